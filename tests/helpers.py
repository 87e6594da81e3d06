"""Shared builders for the parity tests: the reference's golden fixtures and the
oracle-backed ``beagleFactory`` (test infrastructure -- may import oracle/)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import beast_mcmc_b200  # noqa: E402
from harness import evomodel as em  # noqa: E402
from harness import treedatalikelihood as tdl  # noqa: E402
from oracle.felsenstein import OracleBeagle  # noqa: E402

GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "primates.json")))


class _Details:
    def __init__(self, flags):
        self.flags = flags


def oracle_factory(extra_flags=0, report_flags=tdl.FLAG_FRAMEWORK_CPU):
    def make(*args):
        args = list(args)
        args[10] |= extra_flags
        inst = OracleBeagle(*args)
        inst.getDetails = lambda: _Details(report_flags)
        return inst
    return make


def primate_patterns():
    g = GOLDEN["primates"]
    return em.Patterns.fromAlignment(em.encode_nucleotides(g["sequences"]))


def primate_tree():
    g = GOLDEN["primates"]
    h = g["tree"]["heights"]
    spec = (((("human", ("chimp", "bonobo", h["chimp_bonobo"]), h["human_cb"]), "gorilla", h["plus_gorilla"]),
             "orangutan", h["plus_orangutan"]), "siamang", h["root"])
    return em.Tree.fromNested(spec, g["taxa"])


def primate_cases():
    """name -> (substitution model, site model, expected logL): the ten cases of
    TreeDataLikelihoodTest.java:116-314."""
    pats = primate_patterns()
    emp = pats.stateFrequencies()
    uni = np.full(4, 0.25)
    exp = GOLDEN["primates"]["expected_logL"]
    G = em.GammaSiteRateModel
    gtr = lambda: em.GTR(1.0, 1.0, 1.0, 1.0, 1.0, 1.0, emp)
    return {
        "JC69": (em.HKY(1.0, uni), G(), exp["JC69"]),
        "K80": (em.HKY(8.0, uni), G(), exp["K80"]),
        "HKY85": (em.HKY(8.0, emp), G(), exp["HKY85"]),
        "HKY85G": (em.HKY(8.0, emp), G(shape=0.5, gammaCategoryCount=4), exp["HKY85G"]),
        "HKY85I": (em.HKY(8.0, emp), G(pInv=0.75), exp["HKY85I"]),
        "HKY85GI": (em.HKY(8.0, emp), G(shape=0.5, gammaCategoryCount=4, pInv=0.75), exp["HKY85GI"]),
        "GTR": (gtr(), G(), exp["GTR"]),
        "GTRI": (gtr(), G(pInv=0.5), exp["GTRI"]),
        "GTRG": (gtr(), G(shape=0.5, gammaCategoryCount=4), exp["GTRG"]),
        "GTRGI": (gtr(), G(shape=0.5, gammaCategoryCount=4, pInv=0.5), exp["GTRGI"]),
    }


def primate_cases_legacy():
    """name -> (substitution model, site model, expected logL): the ten cases of LikelihoodTest.java:86-341 -- the same
    alignment and tree through the older TreeLikelihood, with that test's own parameters and the older site-model rate
    rule (dr.oldevomodel.sitemodel.GammaSiteModel.java:271-311)."""
    pats = primate_patterns()
    emp = pats.stateFrequencies()
    uni = np.full(4, 0.25)
    out = {}
    for name, c in GOLDEN["primates"]["legacy_likelihood_test"].items():
        freqs = emp if c["empiricalFrequencies"] else uni
        model = em.GTR(1.0, 1.0, 1.0, 1.0, 1.0, 1.0, freqs) if c["model"] == "GTR" else em.HKY(c["kappa"], freqs)
        site = em.GammaSiteModel(shape=c["shape"], gammaCategoryCount=4 if c["shape"] is not None else 1, pInv=c["pInv"])
        out[name] = (model, site, c["logL"])
    return out


def _newick_to_tree(newick):
    """Ultrametric newick with branch lengths -> (em.Tree, tip names in taxonN order); heights from the lengths."""
    import re
    tokens = re.findall(r"[(),;]|[^(),;:]+|:[0-9.eE+-]+", newick.strip())
    pos = [0]

    def node():
        if tokens[pos[0]] == "(":
            pos[0] += 1
            kids = [node()]
            while tokens[pos[0]] == ",":
                pos[0] += 1
                kids.append(node())
            assert tokens[pos[0]] == ")" and len(kids) == 2
            pos[0] += 1
            name = None
        else:
            kids, name = None, tokens[pos[0]]
            pos[0] += 1
        length = 0.0
        if pos[0] < len(tokens) and tokens[pos[0]].startswith(":"):
            length = float(tokens[pos[0]][1:])
            pos[0] += 1
        return (name, kids, length)

    root = node()
    names = sorted(re.findall(r"taxon\d+", newick), key=lambda t: int(t[5:]))

    def build(n):
        name, kids, _ = n
        if kids is None:
            return name, 0.0
        (a, ha), (b, hb) = build(kids[0]), build(kids[1])
        h1, h2 = ha + kids[0][2], hb + kids[1][2]
        assert abs(h1 - h2) < 1e-12, "fixture trees are ultrametric"
        return (a, b, h1), h1

    spec, _ = build(root)
    return em.Tree.fromNested(spec, names), names


def msat_cases():
    """[(tree, patterns, model, site model, expected logL)]: MsatFullLikelihoodTest.java:60-190, pinned there to 1e-10.
    Default AsymmetricQuadraticModel = stepwise mutation model (rate 1 to either neighbour state), normalised to one expected
    substitution, uniform stationary root frequencies."""
    import json
    out = []
    for c in json.load(open(os.path.join(ROOT, "tests", "golden", "msat.json")))["cases"]:
        S = c["stateCount"]
        tree, names = _newick_to_tree(c["newick"])
        rates = [1.0 if j == i + 1 else 0.0 for i in range(S) for j in range(i + 1, S)]
        model = em.SubstitutionModel(rates, np.full(S, 1.0 / S))
        pats = em.Patterns(np.asarray(c["pattern"], dtype=np.int32)[:, None], np.ones(1), S)
        out.append((tree, pats, model, em.GammaSiteRateModel(), c["logL"]))
    return out


def tiny_case():
    g = GOLDEN["tiny"]
    pats = em.Patterns.fromAlignment(em.encode_nucleotides(g["sequences"]), unique=False)
    e = g["edges"]
    spec = (("human", "chimp", e["human"]), "gorilla", e["human"] + e["human_chimp"])
    tree = em.Tree.fromNested(spec, g["taxa"])
    # gorilla edge is 0.2 = root height - 0 => consistent with human/chimp path 0.1 + 0.1
    assert abs(tree.branchLength(2) - e["gorilla"]) < 1e-15
    return tree, pats, em.HKY(1.0, np.full(4, 0.25)), em.GammaSiteRateModel(), g["expected_logL"]


def synthetic_case(tips, patterns, categories=4, seed=11, stateCount=4, rootHeight=0.1):
    """Small seeded instance of the cfg-2/cfg-3 recipe."""
    tree = em.Tree.coalescent(tips, rootHeight, seed)
    if stateCount == 4:
        model = em.GTR(1.0, 4.0, 0.7, 1.2, 5.0, 1.0, np.array([0.30, 0.22, 0.24, 0.24]))
    elif stateCount == 61:
        model = em.MG94HKYCodonModel(1.0, 0.3, 2.0)
    else:
        rng = np.random.default_rng(seed)
        pi = rng.dirichlet(np.full(stateCount, 5.0))
        model = em.SubstitutionModel(rng.uniform(0.2, 3.0, stateCount * (stateCount - 1) // 2), pi)
    site = em.GammaSiteRateModel(shape=0.5, gammaCategoryCount=categories) if categories > 1 \
        else em.GammaSiteRateModel()
    aln = em.simulate_alignment(tree, model, site, patterns, seed + 1)
    pats = em.Patterns(aln, np.random.default_rng(seed + 2).integers(1, 5, patterns).astype(np.float64),
                       stateCount)
    return tree, pats, model, site
