"""Host-side producers of the numbers BEAST hands to BEAGLE (the L4 "model inputs" row of
SURVEY.md section 2.1): data types, site patterns, substitution-model eigen systems,
discretised site-rate categories, trees and a sequence simulator for synthetic workloads.

Nothing here is on the measured path: these objects only build the inputs (tip states,
pattern weights, Evec/Ievc/Eval, category rates/weights, branch lengths) for the tests,
``__graft_entry__.smoke`` and ``bench.py``.  Reference statements followed:

  * nucleotide state codes ........ src/dr/evolution/datatype/Nucleotides.java:51-105
  * unique-pattern compression .... src/dr/evolution/alignment/SitePatterns.java:226-340
  * contiguous pattern sharding ... src/dr/evolution/alignment/Patterns.java:142-169
  * Q set-up and normalisation .... src/dr/evomodel/substmodel/BaseSubstitutionModel.java:256-325
  * eigen-system array layout ..... src/dr/evomodel/substmodel/EigenDecomposition.java:41-121 (row-major)
  * MG94xHKY codon rates .......... src/dr/evomodel/substmodel/codon/MG94HKYCodonModel.java:150-190
  * gamma rate categories ......... src/dr/evomodel/siteratemodel/GammaSiteRateModel.java:233-272,445-472
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# ----------------------------------------------------------------------------------------------
# data types
# ----------------------------------------------------------------------------------------------
_NUC_CHARS = "ACGTURYMWSKBDHVN?-"          # state numbers 0..17 (Nucleotides.java:60-62)
NUCLEOTIDE_UNKNOWN = 16
NUCLEOTIDE_GAP = 17
_NUC_AMBIG = ["A", "C", "G", "T", "T", "AG", "CT", "AC", "AT", "CG", "GT",
              "CGT", "AGT", "ACT", "ACG", "ACGT", "ACGT", "ACGT"]


def nucleotide_state(ch: str) -> int:
    """char -> state code; letters outside the table map to '?', anything else to '-'."""
    c = ch.upper()
    if c == "U":
        return 3
    k = _NUC_CHARS.find(c)
    if k >= 0:
        return k
    return NUCLEOTIDE_UNKNOWN if c.isalpha() else NUCLEOTIDE_GAP


def nucleotide_state_set(state: int) -> np.ndarray:
    out = np.zeros(4)
    for ch in _NUC_AMBIG[state]:
        out["ACGT".index(ch)] = 1.0
    return out


def encode_nucleotides(seqs) -> np.ndarray:
    """list of equal-length strings -> int32 [taxa][sites]."""
    return np.array([[nucleotide_state(c) for c in s] for s in seqs], dtype=np.int32)


# ----------------------------------------------------------------------------------------------
# site patterns
# ----------------------------------------------------------------------------------------------
@dataclass
class Patterns:
    """Unique site patterns with weights. ``states`` is int32 [taxa][patterns]."""
    states: np.ndarray
    weights: np.ndarray
    stateCount: int = 4

    @property
    def patternCount(self) -> int:
        return self.states.shape[1]

    @property
    def taxonCount(self) -> int:
        return self.states.shape[0]

    @staticmethod
    def fromAlignment(states: np.ndarray, stateCount: int = 4, unique: bool = True) -> "Patterns":
        states = np.asarray(states, dtype=np.int32)
        if not unique:
            return Patterns(states.copy(), np.ones(states.shape[1]), stateCount)
        # first-occurrence order, like SitePatterns' incremental insertion
        cols, first, counts = np.unique(states.T, axis=0, return_index=True, return_counts=True)
        order = np.argsort(first, kind="stable")
        return Patterns(np.ascontiguousarray(cols[order].T), counts[order].astype(np.float64), stateCount)

    def subSet(self, subSet: int, subSetCount: int) -> "Patterns":
        """Patterns.java:142-169: contiguous blocks, first (P mod n) shards get one extra."""
        div, rem = divmod(self.patternCount, subSetCount)
        start = sum(div + (1 if i < rem else 0) for i in range(subSet))
        n = div + (1 if subSet < rem else 0)
        return Patterns(np.ascontiguousarray(self.states[:, start:start + n]),
                        self.weights[start:start + n].copy(), self.stateCount)

    def stateFrequencies(self) -> np.ndarray:
        """Empirical frequencies over unambiguous states, weighted by pattern multiplicity."""
        f = np.zeros(self.stateCount)
        for s in range(self.stateCount):
            f[s] = ((self.states == s) * self.weights[None, :]).sum()
        return f / f.sum()


# ----------------------------------------------------------------------------------------------
# substitution models -> eigen systems
# ----------------------------------------------------------------------------------------------
@dataclass
class EigenDecomposition:
    """Row-major Evec[S][S], Ievc[S][S], Eval[S] (or [2S] real||imag), as BEAST passes them."""
    Evec: np.ndarray
    Ievc: np.ndarray
    Eval: np.ndarray


class SubstitutionModel:
    """General time-reversible model on S states: q_ij = r_ij * pi_j, normalised to one
    expected substitution per unit time (BaseSubstitutionModel.java:256-325)."""

    def __init__(self, relativeRates, frequencies):
        self.pi = np.asarray(frequencies, dtype=np.float64)
        self.S = self.pi.shape[0]
        self.rates = np.asarray(relativeRates, dtype=np.float64)   # upper triangle, row-major
        assert self.rates.shape[0] == self.S * (self.S - 1) // 2
        self._eigen = None

    def canReturnComplexDiagonalization(self) -> bool:
        return False

    def getFrequencies(self) -> np.ndarray:
        return self.pi

    def infinitesimalMatrix(self) -> np.ndarray:
        S = self.S
        q = np.zeros((S, S))
        iu = np.triu_indices(S, 1)
        q[iu] = self.rates
        q = q + q.T
        q = q * self.pi[None, :]
        np.fill_diagonal(q, 0.0)
        np.fill_diagonal(q, -q.sum(axis=1))
        norm = -(np.diag(q) * self.pi).sum()
        return q / norm

    def getEigenDecomposition(self) -> EigenDecomposition:
        if self._eigen is None:
            q = self.infinitesimalMatrix()
            # reversible => similar to a symmetric matrix; guarantees a real system
            sq = np.sqrt(self.pi)
            b = (sq[:, None] * q) / sq[None, :]
            b = 0.5 * (b + b.T)
            lam, u = np.linalg.eigh(b)
            evec = u / sq[:, None]
            ievc = u.T * sq[None, :]
            self._eigen = EigenDecomposition(np.ascontiguousarray(evec), np.ascontiguousarray(ievc), lam.copy())
        return self._eigen


class HKY(SubstitutionModel):
    def __init__(self, kappa, frequencies):
        super().__init__([1.0, kappa, 1.0, 1.0, kappa, 1.0], frequencies)


class GTR(SubstitutionModel):
    def __init__(self, ac, ag, at, cg, ct, gt, frequencies):
        super().__init__([ac, ag, at, cg, ct, gt], frequencies)


_CODON_NUC = "ACGT"
_STOP = {"TAA", "TAG", "TGA"}
_AA = dict(zip(
    [a + b + c for a in "TCAG" for b in "TCAG" for c in "TCAG"],
    "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"))
SENSE_CODONS = [a + b + c for a in _CODON_NUC for b in _CODON_NUC for c in _CODON_NUC
                if a + b + c not in _STOP]


class MG94HKYCodonModel(SubstitutionModel):
    """61-state Muse-Gaut x HKY codon model (universal code).  Single-nucleotide changes only:
    synonymous transition alpha*kappa, synonymous transversion alpha, non-synonymous
    transition beta*kappa, non-synonymous transversion beta (MG94HKYCodonModel.java:150-190)."""

    def __init__(self, alpha, beta, kappa, codonFrequencies=None):
        n = len(SENSE_CODONS)
        if codonFrequencies is None:
            codonFrequencies = np.full(n, 1.0 / n)
        rates = []
        transitions = {("A", "G"), ("G", "A"), ("C", "T"), ("T", "C")}
        for i in range(n):
            for j in range(i + 1, n):
                ci, cj = SENSE_CODONS[i], SENSE_CODONS[j]
                diff = [k for k in range(3) if ci[k] != cj[k]]
                if len(diff) != 1:
                    rates.append(0.0)
                    continue
                k = diff[0]
                ts = (ci[k], cj[k]) in transitions
                syn = _AA[ci] == _AA[cj]
                r = (alpha if syn else beta) * (kappa if ts else 1.0)
                rates.append(r)
        super().__init__(rates, codonFrequencies)


# ----------------------------------------------------------------------------------------------
# site rate model
# ----------------------------------------------------------------------------------------------
class GammaSiteRateModel:
    """Median-quantile discretised gamma (+ invariant) categories, normalised over ALL
    categories exactly as GammaSiteRateModel.java:233-272,445-472 does."""

    def __init__(self, shape=None, gammaCategoryCount=1, pInv=None, mu=1.0):
        from scipy.stats import gamma as _gamma
        offset = 0
        if shape is None:
            gammaCategoryCount = 1
        n = gammaCategoryCount + (1 if pInv is not None else 0)
        rates = np.zeros(n)
        props = np.zeros(n)
        if pInv is not None:
            rates[0] = 0.0
            props[0] = pInv
            offset = 1
        if shape is not None:
            k = n - offset
            for i in range(k):
                rates[i + offset] = _gamma.ppf((2.0 * i + 1.0) / (2.0 * k), a=shape, scale=1.0 / shape)
                props[i + offset] = 1.0
            mean = rates.sum() / n
            rates = rates / mean
            props = props / props.sum()
        elif offset > 0:
            rates[offset] = 2.0
            props[offset] = 1.0 - props[0]
        else:
            rates[0] = 1.0
            props[0] = 1.0
        self.rates = rates * mu
        self.proportions = props

    def getCategoryCount(self) -> int:
        return self.rates.shape[0]

    def getCategoryRates(self) -> np.ndarray:
        return self.rates

    def getCategoryProportions(self) -> np.ndarray:
        return self.proportions


class GammaSiteModel(GammaSiteRateModel):
    """The older site model (dr.oldevomodel.sitemodel.GammaSiteModel.java:271-311) that LikelihoodTest.java drives: same
    median-quantile gamma categories, but the rates are normalised so that the proportion-weighted mean over ALL categories,
    the invariant one included, is one (mean = pVariable * sum / K), and the +I-only model uses rate 1 / pVariable."""

    def __init__(self, shape=None, gammaCategoryCount=1, pInv=None, mu=1.0):
        from scipy.stats import gamma as _gamma
        if shape is None:
            gammaCategoryCount = 1
        cat = 1 if pInv is not None else 0
        n = gammaCategoryCount + cat
        rates, props = np.zeros(n), np.zeros(n)
        pVar = 1.0
        if pInv is not None:
            props[0] = pInv
            pVar = 1.0 - pInv
        if shape is not None:
            k = gammaCategoryCount
            for i in range(k):
                rates[i + cat] = _gamma.ppf((2.0 * i + 1.0) / (2.0 * k), a=shape, scale=1.0 / shape)
                props[i + cat] = pVar / k
            rates[cat:] /= (pVar * rates[cat:].sum()) / k
        else:
            rates[cat] = 1.0 / pVar
            props[cat] = pVar
        self.rates = rates * mu
        self.proportions = props


# ----------------------------------------------------------------------------------------------
# trees
# ----------------------------------------------------------------------------------------------
@dataclass
class Tree:
    """Rooted binary tree; nodes 0..N-1 are tips, N..2N-2 internal (BEAST numbering)."""
    parent: np.ndarray            # int [2N-1], -1 at root
    child: np.ndarray             # int [2N-1][2], -1 for tips
    height: np.ndarray            # float [2N-1]
    root: int
    branchRate: np.ndarray = None  # optional per-node clock-rate multipliers

    @property
    def nodeCount(self) -> int:
        return self.parent.shape[0]

    @property
    def tipCount(self) -> int:
        return (self.nodeCount + 1) // 2

    def isExternal(self, n: int) -> bool:
        return n < self.tipCount

    def branchLength(self, n: int) -> float:
        """rate * (parentHeight - nodeHeight)  (TreeTraversal.java:107-124)."""
        r = 1.0 if self.branchRate is None else self.branchRate[n]
        return r * (self.height[self.parent[n]] - self.height[n])

    def copy(self) -> "Tree":
        return Tree(self.parent.copy(), self.child.copy(), self.height.copy(), self.root,
                    None if self.branchRate is None else self.branchRate.copy())

    def depth(self) -> int:
        d = np.zeros(self.nodeCount, dtype=np.int64)
        best = 0
        stack = [self.root]
        while stack:
            n = stack.pop()
            for c in self.child[n]:
                if c >= 0:
                    d[c] = d[n] + 1
                    best = max(best, d[c])
                    stack.append(c)
        return int(best)

    @staticmethod
    def fromNested(spec, tipNames) -> "Tree":
        """spec: nested tuples (left, right, height) with tip names at the leaves."""
        N = len(tipNames)
        parent = -np.ones(2 * N - 1, dtype=np.int64)
        child = -np.ones((2 * N - 1, 2), dtype=np.int64)
        height = np.zeros(2 * N - 1)
        nxt = [N]

        def rec(s):
            if isinstance(s, str):
                return tipNames.index(s)
            l, r, h = s
            a, b = rec(l), rec(r)
            k = nxt[0]
            nxt[0] += 1
            child[k] = (a, b)
            parent[a] = parent[b] = k
            height[k] = h
            return k

        root = rec(spec)
        return Tree(parent, child, height, root)

    @staticmethod
    def coalescent(tipCount: int, rootHeight: float, seed: int) -> "Tree":
        """Kingman coalescent topology/heights (isochronous tips), rescaled to ``rootHeight``."""
        rng = np.random.default_rng(seed)
        N = tipCount
        parent = -np.ones(2 * N - 1, dtype=np.int64)
        child = -np.ones((2 * N - 1, 2), dtype=np.int64)
        height = np.zeros(2 * N - 1)
        active = list(range(N))
        t = 0.0
        for k in range(N, 2 * N - 1):
            n = len(active)
            t += rng.exponential(1.0 / (n * (n - 1) / 2.0))
            i, j = rng.choice(n, size=2, replace=False)
            a, b = active[i], active[j]
            child[k] = (a, b)
            parent[a] = parent[b] = k
            height[k] = t
            active = [x for x in active if x != a and x != b] + [k]
        height *= rootHeight / height[2 * N - 2]
        return Tree(parent, child, height, 2 * N - 2)


# ----------------------------------------------------------------------------------------------
# sequence simulation (synthetic workloads; the reference's analogue is
# src/dr/app/beagle/tools/BeagleSequenceSimulator.java, out of scope and not followed)
# ----------------------------------------------------------------------------------------------
def transition_probabilities(eig: EigenDecomposition, t: float) -> np.ndarray:
    return np.abs((eig.Evec * np.exp(eig.Eval * t)[None, :]) @ eig.Ievc)


def simulate_alignment(tree: Tree, model: SubstitutionModel, siteModel: GammaSiteRateModel,
                       siteCount: int, seed: int) -> np.ndarray:
    """int32 [tips][sites]; each site draws a rate category, then evolves down the tree."""
    rng = np.random.default_rng(seed)
    S = model.S
    eig = model.getEigenDecomposition()
    rates = siteModel.getCategoryRates()
    cat = rng.choice(len(rates), size=siteCount, p=siteModel.getCategoryProportions())
    states = np.zeros((tree.nodeCount, siteCount), dtype=np.int32)
    states[tree.root] = rng.choice(S, size=siteCount, p=model.getFrequencies() / model.getFrequencies().sum())
    order = [tree.root]
    for n in order:
        for c in tree.child[n]:
            if c >= 0:
                order.append(int(c))
    u = rng.random((tree.nodeCount, siteCount))
    for n in order[1:]:
        t = tree.branchLength(n)
        par = states[tree.parent[n]]
        for ci, r in enumerate(rates):
            sel = np.nonzero(cat == ci)[0]
            if sel.size == 0:
                continue
            cdf = np.cumsum(transition_probabilities(eig, t * r), axis=1)
            cdf /= cdf[:, -1:]
            rows = cdf[par[sel]]
            states[n, sel] = (u[n, sel][:, None] > rows).sum(axis=1).clip(0, S - 1)
    return states[:tree.tipCount].copy()


def synthetic_patterns(tree: Tree, model: SubstitutionModel, siteModel: GammaSiteRateModel,
                       patternCount: int, seed: int, batch: int = 0) -> Patterns:
    """Keep simulating sites until exactly ``patternCount`` unique patterns exist
    (SURVEY.md section 8d recipe); weights are the multiplicities."""
    batch = batch or max(1024, patternCount)
    seen = {}
    cols = []
    weights = []
    k = 0
    while len(cols) < patternCount:
        block = simulate_alignment(tree, model, siteModel, batch, seed + 7919 * k)
        k += 1
        for col in block.T:
            key = col.tobytes()
            idx = seen.get(key)
            if idx is None:
                if len(cols) < patternCount:
                    seen[key] = len(cols)
                    cols.append(col.copy())
                    weights.append(1.0)
            else:
                weights[idx] += 1.0
        if k > 10000:
            raise RuntimeError("could not reach the requested number of unique patterns")
    return Patterns(np.ascontiguousarray(np.array(cols, dtype=np.int32).T),
                    np.array(weights), model.S)
