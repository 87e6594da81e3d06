#!/usr/bin/env python3
"""Fixture generator (runs only where /root/reference exists).

Extracts the reference's own golden INPUTS for the hot path into tests/golden/primates.json:
  * the 6-primate x 768-nt alignment   src/test/dr/inference/trace/TraceCorrelationAssert.java:192-198
  * the fixed-height primate tree       TraceCorrelationAssert.java:145-190
  * the ten expected log-likelihoods    src/test/dr/evomodel/treedatalikelihood/TreeDataLikelihoodTest.java:131-314
  * ten more, with their parameters      src/test/dr/evomodel/treelikelihood/LikelihoodTest.java:86-341 (older TreeLikelihood,
                                         dr.oldevomodel.sitemodel.GammaSiteModel rate rule)
  * the BEAGLE tiny test (3 taxa, JC69) src/test/dr/app/beagle/TinyTest.java:101-107 + lib/beagle.jar BeagleFactory.main
"""
import json, re, sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"


def seqs_after(path, marker, count):
    txt = open(path).read()
    tail = txt[txt.index(marker):]
    out = re.findall(r'"([ACGTUacgtu\-\?NRYMWSKBDHVnrymwskbdhv]{100,})"', tail)
    return out[:count]


def main():
    tca = f"{REF}/src/test/dr/inference/trace/TraceCorrelationAssert.java"
    prim = seqs_after(tca, "PRIMATES_TAXON_SEQUENCE", 6)
    tiny = seqs_after(f"{REF}/src/test/dr/app/beagle/TinyTest.java", "static private String sequences", 3)
    test = open(f"{REF}/src/test/dr/evomodel/treedatalikelihood/TreeDataLikelihoodTest.java").read()
    expected = dict(re.findall(r'assertEquals\("treeLikelihood(\w+)", format.format\((-[\d.]+)\)', test))
    heights = re.findall(r"setHeight\((\d+\.\d+)\)", open(tca).read())[:5]
    # the same data through the older pure-Java TreeLikelihood: ten more pinned values with their own parameters
    lt = open(f"{REF}/src/test/dr/evomodel/treelikelihood/LikelihoodTest.java").read()
    legacy = {}
    for body in re.split(r"public void testLikelihood", lt)[1:]:
        name = re.match(r"(\w+)\(\)", body).group(1)
        val = re.search(r'assertEquals\("treeLikelihood\w+", format.format\((-[\d.]+)\)', body)
        if not val:
            continue
        par = lambda key: (lambda m: float(m.group(1)) if m else None)(re.search(key + r",\s*([\d.]+),", body))
        legacy[name] = {"logL": float(val.group(1)), "kappa": par(r"HKYParser\.KAPPA"),
                        "shape": par(r"GammaSiteModelParser\.GAMMA_SHAPE"),
                        "pInv": par(r"GammaSiteModelParser\.PROPORTION_INVARIANT"),
                        "empiricalFrequencies": "alignment.getStateFrequencies()" in body,
                        "model": "GTR" if "new GTR(" in body else "HKY"}
    assert len(legacy) == 10, legacy.keys()
    out = {
        "primates": {
            "taxa": ["human", "chimp", "bonobo", "gorilla", "orangutan", "siamang"],
            "sequences": prim,
            "tree": {"heights": {"chimp_bonobo": float(heights[0]), "human_cb": float(heights[1]),
                                 "plus_gorilla": float(heights[2]), "plus_orangutan": float(heights[3]),
                                 "root": float(heights[4])}},
            "expected_logL": {k: float(v) for k, v in expected.items()},
            "legacy_likelihood_test": legacy,
        },
        "tiny": {
            "taxa": ["human", "chimp", "gorilla"],
            "sequences": tiny,
            "edges": {"human": 0.1, "chimp": 0.1, "human_chimp": 0.1, "gorilla": 0.2},
            "expected_logL": -1574.63623,
        },
    }
    assert len(prim) == 6 and all(len(s) == 768 for s in prim), [len(s) for s in prim]
    assert len(tiny) == 3 and len(set(map(len, tiny))) == 1
    assert len(expected) == 10, expected
    json.dump(out, open(__file__.rsplit("/", 1)[0] + "/primates.json", "w"), indent=1)
    print({k: v for k, v in out["primates"]["expected_logL"].items()}, len(tiny[0]))


if __name__ == "__main__":
    main()
