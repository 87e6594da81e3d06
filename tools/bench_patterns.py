#!/usr/bin/env python3
"""Site-pattern compression at Makona scale (1610 taxa x 18,992 sites, SURVEY.md 8d cfg 4): the GPU hash-table build
(host arrays in, host arrays out, synchronous) next to numpy's sort-based unique on the host.  One JSON line."""
import json, os, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import beast_mcmc_b200  # noqa
from beast_mcmc_b200 import beagle
from harness import evomodel as em

TAXA, SITES = int(os.environ.get("TAXA", 1610)), int(os.environ.get("SITES", 18992))
tree = em.Tree.coalescent(TAXA, 0.0025, 3)
model = em.GTR(1.0, 4.0, 0.7, 1.2, 5.0, 1.0, np.array([0.30, 0.22, 0.24, 0.24]))
site = em.GammaSiteRateModel(shape=0.5, gammaCategoryCount=4)
aln = em.simulate_alignment(tree, model, site, SITES, 11).astype(np.int32)


def best(fn, n=5):
    fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); r = fn(); ts.append(time.perf_counter() - t0)
    return min(ts), r


tg, (pats, w, idx) = best(lambda: beagle.compressSitePatterns(aln))
tc, m = best(lambda: em.Patterns.fromAlignment(aln), 2)
ok = bool(np.array_equal(pats, m.states) and np.array_equal(w, m.weights))
out = {"taxa": TAXA, "sites": SITES, "patterns": int(pats.shape[1]), "gpu_ms": 1e3 * tg, "numpy_unique_ms": 1e3 * tc,
       "identical": ok, "input_MB": aln.nbytes / 1e6,
       "note": "gpu_ms includes the H2D copy of the int32 alignment and the D2H copies of the results"}
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "r01_bench_patterns.json"), "w").write(json.dumps(out) + "\n")
