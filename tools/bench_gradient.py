#!/usr/bin/env python3
"""Timing of the gradient route on one GPU next to the likelihood it extends: post-order
evaluation, pre-order traversal (updatePrePartials), branch-length gradient (calculateEdgeDifferentials) and the
substitution-model cross products (calculateCrossProductDifferentials).  Synchronous calls through the C ABI with
host buffers; wall-clock per call after warm-up.  Prints one JSON line."""
import json, os, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import beast_mcmc_b200  # noqa
from beast_mcmc_b200 import beagle
from harness import evomodel as em, treedatalikelihood as tdl

import bench

WORK = os.environ.get("WORKLOAD", "gtr_g4_1000x10k")          # any bench.py workload name
STEPS = int(os.environ.get("STEPS", 30))
w, tree, pats, model, site = bench.build_workload(WORK, 0, {})
taxa, patterns, cats = w["taxa"], w["patterns"], w["categories"]
d = tdl.BeagleDataLikelihoodDelegate(tree, pats, model, site, beagle.BeagleFactory.loadBeagleInstance,
                                     resourceList=[1, 0], rescalingScheme=tdl.PartialsRescalingScheme.NONE,
                                     usePreOrder=True)
like = tdl.TreeDataLikelihood(d, tree)
g = tdl.SubstitutionModelCrossProductDelegate(tree, d, model)


def timed(fn, steps=STEPS):
    """(host wall-clock ms per call, device ms per call from the engine's CUDA-event hooks)"""
    for _ in range(3):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    wall = (time.perf_counter() - t0) / steps * 1e3
    d.beagle.setKernelTiming(True)
    for _ in range(steps):
        fn()
    dev = sum(d.beagle.getKernelTiming(k)[0] for k in range(3)) / steps
    d.beagle.setKernelTiming(False)
    return {"wall": round(wall, 4), "device": round(dev, 4)}


def evaluate():
    like.updateNode[:] = True
    like.likelihoodKnown = False
    return like.getLogLikelihood()


logl = evaluate()
nodes = [n for n in range(tree.nodeCount) if n != tree.root]
post = np.asarray([d.getPartialBufferIndex(n) for n in nodes], dtype=np.int32)
pre = np.asarray([g.getPreOrderPartialIndex(n) for n in nodes], dtype=np.int32)
der = np.full(len(nodes), g.firstDerivativeMatrixIndex, dtype=np.int32)
lengths = np.asarray([tree.branchLength(n) for n in nodes])
zero = np.zeros(1, dtype=np.int32)
first, sq = np.zeros(len(nodes)), np.zeros(len(nodes))
cross = np.zeros(d.stateCount ** 2)
g.cacheDifferentialMassMatrix()
out = {
    "workload": WORK, "taxa": taxa, "patterns": patterns, "states": d.stateCount, "categories": cats, "logL": logl,
    "ms_likelihood": timed(evaluate),
    "ms_prepartials": timed(g.simulate),
    "ms_edge_differentials": timed(lambda: d.beagle.calculateEdgeDifferentials(post, pre, der, zero, len(nodes), None,
                                                                               first, sq)),
    "ms_cross_products": timed(lambda: d.beagle.calculateCrossProductDifferentials(post, pre, zero, zero, lengths,
                                                                                    len(nodes), cross, None)),
    "timing": "wall = host wall-clock per synchronous call (python driver included); device = CUDA-event time of the engine kernels in that call",
}
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"r01_bench_gradient_{WORK}.json"), "w") as f:
    f.write(json.dumps(out) + "\n")
