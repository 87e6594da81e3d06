#!/usr/bin/env python3
"""cfg 5 (BASELINE.json configs[4]) on ONE GPU: 8 gene-like partitions sharing a 2000-taxon tree, evaluated
(a) through ONE instance with the *ByPartition calls (the route MultiPartitionDataLikelihoodDelegate takes on GPUs,
MPDLD:744-1207) and (b) as 8 single-partition instances evaluated one after the other.  Prints one JSON line.
The 8-GPU form of this config (one partition per GPU, NCCL sum) is what `bench.py --gpus 8` measures."""
import json, os, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import beast_mcmc_b200  # noqa
from beast_mcmc_b200 import beagle
from harness import evomodel as em, treedatalikelihood as tdl

TAXA = int(os.environ.get("TAXA", 2000))
SITES = [2341, 2341, 2233, 1778, 1565, 1413, 1027, 890]
STEPS = int(os.environ.get("STEPS", 200))

tree = em.Tree.coalescent(TAXA, 0.05, 5)
N, n = tree.tipCount, tree.nodeCount
parts, models, sites = [], [], []
for k, ns in enumerate(SITES):
    rng = np.random.default_rng(10 + k)
    model = em.GTR(*rng.uniform(0.5, 4.0, 6), rng.dirichlet(np.full(4, 20.0)))
    site = em.GammaSiteRateModel(shape=float(rng.uniform(0.3, 1.0)), gammaCategoryCount=4)
    aln = em.simulate_alignment(tree, model, site, ns, 100 + k)
    parts.append(em.Patterns.fromAlignment(aln))
    models.append(model); sites.append(site)
K = len(parts)
counts = [p.patternCount for p in parts]
allStates = np.concatenate([p.states for p in parts], axis=1)
allWeights = np.concatenate([p.weights for p in parts])
P = allStates.shape[1]
like = tdl.TreeDataLikelihood.__new__(tdl.TreeDataLikelihood)
like.tree, like.traversalType, like.updateNode = tree, "REVERSE_LEVEL_ORDER", np.ones(n, dtype=bool)
like._dispatch()
branches = like.branchOperations
nodeOps = like.nodeOperations

# ---- (a) one instance, by-partition calls ---------------------------------------------------------------
b = beagle.BeagleFactory.loadBeagleInstance(N, n, N, 4, P, K, K * n, 4, 1, [1, 0], 0, 0)
for t in range(N):
    b.setTipStates(t, allStates[t])
b.setPatternWeights(allWeights)
b.setPatternPartitions(K, np.repeat(np.arange(K, dtype=np.int32), counts))
eig, rate, prob, lens = [], [], [], []
for k in range(K):
    e = models[k].getEigenDecomposition()
    b.setEigenDecomposition(k, e.Evec, e.Ievc, e.Eval)
    b.setCategoryRatesWithIndex(k, sites[k].getCategoryRates())
    b.setCategoryWeights(k, sites[k].getCategoryProportions())
    b.setStateFrequencies(k, models[k].getFrequencies())
    for node, t in branches:
        eig.append(k); rate.append(k); prob.append(node + k * n); lens.append(t)
eig, rate, prob = (np.array(x, dtype=np.int32) for x in (eig, rate, prob))
lens = np.array(lens)
ops = []
for node, c1, c2 in nodeOps:
    for k in range(K):
        ops += [node, -1, -1, c1, c1 + k * n, c2, c2 + k * n, k, -1]
ops = np.array(ops, dtype=np.int32)
root = np.full(K, tree.root, dtype=np.int32)
idx = np.arange(K, dtype=np.int32)
none = np.full(K, -1, dtype=np.int32)
byPart, total = np.zeros(K), np.zeros(1)


def step_single():
    b.updateTransitionMatricesWithMultipleModels(eig, rate, prob, None, None, lens, len(lens))
    b.updatePartialsByPartition(ops, len(ops) // 9)
    b.calculateRootLogLikelihoodsByPartition(root, idx, idx, none, idx, K, 1, byPart, total)
    return total[0]


for _ in range(5):
    joint_single = step_single()
t0 = time.perf_counter()
for _ in range(STEPS):
    step_single()
dt_single = (time.perf_counter() - t0) / STEPS
b.finalize()

# ---- (b) eight single-partition instances -------------------------------------------------------------------
insts = []
for k in range(K):
    i = beagle.BeagleFactory.loadBeagleInstance(N, n, N, 4, counts[k], 1, n, 4, 1, [1, 0], 0, 0)
    for t in range(N):
        i.setTipStates(t, parts[k].states[t])
    i.setPatternWeights(parts[k].weights)
    e = models[k].getEigenDecomposition()
    i.setEigenDecomposition(0, e.Evec, e.Ievc, e.Eval)
    i.setCategoryRates(sites[k].getCategoryRates())
    i.setCategoryWeights(0, sites[k].getCategoryProportions())
    i.setStateFrequencies(0, models[k].getFrequencies())
    insts.append(i)
bn = np.array([node for node, _ in branches], dtype=np.int32)
bl = np.array([t for _, t in branches])
ops7 = np.array([v for node, c1, c2 in nodeOps for v in (node, -1, -1, c1, c1, c2, c2)], dtype=np.int32)
one, z, m1, out = np.array([tree.root], dtype=np.int32), np.zeros(1, dtype=np.int32), np.full(1, -1, dtype=np.int32), np.zeros(1)


def step_separate():
    tot = 0.0
    for i in insts:
        i.updateTransitionMatrices(0, bn, None, None, bl, len(bl))
        i.updatePartials(ops7, len(nodeOps), -1)
    for i in insts:
        i.calculateRootLogLikelihoods(one, z, z, m1, 1, out)
        tot += out[0]
    return tot


for _ in range(5):
    joint_sep = step_separate()
t0 = time.perf_counter()
for _ in range(STEPS):
    step_separate()
dt_sep = (time.perf_counter() - t0) / STEPS
print(json.dumps({"workload": "flu8_like: 8 partitions, %d taxa, GTR+G4 each" % TAXA, "patterns": counts, "total_patterns": int(P),
                  "single_instance_by_partition": {"joint_evals_per_s": 1 / dt_single, "ms_per_eval": 1e3 * dt_single, "logL": joint_single},
                  "eight_instances_one_gpu": {"joint_evals_per_s": 1 / dt_sep, "ms_per_eval": 1e3 * dt_sep, "logL": joint_sep},
                  "rel_diff": abs(joint_single - joint_sep) / abs(joint_sep), "steps": STEPS}))
