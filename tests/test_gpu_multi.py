"""GPU tests of the engine's own multi-GPU layer (csrc/multi.cu, SURVEY.md 8e mode B) and of the reduce groups whose sum
is fused into the root kernel.  Every case runs on a ONE-GPU box too: the shard device list may name a device several
times (the shards then share it; the exchange goes through the same code, peer pointers included), and the cases that
need distinct devices skip themselves.

Checked against: the unsharded engine instance and the numpy oracle (1e-10), and BIT-equality with mode A -- the sum, in
shard order, of independent per-shard instances over the Patterns.java:142-169 blocks (what CompoundLikelihood adds up)."""
import ctypes as C
import math
import threading

import numpy as np
import pytest

import helpers as H
from beast_mcmc_b200 import beagle
from harness import evomodel as em, treedatalikelihood as tdl

pytestmark = pytest.mark.gpu

REL = 1e-10
GPU = beagle.BeagleFactory.loadBeagleInstance
ORACLE = H.oracle_factory(report_flags=0)
S_ = tdl.PartialsRescalingScheme


def _rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def _gpu_count():
    import torch
    return torch.cuda.device_count()


def _shard_resource(devices):
    lib = beagle.load_library()
    res = beagle.BeagleFactory.getResourceDetails()
    number = [r.number for r in res if "pattern-sharded" in r.name]
    assert number, [r.name for r in res]
    arr = (C.c_int * len(devices))(*devices)
    assert lib.b200SetShardDevices(arr, len(devices)) == 0
    return number[0]


def _devices(g):
    n = _gpu_count()
    return [k % n for k in range(g)]


@pytest.mark.parametrize("g", [2, 3, 8])
@pytest.mark.parametrize("states,cats,scheme", [(4, 4, S_.NONE), (4, 4, S_.ALWAYS), (20, 2, S_.ALWAYS), (61, 1, S_.NONE)])
def test_sharded_instance_equals_whole_and_mode_a(g, states, cats, scheme):
    # 4 states: more than 64 operations, so that neither side takes the one-launch route of short lists (csrc/incr.cu), whose
    # root reduction has its own (equally deterministic) summation order -- bit-equality is asserted between like paths
    tips, patterns = (100, 1003) if states == 4 else ((14, 203) if states == 20 else (9, 77))
    tree, pats, model, site = H.synthetic_case(tips, patterns, cats, seed=17 + states, stateCount=states)
    res = _shard_resource(_devices(g))
    kw = dict(rescalingScheme=scheme, delayRescalingUntilUnderflow=False)
    sharded = tdl.BeagleDataLikelihoodDelegate(tree, pats, model, site, GPU, resourceList=[res, 0], **kw)
    assert sharded.beagle.getDetails().getResourceNumber() == res
    whole = tdl.BeagleDataLikelihoodDelegate(tree, pats, model, site, GPU, resourceList=[1, 0], **kw)
    oracle = tdl.BeagleDataLikelihoodDelegate(tree, pats, model, site, ORACLE, **kw)
    ls, lw, lo = (tdl.TreeDataLikelihood(d, tree) for d in (sharded, whole, oracle))
    vs, vw, vo = ls.getLogLikelihood(), lw.getLogLikelihood(), lo.getLogLikelihood()
    assert math.isfinite(vo) and _rel(vs, vo) <= REL and _rel(vs, vw) <= REL, (vs, vw, vo)
    # mode A: BEAST's own split, one instance per block, summed in shard order on the host
    parts = []
    for k in range(g):
        d = tdl.BeagleDataLikelihoodDelegate(tree, pats.subSet(k, g), model, site, GPU, resourceList=[1, 0], **kw)
        parts.append(tdl.TreeDataLikelihood(d, tree).getLogLikelihood())
        d.finalize()
    total = 0.0
    for v in parts:
        total += v
    assert vs == total, (vs, total, parts)
    # gathered per-pattern outputs
    assert np.allclose(sharded.getSiteLogLikelihoods(), oracle.getSiteLogLikelihoods(), rtol=1e-10, atol=1e-12)
    assert np.array_equal(sharded.getSiteLogLikelihoods(), whole.getSiteLogLikelihoods())
    for node in (tree.tipCount + 1, tree.root):
        assert np.array_equal(sharded.getPartials(node), whole.getPartials(node)), node
    if scheme == S_.ALWAYS:
        a, b = np.zeros(pats.patternCount), np.zeros(pats.patternCount)
        sharded.beagle.getLogScaleFactors(sharded.scaleBufferIndices[2], a)
        whole.beagle.getLogScaleFactors(whole.scaleBufferIndices[2], b)
        assert np.array_equal(a, b)
    # an incremental update and a store / restore cycle through the sharded instance
    for like in (ls, lo):
        like.storeState()
        t = like.tree
    node = tree.tipCount + 3
    old = tree.height[node]
    tree.height[node] = 0.5 * (max(tree.height[c] for c in tree.child[node]) + tree.height[tree.parent[node]])
    for like in (ls, lo):
        like.updateNodeAndChildren(node)
    assert _rel(ls.getLogLikelihood(), lo.getLogLikelihood()) <= REL
    tree.height[node] = old
    for like in (ls, lo):
        like.restoreState()
    assert ls.getLogLikelihood() == vs
    sharded.finalize()
    whole.finalize()


def test_sharded_instance_with_fewer_patterns_than_shards_and_errors():
    tree, pats, model, site = H.synthetic_case(6, 3, 2, seed=5)
    res = _shard_resource(_devices(5))                        # 5 shards, 3 patterns: two shards are empty
    sharded = tdl.BeagleDataLikelihoodDelegate(tree, pats, model, site, GPU, resourceList=[res, 0], rescalingScheme=S_.NONE)
    oracle = tdl.BeagleDataLikelihoodDelegate(tree, pats, model, site, ORACLE, rescalingScheme=S_.NONE)
    assert _rel(tdl.TreeDataLikelihood(sharded, tree).getLogLikelihood(),
                tdl.TreeDataLikelihood(oracle, tree).getLogLikelihood()) <= REL
    with pytest.raises(beagle.BeagleException) as e:         # *ByPartition is not offered on a sharded instance
        sharded.beagle.setPatternPartitions(1, np.zeros(3, dtype=np.int32))
    assert e.value.errCode == beagle.BeagleErrorCode.NO_IMPLEMENTATION_ERROR
    with pytest.raises(beagle.BeagleException) as e:
        sharded.beagle.updatePartials(np.array([99, -1, -1, 0, 0, 1, 1], dtype=np.int32), 1, -1)
    assert e.value.errCode == beagle.BeagleErrorCode.OUT_OF_RANGE_ERROR
    sharded.finalize()


def test_sharded_gradient_calls():
    """pre-order partials, edge derivatives (sums + per-pattern gather) and cross products through a sharded instance."""
    tree, pats, model, site = H.synthetic_case(12, 301, 4, seed=31)
    res = _shard_resource(_devices(3))
    out = []
    for factory, rl in ((GPU, [res, 0]), (GPU, [1, 0])):
        d = tdl.BeagleDataLikelihoodDelegate(tree, pats, model, site, factory, resourceList=rl, rescalingScheme=S_.NONE,
                                             usePreOrder=True)
        like = tdl.TreeDataLikelihood(d, tree)
        like.getLogLikelihood()
        grad = tdl.DiscreteTraitBranchRateDelegate(tree, d, model)
        g = grad.getGradient()
        cross = tdl.SubstitutionModelCrossProductDelegate(tree, d, model).getCrossProducts()
        out.append((g, cross))
        d.finalize()
    assert np.allclose(out[0][0], out[1][0], rtol=1e-11, atol=1e-13)
    assert np.allclose(out[0][1], out[1][1], rtol=1e-11, atol=1e-13)


def test_reduce_group_of_plain_instances_from_threads():
    """Mode A inside one process with the sum moved into the root kernel: g ordinary instances (one per pattern block),
    b200ExchangeConnectLocal, each driven from its own thread -- every member's calculateRootLogLikelihoods returns the
    JOINT value, bit-equal to the host-side sum in shard order.  Distinct devices when the box has them."""
    g = 4
    tree, pats, model, site = H.synthetic_case(100, 2003, 4, seed=3)      # > 64 operations: the ordinary launches on both sides
    devices = _devices(g)
    shards = [pats.subSet(k, g) for k in range(g)]
    plain = []
    for k in range(g):
        d = tdl.BeagleDataLikelihoodDelegate(tree, shards[k], model, site, GPU, resourceList=[devices[k] + 1, 0], rescalingScheme=S_.NONE)
        plain.append(tdl.TreeDataLikelihood(d, tree).getLogLikelihood())
        d.finalize()
    total = 0.0
    for v in plain:
        total += v
    delegates = [tdl.BeagleDataLikelihoodDelegate(tree, shards[k], model, site, GPU, resourceList=[devices[k] + 1, 0],
                                                  rescalingScheme=S_.NONE) for k in range(g)]
    ids = (C.c_int * g)(*[d.beagle.instance for d in delegates])
    assert beagle.load_library().b200ExchangeConnectLocal(ids, g) == 0
    results = [[] for _ in range(g)]

    def work(k):
        like = tdl.TreeDataLikelihood(delegates[k], tree)
        for _ in range(12):
            like.makeDirty()
            results[k].append(like.getLogLikelihood())

    threads = [threading.Thread(target=work, args=(k,)) for k in range(g)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    for k in range(g):
        assert results[k] == [total] * 12, (k, results[k][:2], total)
    whole = tdl.BeagleDataLikelihoodDelegate(tree, pats, model, site, ORACLE, rescalingScheme=S_.NONE)
    assert _rel(total, tdl.TreeDataLikelihood(whole, tree).getLogLikelihood()) <= REL
    [d.finalize() for d in delegates]


def test_sharded_instance_on_distinct_devices():
    n = _gpu_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    tree, pats, model, site = H.synthetic_case(200, 6000, 4, seed=41)
    res = _shard_resource(list(range(n)))
    sharded = tdl.BeagleDataLikelihoodDelegate(tree, pats, model, site, GPU, resourceList=[res, 0], rescalingScheme=S_.ALWAYS,
                                               delayRescalingUntilUnderflow=False)
    whole = tdl.BeagleDataLikelihoodDelegate(tree, pats, model, site, GPU, resourceList=[1, 0], rescalingScheme=S_.ALWAYS,
                                             delayRescalingUntilUnderflow=False)
    ls, lw = tdl.TreeDataLikelihood(sharded, tree), tdl.TreeDataLikelihood(whole, tree)
    for _ in range(5):
        ls.makeDirty()
        lw.makeDirty()
        assert _rel(ls.getLogLikelihood(), lw.getLogLikelihood()) <= REL
    assert np.array_equal(sharded.getSiteLogLikelihoods(), whole.getSiteLogLikelihoods())
    sharded.finalize()
    whole.finalize()
