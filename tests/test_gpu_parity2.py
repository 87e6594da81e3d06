"""GPU parity tests, round 2: every exported entry point the round-1 suite left untested, driven through the C ABI and
checked against the ORACLE (numpy restatement / C port) -- never GPU against GPU:

  * the *ByPartition route through a re-enactment of MultiPartitionDataLikelihoodDelegate (>= 3 unequal partitions,
    S in {4, 20, 61}, with and without per-partition rescaling, store/restore, single-partition updates)
  * updatePrePartialsByPartition
  * removeScaleFactors[ByPartition], copyScaleFactors, getScaleFactors, getLogScaleFactors on cumulative buffers
  * setTipPartials through the BeagleTreeLikelihood re-enactment (tip-error model), its in-call underflow retry and the
    ascertainment correction from getSiteLogLikelihoods
  * updatePartials with a non-NONE cumulativeScaleIndex (and field 9 of the ByPartition tuple) on trees with many
    concurrent subtrees
  * BASELINE configs 3 (codon) and 4 (Makona-like) at FULL size against the C port: root value, every site
    log-likelihood and a sample of internal nodes' partials
  * two instances on two different resources driven from two threads (needs >= 2 GPUs)

Tolerance: 1e-10 relative on log-likelihoods (north_star), 1e-9 relative on partials with an absolute floor for entries
that are cancellation residue (as in test_gpu_parity.py)."""
import math
import threading

import numpy as np
import pytest

import helpers as H
from beast_mcmc_b200 import beagle
from harness import evomodel as em, treedatalikelihood as tdl
from harness.beagletreelikelihood import BeagleTreeLikelihood, TipPartialsModel
from harness.multipartition import MultiPartitionDataLikelihoodDelegate

pytestmark = pytest.mark.gpu

REL = 1e-10
GPU = beagle.BeagleFactory.loadBeagleInstance
ORACLE = H.oracle_factory(report_flags=0)
S_ = tdl.PartialsRescalingScheme
NONE = -1


def _rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def _i(v):
    return np.asarray(v, dtype=np.int32)


def _partition_case(stateCount, sizes, tips, cats, seed):
    tree, pats, model, site = H.synthetic_case(tips, sum(sizes), cats, seed=seed, stateCount=stateCount)
    cuts = np.cumsum((0,) + tuple(sizes))
    parts = [em.Patterns(np.ascontiguousarray(pats.states[:, a:b]), pats.weights[a:b].copy(), stateCount)
             for a, b in zip(cuts[:-1], cuts[1:])]
    rng = np.random.default_rng(seed)
    models, sites = [], []
    for k in range(len(sizes)):
        if stateCount == 4:
            models.append(em.HKY(1.5 + k, rng.dirichlet(np.full(4, 8.0))))
        elif stateCount == 61:
            models.append(em.MG94HKYCodonModel(1.0, 0.2 + 0.1 * k, 2.0 + k))
        else:
            models.append(em.SubstitutionModel(rng.uniform(0.2, 3.0, stateCount * (stateCount - 1) // 2),
                                               rng.dirichlet(np.full(stateCount, 5.0))))
        sites.append(em.GammaSiteRateModel(shape=0.4 + 0.3 * k, gammaCategoryCount=cats) if cats > 1
                     else em.GammaSiteRateModel())
    return tree, parts, models, sites


# ---- *ByPartition through the MultiPartitionDataLikelihoodDelegate re-enactment ---------------------------------------
@pytest.mark.parametrize("stateCount,sizes,tips,cats", [
    (4, (37, 333, 90, 51), 40, 4),        # 4 unequal partitions, windows not aligned to the warp's 32 patterns
    (4, (700, 64, 1, 129), 90, 2),        # a one-pattern partition; > 24 ops so several subtrees run concurrently
    (20, (45, 130, 70), 16, 2),
    (61, (40, 72, 33), 9, 1),
])
@pytest.mark.parametrize("scheme", [S_.NONE, S_.ALWAYS])
def test_by_partition_route_matches_oracle(stateCount, sizes, tips, cats, scheme):
    tree, parts, models, sites = _partition_case(stateCount, sizes, tips, cats, seed=101 + stateCount)
    g = MultiPartitionDataLikelihoodDelegate(tree, parts, models, sites, GPU, resourceList=[1, 0], rescalingScheme=scheme,
                                             delayRescalingUntilUnderflow=False)
    o = MultiPartitionDataLikelihoodDelegate(tree, parts, models, sites, ORACLE, rescalingScheme=scheme,
                                             delayRescalingUntilUnderflow=False)
    lg, lo = tdl.TreeDataLikelihood(g, tree), tdl.TreeDataLikelihood(o, tree)
    for evaluation in range(2):            # the second one rescales under ALWAYS (MPDLD:746-790)
        vg, vo = lg.getLogLikelihood(), lo.getLogLikelihood()
        assert math.isfinite(vo) and _rel(vg, vo) <= REL, (evaluation, vg, vo)
        for k in range(len(parts)):
            assert _rel(g.cachedLogLikelihoodsByPartition[k], o.cachedLogLikelihoodsByPartition[k]) <= REL, k
        assert np.allclose(g.getSiteLogLikelihoods(), o.getSiteLogLikelihoods(), rtol=1e-10, atol=1e-12)
        for node in (tree.tipCount, tree.tipCount + (tree.nodeCount - tree.tipCount) // 2, tree.root):
            for part in (0, len(parts) - 1):
                pg, po = g.getPartials(part, node), o.getPartials(part, node)
                C, P, S = g.categoryCount, g.totalPatternCount, stateCount
                a, b = int(np.sum(g.patternCounts[:part])), int(np.sum(g.patternCounts[:part + 1]))
                pg, po = pg.reshape(C, P, S)[:, a:b], po.reshape(C, P, S)[:, a:b]
                assert np.allclose(pg, po, rtol=1e-9, atol=1e-13 * po.max()), (node, part)
        if scheme == S_.ALWAYS and evaluation == 1:
            assert all(g.useScaleFactors)
            cg, co = np.zeros(g.totalPatternCount), np.zeros(g.totalPatternCount)
            idx = g.scaleBufferHelper[0].getOffsetIndex(g.internalNodeCount)
            g.beagle.getLogScaleFactors(idx, cg)
            o.beagle.getLogScaleFactors(o.scaleBufferHelper[0].getOffsetIndex(o.internalNodeCount), co)
            a, b = 0, g.patternCounts[0]
            assert np.allclose(cg[a:b], co[a:b], rtol=1e-11, atol=1e-12)     # a cumulative buffer: logs, no second log()
        lg.makeDirty()
        lo.makeDirty()
    g.finalize()


def test_by_partition_mcmc_walk_store_restore_and_partial_updates():
    """25 steps: node-height moves (all partitions), site-model moves (ONE partition in the ByPartition calls), accept or
    reject; every value equals the oracle driven by the identical sequence, and a from-scratch instance at the end."""
    tree, parts, models, sites = _partition_case(4, (150, 420, 33), 50, 4, seed=77)
    tg, to = tree.copy(), tree.copy()
    sg, so = list(sites), list(sites)
    g = MultiPartitionDataLikelihoodDelegate(tg, parts, models, sg, GPU, resourceList=[1, 0], rescalingScheme=S_.DYNAMIC)
    o = MultiPartitionDataLikelihoodDelegate(to, parts, models, so, ORACLE, rescalingScheme=S_.DYNAMIC)
    lg, lo = tdl.TreeDataLikelihood(g, tg), tdl.TreeDataLikelihood(o, to)
    assert _rel(lg.getLogLikelihood(), lo.getLogLikelihood()) <= REL
    rng = np.random.default_rng(4)
    for step in range(25):
        accept = rng.random() < 0.5
        if step % 3 == 2:
            k = int(rng.integers(0, len(parts)))
            newSite = em.GammaSiteRateModel(shape=float(rng.uniform(0.2, 2.0)), gammaCategoryCount=4)
            for like, d in ((lg, g), (lo, o)):
                like.storeState()
                old = d.siteRateModels[k]
                d.siteRateModels[k] = newSite
                d.updateSiteRateModels[k] = True
                like.updateAllNodes()
                like._last = like.getLogLikelihood()
                if not accept:
                    d.siteRateModels[k] = old
                    like.restoreState()
        else:
            node = int(rng.integers(tree.tipCount, tree.nodeCount - 1))
            lo_h = max(tg.height[c] for c in tg.child[node])
            new_h = lo_h + (tg.height[tg.parent[node]] - lo_h) * rng.uniform(0.05, 0.95)
            for like, t in ((lg, tg), (lo, to)):
                like.storeState()
                old = t.height[node]
                t.height[node] = new_h
                like.updateNodeAndChildren(node)
                like._last = like.getLogLikelihood()
                if not accept:
                    t.height[node] = old
                    like.restoreState()
        assert _rel(lg._last, lo._last) <= REL, step
        assert _rel(lg.getLogLikelihood(), lo.getLogLikelihood()) <= REL, step
    fresh = MultiPartitionDataLikelihoodDelegate(tg, parts, models, g.siteRateModels, ORACLE, rescalingScheme=S_.NONE)
    assert _rel(lg.getLogLikelihood(), tdl.TreeDataLikelihood(fresh, tg).getLogLikelihood()) <= REL
    g.finalize()


def _post_and_pre_lists(tree, N, nodeCount, partitions):
    """post-order 9-tuples and the matching pre-order 9-tuples (pre buffers at nodeCount + node), all partitions."""
    like = tdl.TreeDataLikelihood.__new__(tdl.TreeDataLikelihood)
    like.tree, like.traversalType, like.updateNode = tree, "POST_ORDER", np.ones(nodeCount, dtype=bool)
    like._dispatch()
    post, pre = [], []
    for node, c1, c2 in like.nodeOperations:
        for k in range(partitions):
            post += [node, NONE, NONE, c1, c1 + k * nodeCount, c2, c2 + k * nodeCount, k, NONE]
    for node, c1, c2 in reversed(like.nodeOperations):            # parents before children
        for child, sib in ((c1, c2), (c2, c1)):
            for k in range(partitions):
                pre += [nodeCount + child, NONE, NONE, nodeCount + node, child + k * nodeCount, sib, sib + k * nodeCount, k, NONE]
    return like, post, pre


@pytest.mark.parametrize("stateCount,sizes,tips,cats", [(4, (70, 33, 129), 30, 4), (20, (40, 24), 10, 2)])
def test_update_pre_partials_by_partition_matches_oracle(stateCount, sizes, tips, cats):
    tree, parts, models, sites = _partition_case(stateCount, sizes, tips, cats, seed=55)
    N, nodeCount, K = tree.tipCount, tree.nodeCount, len(sizes)
    P, S, C = sum(sizes), stateCount, cats
    like, post, pre = _post_and_pre_lists(tree, N, nodeCount, K)
    out = []
    for factory, res in ((GPU, [1, 0]), (H.oracle_factory(), None)):
        b = factory(N, 2 * nodeCount, N, S, P, K, K * nodeCount, C, 4, res, 0, beagle.BeagleFlag.PREORDER_TRANSPOSE_AUTO if S > 4 else 0)
        for t in range(N):
            b.setTipStates(t, np.ascontiguousarray(np.concatenate([p.states[t] for p in parts]), dtype=np.int32))
        b.setPatternWeights(np.concatenate([p.weights for p in parts]))
        b.setPatternPartitions(K, np.concatenate([np.full(n, j, dtype=np.int32) for j, n in enumerate(sizes)]))
        eig, rate, prob, lens = [], [], [], []
        for k in range(K):
            e = models[k].getEigenDecomposition()
            b.setEigenDecomposition(k, e.Evec, e.Ievc, e.Eval)
            b.setCategoryRatesWithIndex(k, sites[k].getCategoryRates())
            b.setCategoryWeights(k, sites[k].getCategoryProportions())
            b.setStateFrequencies(k, models[k].getFrequencies())
            for node, t in like.branchOperations:
                eig.append(k); rate.append(k); prob.append(node + k * nodeCount); lens.append(t)
        b.updateTransitionMatricesWithMultipleModels(_i(eig), _i(rate), _i(prob), None, None, np.array(lens), len(eig))
        b.updatePartialsByPartition(_i(post), len(post) // 9)
        # root pre-order partial = frequencies of the pattern's partition (AbstractBeagleGradientDelegate.java:139-149)
        rootPre = np.zeros((C, P, S))
        a = 0
        for k, n in enumerate(sizes):
            rootPre[:, a:a + n, :] = models[k].getFrequencies()[None, None, :]
            a += n
        b.setPartials(nodeCount + tree.root, rootPre.reshape(-1))
        b.updatePrePartialsByPartition(_i(pre), len(pre) // 9)
        got = {}
        for node in (0, N - 1, N + 1, nodeCount - 2):
            x = np.zeros(C * P * S)
            b.getPartials(nodeCount + node, NONE, x)
            got[node] = x.reshape(C, P, S)
        out.append(got)
        b.finalize()
    for node in out[0]:
        assert np.allclose(out[0][node], out[1][node], rtol=1e-9, atol=1e-13 * out[1][node].max()), node


# ---- scale-factor entry points ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("log_scalers", [False, True])
def test_scale_factor_calls_match_oracle(log_scalers):
    """remove / copy / get(Log)ScaleFactors (+ByPartition) against the oracle after a rescaled evaluation."""
    tree, pats, model, site = H.synthetic_case(40, 300, 4, seed=3)
    flag = beagle.BeagleFlag.SCALERS_LOG if log_scalers else 0
    P = pats.patternCount
    res = []
    for factory, rl, extra in ((GPU, [1, 0], {"preferenceFlags": flag}), (H.oracle_factory(extra_flags=flag, report_flags=0), None, {})):
        d = tdl.BeagleDataLikelihoodDelegate(tree, pats, model, site, factory, resourceList=rl, rescalingScheme=S_.ALWAYS,
                                             delayRescalingUntilUnderflow=False, **extra)
        tdl.TreeDataLikelihood(d, tree).getLogLikelihood()
        b = d.beagle
        b.setPatternPartitions(2, _i([0] * 100 + [1] * (P - 100)))
        idx = list(d.scaleBufferIndices)
        cum = d.scaleBufferHelper.getOffsetIndex(d.internalNodeCount)
        spare = [k for k in range(d.scaleBufferHelper.getBufferCount()) if k not in idx and k != cum][:2]
        got = {}

        def grab(name, index, logs):
            x = np.zeros(P)
            (b.getLogScaleFactors if logs else b.getScaleFactors)(index, x)
            got[name] = x.copy()

        grab("node_raw", idx[3], False)
        grab("node_log", idx[3], True)
        grab("cum_log", cum, True)
        b.removeScaleFactors(_i(idx[:7]), 7, cum)
        grab("cum_after_remove", cum, True)
        b.copyScaleFactors(spare[0], cum)
        grab("copy_of_cum", spare[0], True)
        b.copyScaleFactors(spare[1], idx[5])
        grab("copy_of_node", spare[1], True)
        b.removeScaleFactorsByPartition(_i(idx[7:12]), 5, cum, 1)
        grab("cum_after_remove_part1", cum, True)
        b.resetScaleFactorsByPartition(cum, 0)
        b.accumulateScaleFactorsByPartition(_i(idx), len(idx), cum, 0)
        grab("cum_part0_reaccumulated", cum, True)
        res.append(got)
        d.finalize()
    for name in res[0]:
        assert np.allclose(res[0][name], res[1][name], rtol=1e-11, atol=1e-12), name
    assert np.allclose(res[0]["cum_after_remove_part1"][:100], res[0]["cum_after_remove"][:100])


# ---- in-list cumulative scaling -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("tips,patterns,cats,states", [(96, 500, 4, 4), (200, 97, 1, 4), (70, 80, 2, 20), (40, 48, 1, 61)])
@pytest.mark.parametrize("log_scalers", [False, True])
def test_update_partials_with_cumulative_scale_index(tips, patterns, cats, states, log_scalers):
    """beagleUpdatePartials(..., cumulativeScaleIndex != NONE): every op adds log(max) of its patterns to the cumulative
    buffer.  The lists have many concurrent subtrees (n > 24 ops), the case where an in-kernel '+=' would race."""
    tree, pats, model, site = H.synthetic_case(tips, patterns, cats, seed=tips + 1, stateCount=states)
    N, nodeCount, P, S, C = tree.tipCount, tree.nodeCount, pats.patternCount, states, cats
    flag = beagle.BeagleFlag.SCALERS_LOG if log_scalers else 0
    like = tdl.TreeDataLikelihood.__new__(tdl.TreeDataLikelihood)
    like.tree, like.traversalType, like.updateNode = tree, "REVERSE_LEVEL_ORDER", np.ones(nodeCount, dtype=bool)
    like._dispatch()
    ops = []
    for node, c1, c2 in like.nodeOperations:
        ops += [node, node - N, NONE, c1, c1, c2, c2]
    cum = nodeCount - N
    vals = []
    for factory, res in ((GPU, [1, 0]), (H.oracle_factory(extra_flags=flag), None)):
        b = factory(N, nodeCount, N, S, P, 1, nodeCount, C, cum + 1, res, flag, 0)
        for t in range(N):
            b.setTipStates(t, np.ascontiguousarray(pats.states[t], dtype=np.int32))
        b.setPatternWeights(pats.weights)
        e = model.getEigenDecomposition()
        b.setEigenDecomposition(0, e.Evec, e.Ievc, e.Eval)
        b.setCategoryRates(site.getCategoryRates())
        b.setCategoryWeights(0, site.getCategoryProportions())
        b.setStateFrequencies(0, model.getFrequencies())
        b.updateTransitionMatrices(0, _i([n for n, _ in like.branchOperations]), None, None,
                                   np.array([t for _, t in like.branchOperations]), len(like.branchOperations))
        out = np.zeros(1)
        for rep in range(3):                 # repeats go through the plan cache
            b.resetScaleFactors(cum)
            b.updatePartials(_i(ops), len(ops) // 7, cum)
            b.calculateRootLogLikelihoods(_i([tree.root]), _i([0]), _i([0]), _i([cum]), 1, out)
        c = np.zeros(P)
        b.getLogScaleFactors(cum, c)
        vals.append((out[0], c.copy()))
        b.finalize()
    assert math.isfinite(vals[1][0]) and _rel(vals[0][0], vals[1][0]) <= REL
    assert np.allclose(vals[0][1], vals[1][1], rtol=1e-11, atol=1e-11)


def test_by_partition_tuple_cumulative_field():
    """field 9 of the ByPartition tuple: per-partition cumulative buffers filled by the list itself."""
    tree, parts, models, sites = _partition_case(4, (90, 40, 200), 64, 4, seed=12)
    N, nodeCount, K = tree.tipCount, tree.nodeCount, 3
    P = sum(p.patternCount for p in parts)
    like, _, _ = _post_and_pre_lists(tree, N, nodeCount, K)
    nScale = (nodeCount - N) + K
    ops = []
    for node, c1, c2 in like.nodeOperations:
        for k in range(K):
            ops += [node, node - N, NONE, c1, c1 + k * nodeCount, c2, c2 + k * nodeCount, k, (nodeCount - N) + k]
    res = []
    for factory, rl in ((GPU, [1, 0]), (H.oracle_factory(), None)):
        b = factory(N, nodeCount, N, 4, P, K, K * nodeCount, 4, nScale, rl, 0, 0)
        for t in range(N):
            b.setTipStates(t, np.ascontiguousarray(np.concatenate([p.states[t] for p in parts]), dtype=np.int32))
        b.setPatternWeights(np.concatenate([p.weights for p in parts]))
        b.setPatternPartitions(K, np.concatenate([np.full(p.patternCount, j, dtype=np.int32) for j, p in enumerate(parts)]))
        eig, rate, prob, lens = [], [], [], []
        for k in range(K):
            e = models[k].getEigenDecomposition()
            b.setEigenDecomposition(k, e.Evec, e.Ievc, e.Eval)
            b.setCategoryRatesWithIndex(k, sites[k].getCategoryRates())
            b.setCategoryWeights(k, sites[k].getCategoryProportions())
            b.setStateFrequencies(k, models[k].getFrequencies())
            for node, t in like.branchOperations:
                eig.append(k); rate.append(k); prob.append(node + k * nodeCount); lens.append(t)
        b.updateTransitionMatricesWithMultipleModels(_i(eig), _i(rate), _i(prob), None, None, np.array(lens), len(eig))
        for k in range(K):
            b.resetScaleFactorsByPartition((nodeCount - N) + k, k)
        b.updatePartialsByPartition(_i(ops), len(ops) // 9)
        byPart, total = np.zeros(K), np.zeros(1)
        b.calculateRootLogLikelihoodsByPartition(_i([tree.root] * K), _i(range(K)), _i(range(K)),
                                                 _i([(nodeCount - N) + k for k in range(K)]), _i(range(K)), K, 1, byPart, total)
        res.append((byPart.copy(), total[0]))
        b.finalize()
    assert _rel(res[0][1], res[1][1]) <= REL
    assert all(_rel(res[0][0][k], res[1][0][k]) <= REL for k in range(K))


# ---- BeagleTreeLikelihood re-enactment: setTipPartials, in-call underflow retry, ascertainment ---------------------------
def test_beagle_tree_likelihood_route_matches_oracle():
    tree, pats, model, site = H.synthetic_case(48, 333, 4, seed=8)
    rng = np.random.default_rng(8)
    noisy = []
    for t in range(tree.tipCount):
        p = np.eye(4)[np.minimum(pats.states[t], 3)] * 0.96 + 0.01
        p[rng.random(pats.patternCount) < 0.05] = 1.0             # some missing data
        noisy.append(p)
    likes = []
    for factory, res in ((GPU, [1, 0]), (ORACLE, None)):
        tp = tree.copy()
        likes.append((BeagleTreeLikelihood(pats, tp, model, site, factory, tipStatesModel=TipPartialsModel(noisy),
                                           resourceList=res, rescalingScheme=S_.DYNAMIC,
                                           ascertainedExclude=[0, 5, 17]), tp))
    (lg, tg), (lo, to) = likes
    assert _rel(lg.getLogLikelihood(), lo.getLogLikelihood()) <= REL
    rng = np.random.default_rng(1)
    for step in range(12):
        node = int(rng.integers(tree.tipCount, tree.nodeCount - 1))
        lo_h = max(tg.height[c] for c in tg.child[node])
        new_h = lo_h + (tg.height[tg.parent[node]] - lo_h) * rng.uniform(0.05, 0.95)
        accept = rng.random() < 0.5
        for like, t in ((lg, tg), (lo, to)):
            like.storeState()
            old = t.height[node]
            t.height[node] = new_h
            like.updateNodeAndChildren(node)
            like._last = like.getLogLikelihood()
            if not accept:
                t.height[node] = old
                like.restoreState()
        assert _rel(lg._last, lo._last) <= REL, step
    assert np.allclose(lg.getPartials(tree.root), lo.getPartials(tree.root), rtol=1e-9, atol=1e-300)
    lg.finalize()


def test_beagle_tree_likelihood_underflow_retry_inside_the_call():
    tree, pats, model, site = H.synthetic_case(700, 48, 4, seed=9, rootHeight=3000.0)
    vals = []
    for factory, res in ((GPU, [1, 0]), (ORACLE, None)):
        like = BeagleTreeLikelihood(pats, tree.copy(), model, site, factory, resourceList=res, rescalingScheme=S_.DYNAMIC)
        vals.append(like.getLogLikelihood())
        assert like.everUnderflowed and like.useScaleFactors
        like.finalize()
    assert math.isfinite(vals[1]) and _rel(vals[0], vals[1]) <= REL


# ---- BASELINE configs 3 and 4 at full size against the C port ------------------------------------------------------------
@pytest.mark.parametrize("name", ["codon_mg94_500x5k", "makona_like_1610x6k"])
def test_full_size_configs_match_c_port(name):
    import bench
    from oracle import cpu
    from beast_mcmc_b200 import build
    build.build_oracle()
    w, tree, pats, model, site = bench.build_workload(name, 0, {})
    S, C, P = w["states"], site.getCategoryCount(), pats.patternCount
    scheme = S_.ALWAYS if name.startswith("makona") else S_.NONE        # cfg 4 runs rescaled, as BEAST would
    nodes = [tree.tipCount, tree.tipCount + 7, tree.tipCount + (tree.nodeCount - tree.tipCount) // 2, tree.root]
    got = []
    for factory, res in ((GPU, [1, 0]), (cpu.factory(threads=16, reportFlags=0), None)):
        d = tdl.BeagleDataLikelihoodDelegate(tree, pats, model, site, factory, resourceList=res, rescalingScheme=scheme,
                                             delayRescalingUntilUnderflow=False)
        v = tdl.TreeDataLikelihood(d, tree).getLogLikelihood()
        got.append((v, d.getSiteLogLikelihoods(), [d.getPartials(n).reshape(C, P, S) for n in nodes]))
        d.finalize()
    (vg, sg, pg), (vc, sc, pc) = got
    assert math.isfinite(vc) and _rel(vg, vc) <= REL, (vg, vc)
    # Per-site values: a pattern whose likelihood hinges on ONE substitution along a branch of length t carries the
    # cancellation error of P_ij(t) = sum_k V_ik e^(lambda_k t) V^-1_kj ~ 1e-16 / (q t) in EVERY implementation of the
    # reference's formula (BaseSubstitutionModel.java:206-241); the epidemic-scale tree has branches down to ~1e-7, so
    # single sites agree to ~1e-9, not 1e-10 -- the bar of north_star (1e-10) is on the weighted root sum, asserted above.
    err = np.abs(sg - sc) / np.maximum(1.0, np.abs(sc))
    print(f"{name}: root rel.err {_rel(vg, vc):.2e}; site log-likelihoods max rel.err {err.max():.2e}, "
          f"99.9th percentile {np.quantile(err, 0.999):.2e}")
    assert err.max() <= 5e-9 and np.quantile(err, 0.99) <= 1e-10, (err.max(), np.quantile(err, 0.99))
    # rescaled partials inherit the same effect (a cherry with t ~ 1e-9 whose tips differ has ALL its entries ~ q t before
    # the division by their maximum): 1e-6 there, 1e-8 on the codon tree with ordinary branch lengths
    rtol = 1e-6 if name.startswith("makona") else 1e-8
    for a, b, n in zip(pg, pc, nodes):
        assert np.allclose(a, b, rtol=rtol, atol=1e-13 * b.max()), (n, float(np.max(np.abs(a - b) / (np.abs(b) + 1e-13 * b.max()))))


# ---- two devices from two threads -----------------------------------------------------------------------------------------
def test_two_instances_on_two_resources_from_two_threads():
    """What a JVM does with -beagle_instances 2 -beagle_order 1,2 (BDLD:275-281, CompoundLikelihood.java:63-82): one
    instance per device, one pool thread each, concurrently."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    tree, pats, model, site = H.synthetic_case(60, 2001, 4, seed=23)
    shards = [pats.subSet(k, 2) for k in range(2)]
    serial = []
    for k in range(2):
        d = tdl.BeagleDataLikelihoodDelegate(tree, shards[k], model, site, GPU, resourceList=[1, 0], rescalingScheme=S_.NONE)
        serial.append(tdl.TreeDataLikelihood(d, tree).getLogLikelihood())
        d.finalize()
    results = [[], []]
    details = [None, None]

    def work(k):
        d = tdl.BeagleDataLikelihoodDelegate(tree, shards[k], model, site, GPU, resourceList=[k + 1, 0],
                                             rescalingScheme=S_.NONE)
        details[k] = d.beagle.getDetails().getResourceNumber()
        like = tdl.TreeDataLikelihood(d, tree)
        for _ in range(20):
            like.makeDirty()
            results[k].append(like.getLogLikelihood())
        d.finalize()

    threads = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert details == [1, 2]
    whole = tdl.BeagleDataLikelihoodDelegate(tree, pats, model, site, ORACLE, rescalingScheme=S_.NONE)
    lw = tdl.TreeDataLikelihood(whole, tree).getLogLikelihood()
    for k in range(2):
        assert len(results[k]) == 20 and all(v == serial[k] for v in results[k])
    assert _rel(results[0][0] + results[1][0], lw) <= REL


# ---- epoch-model matrices (convolved on the device) inside an operation list ------------------------------------------------
@pytest.mark.parametrize("states,cats", [(4, 4), (20, 2)])
def test_convolved_matrices_in_an_operation_list(states, cats):
    """SubstitutionModelDelegate.java:303-470: P(t1) x P(t2) per branch through convolveTransitionMatrices, then the usual
    list.  For S = 4 such matrices carry no spectrum, so the list runs on the matrix-form walk; the value equals the oracle's
    and (Chapman-Kolmogorov) the plain evaluation with t1 + t2."""
    tree, pats, model, site = H.synthetic_case(24, 257, cats, seed=61, stateCount=states)
    N, n, P, S, C = tree.tipCount, tree.nodeCount, pats.patternCount, states, cats
    branches, nodeOps = [], []
    like = tdl.TreeDataLikelihood.__new__(tdl.TreeDataLikelihood)
    like.tree, like.traversalType, like.updateNode = tree, "REVERSE_LEVEL_ORDER", np.ones(n, dtype=bool)
    like._dispatch()
    nodes = [b for b, _ in like.branchOperations]
    lens = np.array([t for _, t in like.branchOperations])
    ops = []
    for node, c1, c2 in like.nodeOperations:
        ops += [node, NONE, NONE, c1, 2 * n + c1, c2, 2 * n + c2]          # the convolved matrices live at 2n + branch
    vals = []
    for factory, rl in ((GPU, [1, 0]), (H.oracle_factory(), None)):
        b = factory(N, n, N, S, P, 1, 3 * n, C, 0, rl, 0, 0)
        for t in range(N):
            b.setTipStates(t, np.ascontiguousarray(pats.states[t], dtype=np.int32))
        b.setPatternWeights(pats.weights)
        e = model.getEigenDecomposition()
        b.setEigenDecomposition(0, e.Evec, e.Ievc, e.Eval)
        b.setCategoryRates(site.getCategoryRates())
        b.setCategoryWeights(0, site.getCategoryProportions())
        b.setStateFrequencies(0, model.getFrequencies())
        b.updateTransitionMatrices(0, _i(nodes), None, None, 0.3 * lens, len(nodes))
        b.updateTransitionMatrices(0, _i([n + k for k in nodes]), None, None, 0.7 * lens, len(nodes))
        b.convolveTransitionMatrices(_i(nodes), _i([n + k for k in nodes]), _i([2 * n + k for k in nodes]), len(nodes))
        out = np.zeros(1)
        b.updatePartials(_i(ops), len(ops) // 7, NONE)
        b.calculateRootLogLikelihoods(_i([tree.root]), _i([0]), _i([0]), _i([NONE]), 1, out)
        vals.append(out[0])
        # the same tree with the whole branch in one matrix
        b.updateTransitionMatrices(0, _i([2 * n + k for k in nodes]), None, None, lens, len(nodes))
        b.updatePartials(_i(ops), len(ops) // 7, NONE)
        b.calculateRootLogLikelihoods(_i([tree.root]), _i([0]), _i([0]), _i([NONE]), 1, out)
        vals.append(out[0])
        b.finalize()
    assert math.isfinite(vals[2]) and _rel(vals[0], vals[2]) <= REL, vals
    assert _rel(vals[1], vals[3]) <= REL and _rel(vals[0], vals[1]) <= 1e-9, vals


# ---- deferred small evaluations: matrices + list + root as one launch -------------------------------------------------------
@pytest.mark.parametrize("cats,scheme", [(4, S_.NONE), (1, S_.NONE), (4, S_.DYNAMIC), (2, S_.ALWAYS)])
def test_fused_incremental_evaluations_match_oracle(cats, scheme):
    """An MCMC-like walk of node-height moves with accept / reject: every incremental evaluation (3 branches, one root path)
    goes through the one-launch route (csrc/incr.cu) and equals the oracle driven by the identical call sequence; the
    same walk with B200_FUSE=0 gives the identical doubles (same arithmetic in both kernels for the per-cell work)."""
    import os
    rootHeight = 3000.0 if scheme == S_.DYNAMIC else 0.1                  # DYNAMIC: underflow -> scale buffers read by the ops
    tips = 700 if scheme == S_.DYNAMIC else 80
    tree, pats, model, site = H.synthetic_case(tips, 64 if scheme == S_.DYNAMIC else 501, cats, seed=29, rootHeight=rootHeight)

    def walk(factory, res, fuse):
        os.environ["B200_FUSE"] = fuse
        try:
            t = tree.copy()
            d = tdl.BeagleDataLikelihoodDelegate(t, pats, model, site, factory, resourceList=res, rescalingScheme=scheme,
                                                 delayRescalingUntilUnderflow=scheme == S_.DYNAMIC)
        finally:
            os.environ.pop("B200_FUSE", None)
        like = tdl.TreeDataLikelihood(d, t)
        vals = [like.getLogLikelihood()]
        rng = np.random.default_rng(7)
        for step in range(30):
            node = int(rng.integers(t.tipCount, t.nodeCount - 1))
            lo_h = max(t.height[c] for c in t.child[node])
            new_h = lo_h + (t.height[t.parent[node]] - lo_h) * rng.uniform(0.05, 0.95)
            like.storeState()
            old = t.height[node]
            t.height[node] = new_h
            like.updateNodeAndChildren(node)
            vals.append(like.getLogLikelihood())
            if rng.random() < 0.5:
                t.height[node] = old
                like.restoreState()
            vals.append(like.getLogLikelihood())
        sites = d.getSiteLogLikelihoods()
        fused = beagle.load_library().b200GetFusedLaunches(d.beagle.instance) if factory is GPU else 0
        d.finalize()
        return vals, sites, fused

    vg, sg, fused = walk(GPU, [1, 0], "1")
    vp, sp, plain = walk(GPU, [1, 0], "0")
    vo, so, _ = walk(ORACLE, None, "1")
    # ALWAYS resets and accumulates the scale factors between the list and the root call: such evaluations keep the ordinary
    # launches; NONE and DYNAMIC (scale factors only READ by the ops, BEAST's default) fuse
    assert plain == 0 and (fused == 0 if scheme == S_.ALWAYS else fused >= 25), (fused, plain)
    assert all(math.isfinite(v) for v in vo)
    assert all(_rel(a, b) <= REL for a, b in zip(vg, vo)), max(_rel(a, b) for a, b in zip(vg, vo))
    assert all(_rel(a, b) <= 1e-13 for a, b in zip(vg, vp))
    assert np.allclose(sg, so, rtol=1e-10, atol=1e-11)
