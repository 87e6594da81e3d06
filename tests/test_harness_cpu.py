"""CPU suite for the caller re-enactments added in round 2 (oracle only, no GPU): the multi-partition delegate and the
older BeagleTreeLikelihood front-end must reproduce, on the numpy oracle, what the single-partition delegate produces --
and the reference's golden values -- before they are trusted to drive the CUDA engine (tests/test_gpu_parity2.py)."""
import math

import numpy as np
import pytest

import helpers as H
from harness import evomodel as em, treedatalikelihood as tdl
from harness.beagletreelikelihood import BeagleTreeLikelihood, TipPartialsModel
from harness.multipartition import MultiPartitionDataLikelihoodDelegate

S_ = tdl.PartialsRescalingScheme
ORACLE = H.oracle_factory(report_flags=0)


@pytest.mark.parametrize("name", ["JC69", "HKY85G", "GTRGI"])
def test_beagle_tree_likelihood_reproduces_reference_values(name):
    """Same alignment / tree / models as TreeDataLikelihoodTest.java through the BeagleTreeLikelihood call sequence."""
    model, site, expected = H.primate_cases()[name]
    for scheme in (S_.NONE, S_.ALWAYS):
        like = BeagleTreeLikelihood(H.primate_patterns(), H.primate_tree(), model, site, ORACLE, rescalingScheme=scheme,
                                    delayRescalingUntilUnderflow=False)
        assert f"{like.getLogLikelihood():.5f}" == f"{expected:.5f}"


def test_beagle_tree_likelihood_underflow_retry_and_store_restore():
    tree, pats, model, site = H.synthetic_case(700, 16, 2, seed=9, rootHeight=3000.0)
    like = BeagleTreeLikelihood(pats, tree.copy(), model, site, ORACLE, rescalingScheme=S_.DYNAMIC)
    ref = tdl.TreeDataLikelihood(tdl.BeagleDataLikelihoodDelegate(tree, pats, model, site, ORACLE, rescalingScheme=S_.ALWAYS,
                                                                  delayRescalingUntilUnderflow=False), tree)
    v = like.getLogLikelihood()
    assert like.everUnderflowed and math.isfinite(v)
    assert abs(v - ref.getLogLikelihood()) <= 1e-10 * abs(v)
    t = like.treeModel
    node = t.tipCount + 5
    like.storeState()
    old = t.height[node]
    t.height[node] = 0.5 * (max(t.height[c] for c in t.child[node]) + t.height[t.parent[node]])
    like.updateNodeAndChildren(node)
    moved = like.getLogLikelihood()
    t.height[node] = old
    like.restoreState()
    assert like.getLogLikelihood() == v and moved != v
    like.makeDirty()
    assert abs(like.getLogLikelihood() - v) <= 1e-10 * abs(v)


def test_beagle_tree_likelihood_tip_partials_model():
    """setTipPartials route (BeagleTreeLikelihood.java:917-930): one-hot partials reproduce the compact-state value; an
    error model (mass spread over the other states) changes it."""
    tree, pats, model, site = H.synthetic_case(12, 60, 3, seed=5)
    plain = BeagleTreeLikelihood(pats, tree, model, site, ORACLE, rescalingScheme=S_.NONE).getLogLikelihood()
    onehot = [np.eye(4)[pats.states[t]] for t in range(tree.tipCount)]
    a = BeagleTreeLikelihood(pats, tree, model, site, ORACLE, tipStatesModel=TipPartialsModel(onehot),
                             rescalingScheme=S_.NONE).getLogLikelihood()
    assert abs(a - plain) <= 1e-12 * abs(plain)
    noisy = [0.97 * p + 0.01 for p in onehot]
    b = BeagleTreeLikelihood(pats, tree, model, site, ORACLE, tipStatesModel=TipPartialsModel(noisy),
                             rescalingScheme=S_.NONE).getLogLikelihood()
    assert math.isfinite(b) and abs(b - plain) > 1e-6


def _partition_case(stateCount, sizes=(37, 90, 51), tips=14, cats=3, seed=31):
    tree, pats, model, site = H.synthetic_case(tips, sum(sizes), cats, seed=seed, stateCount=stateCount)
    cuts = np.cumsum((0,) + tuple(sizes))
    parts = [em.Patterns(np.ascontiguousarray(pats.states[:, a:b]), pats.weights[a:b].copy(), stateCount)
             for a, b in zip(cuts[:-1], cuts[1:])]
    rng = np.random.default_rng(seed)
    models, sites = [], []
    for k in range(len(sizes)):
        if stateCount == 4:
            models.append(em.HKY(1.5 + k, rng.dirichlet(np.full(4, 8.0))))
        else:
            models.append(em.SubstitutionModel(rng.uniform(0.2, 3.0, stateCount * (stateCount - 1) // 2),
                                               rng.dirichlet(np.full(stateCount, 5.0))))
        sites.append(em.GammaSiteRateModel(shape=0.4 + 0.3 * k, gammaCategoryCount=cats))
    return tree, parts, models, sites


@pytest.mark.parametrize("stateCount", [4, 7])
@pytest.mark.parametrize("scheme", [S_.NONE, S_.ALWAYS])
def test_multipartition_delegate_equals_one_delegate_per_partition(stateCount, scheme):
    tree, parts, models, sites = _partition_case(stateCount)
    mp = MultiPartitionDataLikelihoodDelegate(tree, parts, models, sites, ORACLE, rescalingScheme=scheme,
                                              delayRescalingUntilUnderflow=False)
    like = tdl.TreeDataLikelihood(mp, tree)
    total = like.getLogLikelihood()
    like.makeDirty()                       # second evaluation: ALWAYS rescales from here on (MPDLD:746-790)
    total2 = like.getLogLikelihood()
    separate = [tdl.TreeDataLikelihood(tdl.BeagleDataLikelihoodDelegate(tree, parts[k], models[k], sites[k], ORACLE,
                                                                         rescalingScheme=S_.NONE), tree).getLogLikelihood()
                for k in range(len(parts))]
    assert abs(total - sum(separate)) <= 1e-11 * abs(total)
    assert abs(total2 - sum(separate)) <= 1e-11 * abs(total)
    for k in range(len(parts)):
        assert abs(mp.cachedLogLikelihoodsByPartition[k] - separate[k]) <= 1e-11 * abs(separate[k])
    assert mp.useScaleFactors == [scheme == S_.ALWAYS] * len(parts)


def test_multipartition_store_restore_and_single_partition_update():
    tree, parts, models, sites = _partition_case(4, tips=20)
    t = tree.copy()
    mp = MultiPartitionDataLikelihoodDelegate(t, parts, models, sites, ORACLE, rescalingScheme=S_.NONE)
    like = tdl.TreeDataLikelihood(mp, t)
    v0 = like.getLogLikelihood()
    like.storeState()
    node = t.tipCount + 3
    old = t.height[node]
    t.height[node] = 0.5 * (max(t.height[c] for c in t.child[node]) + t.height[t.parent[node]])
    like.updateNodeAndChildren(node)
    v1 = like.getLogLikelihood()
    fresh = tdl.TreeDataLikelihood(MultiPartitionDataLikelihoodDelegate(t, parts, models, sites, ORACLE,
                                                                        rescalingScheme=S_.NONE), t).getLogLikelihood()
    assert abs(v1 - fresh) <= 1e-11 * abs(fresh) and v1 != v0
    t.height[node] = old
    like.restoreState()
    assert like.getLogLikelihood() == v0
    # only partition 1's site model changes: one partition in the ByPartition calls (MPDLD:826-846,1059-1071)
    mp.siteRateModels[1] = em.GammaSiteRateModel(shape=2.5, gammaCategoryCount=3)
    mp.updateSiteRateModels[1] = True
    like.updateAllNodes()
    v2 = like.getLogLikelihood()
    sites2 = list(sites)
    sites2[1] = mp.siteRateModels[1]
    fresh2 = tdl.TreeDataLikelihood(MultiPartitionDataLikelihoodDelegate(t, parts, models, sites2, ORACLE,
                                                                         rescalingScheme=S_.NONE), t).getLogLikelihood()
    assert abs(v2 - fresh2) <= 1e-11 * abs(fresh2) and v2 != v0
