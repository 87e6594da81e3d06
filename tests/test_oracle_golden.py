"""Pin the oracle (and the caller re-enactment) to the reference's own golden vectors:
TreeDataLikelihoodTest.java:131-314 (ten values, 5 decimals), LikelihoodTest.java:106-341 (ten more, other parameters,
the older site-model rate rule) and the BEAGLE tiny test."""
import numpy as np
import pytest

import helpers as H
from harness import treedatalikelihood as tdl


def _fmt(x):
    return f"{x:.5f}"


@pytest.mark.parametrize("name", list(H.primate_cases().keys()))
@pytest.mark.parametrize("traversal_flags", [tdl.FLAG_FRAMEWORK_CPU, 0])
def test_primates_through_delegate(name, traversal_flags):
    model, site, expected = H.primate_cases()[name]
    delegate = tdl.BeagleDataLikelihoodDelegate(
        H.primate_tree(), H.primate_patterns(), model, site,
        H.oracle_factory(report_flags=traversal_flags),
        useAmbiguities=False, rescalingScheme=tdl.PartialsRescalingScheme.DEFAULT,
        delayRescalingUntilUnderflow=False)     # exactly the ctor arguments of the JUnit test
    like = tdl.TreeDataLikelihood(delegate, H.primate_tree())
    assert _fmt(like.getLogLikelihood()) == _fmt(expected)
    # the DYNAMIC scheme with delay=false rescales on the very first evaluation
    assert delegate.useScaleFactors


@pytest.mark.parametrize("name", list(H.primate_cases_legacy().keys()))
def test_primates_legacy_likelihood_test_values(name):
    """LikelihoodTest.java:106-341: K80 kappa=27.402591, HKY85+G kappa=38.829740 alpha=0.137064, HKY85+I pInv=0.701211, ..."""
    model, site, expected = H.primate_cases_legacy()[name]
    d = tdl.BeagleDataLikelihoodDelegate(H.primate_tree(), H.primate_patterns(), model, site, H.oracle_factory(),
                                         rescalingScheme=tdl.PartialsRescalingScheme.NONE)
    assert _fmt(tdl.TreeDataLikelihood(d, H.primate_tree()).getLogLikelihood()) == _fmt(expected)


@pytest.mark.parametrize("k", [0, 1, 2])
def test_msat_hand_calculated_values_to_1e10(k):
    """MsatFullLikelihoodTest.java:181-189: the only values the reference pins to 1e-10 (3- and 4-state stepwise models)."""
    tree, pats, model, site, expected = H.msat_cases()[k]
    for scheme in (tdl.PartialsRescalingScheme.NONE, tdl.PartialsRescalingScheme.ALWAYS):
        d = tdl.BeagleDataLikelihoodDelegate(tree, pats, model, site, H.oracle_factory(), rescalingScheme=scheme,
                                             delayRescalingUntilUnderflow=False)
        assert abs(tdl.TreeDataLikelihood(d, tree).getLogLikelihood() - expected) <= 1e-10


def test_primates_unscaled_equals_scaled():
    model, site, expected = H.primate_cases()["GTRGI"]
    vals = []
    for scheme, delay in [(tdl.PartialsRescalingScheme.NONE, True), (tdl.PartialsRescalingScheme.ALWAYS, False),
                          (tdl.PartialsRescalingScheme.DYNAMIC, False)]:
        for log_flag in (0, 1 << 10):
            d = tdl.BeagleDataLikelihoodDelegate(H.primate_tree(), H.primate_patterns(), model, site,
                                                 H.oracle_factory(extra_flags=log_flag), rescalingScheme=scheme,
                                                 delayRescalingUntilUnderflow=delay)
            vals.append(tdl.TreeDataLikelihood(d, H.primate_tree()).getLogLikelihood())
    assert np.allclose(vals, vals[0], rtol=1e-13, atol=0)
    assert _fmt(vals[0]) == _fmt(expected)


def test_primates_ambiguities_as_partials():
    from harness.evomodel import nucleotide_state_set
    model, site, expected = H.primate_cases()["HKY85G"]
    d = tdl.BeagleDataLikelihoodDelegate(H.primate_tree(), H.primate_patterns(), model, site, H.oracle_factory(),
                                         useAmbiguities=True, stateSetFn=nucleotide_state_set)
    assert _fmt(tdl.TreeDataLikelihood(d, H.primate_tree()).getLogLikelihood()) == _fmt(expected)


def test_beagle_tiny_test():
    tree, pats, model, site, expected = H.tiny_case()
    d = tdl.BeagleDataLikelihoodDelegate(tree, pats, model, site, H.oracle_factory(),
                                         rescalingScheme=tdl.PartialsRescalingScheme.NONE)
    assert _fmt(tdl.TreeDataLikelihood(d, tree).getLogLikelihood()) == _fmt(expected)


def test_jc69_eigenvalues_match_beagle_tiny_constants():
    # BeagleFactory.main hard-codes Eval = {0, -4/3, -4/3, -4/3} for JC69
    _, _, model, _, _ = H.tiny_case()
    lam = np.sort(model.getEigenDecomposition().Eval)
    assert np.allclose(lam, [-4 / 3, -4 / 3, -4 / 3, 0.0], atol=1e-14)


# ---- the C restatement (oracle/beagle_cpu.c: checker + cpu_baseline) is pinned the same way ------
@pytest.fixture(scope="module")
def cport():
    from beast_mcmc_b200 import build
    build.build_oracle()
    from oracle import cpu
    return cpu


@pytest.mark.parametrize("name", list(H.primate_cases().keys()))
def test_c_port_primates_golden(cport, name):
    model, site, expected = H.primate_cases()[name]
    d = tdl.BeagleDataLikelihoodDelegate(H.primate_tree(), H.primate_patterns(), model, site, cport.factory(threads=3),
                                         delayRescalingUntilUnderflow=False)
    assert _fmt(tdl.TreeDataLikelihood(d, H.primate_tree()).getLogLikelihood()) == _fmt(expected)
    d.finalize()


@pytest.mark.parametrize("name", list(H.primate_cases_legacy().keys()))
def test_c_port_primates_legacy_golden(cport, name):
    model, site, expected = H.primate_cases_legacy()[name]
    d = tdl.BeagleDataLikelihoodDelegate(H.primate_tree(), H.primate_patterns(), model, site, cport.factory(threads=2),
                                         delayRescalingUntilUnderflow=False)
    assert _fmt(tdl.TreeDataLikelihood(d, H.primate_tree()).getLogLikelihood()) == _fmt(expected)
    d.finalize()


@pytest.mark.parametrize("k", [0, 1, 2])
def test_c_port_msat_hand_calculated_values(cport, k):
    tree, pats, model, site, expected = H.msat_cases()[k]
    d = tdl.BeagleDataLikelihoodDelegate(tree, pats, model, site, cport.factory(threads=1))
    assert abs(tdl.TreeDataLikelihood(d, tree).getLogLikelihood() - expected) <= 1e-10
    d.finalize()


@pytest.mark.parametrize("states,cats,scheme", [(4, 4, "none"), (4, 5, "always"), (20, 2, "always"), (61, 1, "none")])
def test_c_port_matches_numpy_oracle(cport, states, cats, scheme):
    tree, pats, model, site = H.synthetic_case(24, 150, cats, seed=states + cats, stateCount=states)
    vals = []
    for f in (cport.factory(threads=4), H.oracle_factory()):
        d = tdl.BeagleDataLikelihoodDelegate(tree, pats, model, site, f, rescalingScheme=scheme,
                                             delayRescalingUntilUnderflow=False)
        vals.append(tdl.TreeDataLikelihood(d, tree).getLogLikelihood())
        sites = d.getSiteLogLikelihoods()
        vals.append(sites)
    assert abs(vals[0] - vals[2]) <= 1e-12 * abs(vals[2])
    assert np.allclose(vals[1], vals[3], rtol=1e-12, atol=1e-13)
