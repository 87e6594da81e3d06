"""Site-pattern compression: the GPU hash-table build against a literal restatement of
SitePatterns.addPattern (src/dr/evolution/alignment/SitePatterns.java:356-372: scan the patterns found so far, merge on exact
equality, append otherwise) -- bit-exact patterns, first-occurrence order, weights, per-site indices."""
import numpy as np
import pytest

import helpers as H
from harness import evomodel as em


def java_add_patterns(states, siteWeights=None):
    """The reference's O(sites x patterns) loop, literally."""
    taxa, sites = states.shape
    patterns, weights, index = [], [], np.zeros(sites, dtype=np.int32)
    for s in range(sites):
        col = states[:, s]
        w = 1.0 if siteWeights is None else float(siteWeights[s])
        for i, p in enumerate(patterns):
            if np.array_equal(p, col):
                weights[i] += w
                index[s] = i
                break
        else:
            index[s] = len(patterns)
            patterns.append(col.copy())
            weights.append(w)
    P = len(patterns)
    return (np.stack(patterns, axis=1) if P else np.zeros((taxa, 0), dtype=np.int32)), np.asarray(weights), index


def test_numpy_mirror_equals_java_loop():
    rng = np.random.default_rng(3)
    states = rng.integers(0, 3, size=(5, 400)).astype(np.int32)
    pats, w, idx = java_add_patterns(states)
    m = em.Patterns.fromAlignment(states)
    assert np.array_equal(m.states, pats) and np.array_equal(m.weights, w)
    assert pats.shape[1] < 400                      # the case really has duplicates
    assert np.array_equal(pats[:, idx], states)


def primate_alignment():
    return em.encode_nucleotides(H.GOLDEN["primates"]["sequences"]).astype(np.int32)


def test_reference_pattern_counts_on_the_primate_alignment():
    """SitePatternsTest.java:75-99: 768 sites, 69 unique patterns, 37 unique patterns among the third codon positions
    (from = 2, every = 3)."""
    aln = primate_alignment()
    assert aln.shape == (6, 768)
    for cols, expected in ((aln, 69), (aln[:, 2::3], 37)):
        pats, w, idx = java_add_patterns(np.ascontiguousarray(cols))
        m = em.Patterns.fromAlignment(cols)
        assert pats.shape[1] == expected == m.patternCount and w.sum() == cols.shape[1]
        assert np.array_equal(m.states, pats) and np.array_equal(m.weights, w)


@pytest.mark.gpu
def test_gpu_reference_pattern_counts_on_the_primate_alignment():
    from beast_mcmc_b200 import beagle
    aln = primate_alignment()
    for cols, expected in ((aln, 69), (np.ascontiguousarray(aln[:, 2::3]), 37)):
        pats, w, idx = beagle.compressSitePatterns(cols)
        epats, ew, eidx = java_add_patterns(cols)
        assert pats.shape[1] == expected
        assert np.array_equal(pats, epats) and np.array_equal(w, ew) and np.array_equal(idx, eidx)


CASES = [(5, 400, 3, 1), (1, 50, 2, 2), (40, 3000, 4, 3), (7, 1, 4, 4), (3, 2000, 18, 5), (300, 700, 2, 6)]


@pytest.mark.gpu
@pytest.mark.parametrize("taxa,sites,alphabet,seed", CASES)
def test_gpu_compression_is_bit_exact(taxa, sites, alphabet, seed):
    from beast_mcmc_b200 import beagle
    rng = np.random.default_rng(seed)
    states = rng.integers(0, alphabet, size=(taxa, sites)).astype(np.int32)
    if sites > 10:                                   # force long runs of duplicates and a late first occurrence
        states[:, sites // 2:sites // 2 + 5] = states[:, [0]]
        states[:, -1] = states[:, 3]
    pats, w, idx = beagle.compressSitePatterns(states)
    epats, ew, eidx = java_add_patterns(states)
    assert np.array_equal(pats, epats) and np.array_equal(w, ew) and np.array_equal(idx, eidx)
    sw = rng.uniform(0.5, 2.0, sites)
    pats2, w2, idx2 = beagle.compressSitePatterns(states, sw)
    _, ew2, _ = java_add_patterns(states, sw)
    assert np.array_equal(pats2, epats) and np.array_equal(idx2, eidx) and np.array_equal(w2, ew2)   # same order of additions


@pytest.mark.gpu
def test_gpu_compression_edge_cases_and_reference_alignment():
    from beast_mcmc_b200 import beagle
    pats, w, idx = beagle.compressSitePatterns(np.zeros((4, 0), dtype=np.int32))
    assert pats.shape == (4, 0) and w.size == 0 and idx.size == 0
    same = np.tile(np.array([[1], [2], [3]], dtype=np.int32), (1, 1000))
    pats, w, idx = beagle.compressSitePatterns(same)
    assert pats.shape == (3, 1) and w[0] == 1000.0 and not idx.any()
    uniq = np.arange(5000, dtype=np.int32)[None, :].repeat(2, axis=0)
    pats, w, idx = beagle.compressSitePatterns(uniq)
    assert np.array_equal(pats, uniq) and np.all(w == 1.0) and np.array_equal(idx, np.arange(5000))
    # the reference's benchmark2 alignment: expand its patterns back into sites, shuffle, compress again
    z = np.load(H.ROOT + "/tests/golden/benchmark2_patterns.npz")
    st, wt = z["states"].astype(np.int32), z["weights"].astype(np.int64)
    sites = np.repeat(np.arange(st.shape[1]), wt)
    np.random.default_rng(9).shuffle(sites)
    aln = st[:, sites]
    pats, w, idx = beagle.compressSitePatterns(aln)
    m = em.Patterns.fromAlignment(aln)
    assert np.array_equal(pats, m.states) and np.array_equal(w, m.weights)
    assert pats.shape[1] == st.shape[1] and w.sum() == wt.sum()
    assert np.array_equal(pats[:, idx], aln)
    with pytest.raises(beagle.BeagleException):
        beagle.compressSitePatterns(aln, resource=0)          # resource 0 (host) is not implemented: no fallback
