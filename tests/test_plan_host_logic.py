"""Host logic of updatePartials (no GPU): the execution plan the engine derives from an operation list --
phases of disjoint subtrees for post-order lists, the mirrored out-forest plan (or depth levels) for pre-order lists -- must be a permutation of the
list in which every op runs after the ops producing its inputs, ops of one phase's different subtrees are
independent, and hazard lists fall back to the caller's order."""
import numpy as np
import pytest

import helpers as H
from beast_mcmc_b200 import beagle, build
from harness import evomodel as em, treedatalikelihood as tdl


@pytest.fixture(scope="module")
def lib():
    build.build_engine()
    return beagle.load_library()


def plan(lib, ops, nbuf, T=0, want=8, minT=4, small=24, pre=0):
    ops = np.ascontiguousarray(ops, dtype=np.int32).reshape(-1)
    n = len(ops) // 7
    order = np.zeros(max(n, 1), dtype=np.int32)
    subs = np.zeros(2 * max(n, 1), dtype=np.int32)
    ph = np.zeros(n + 2, dtype=np.int32)
    cnt = np.zeros(2, dtype=np.int32)
    ip = lambda a: a.ctypes.data_as(beagle._IP)
    rc = lib.b200DebugPlan(ip(ops), n, nbuf, T, want, minT, small, pre, ip(order), ip(subs), ip(ph), ip(cnt))
    assert rc == 0
    return order[:n], subs[:2 * cnt[0]].reshape(-1, 2), ph[:cnt[1] + 1]


def tree_ops(tips, seed, traversal):
    tree = em.Tree.coalescent(tips, 1.0, seed)
    like = tdl.TreeDataLikelihood.__new__(tdl.TreeDataLikelihood)
    like.tree, like.traversalType, like.updateNode = tree, traversal, np.ones(tree.nodeCount, dtype=bool)
    like._dispatch()
    ops = []
    for node, c1, c2 in like.nodeOperations:
        ops += [node, -1, -1, c1, c1, c2, c2]
    return tree, np.array(ops, dtype=np.int32)


def check_valid(ops, order, subs, phases):
    ops = ops.reshape(-1, 7)
    n = len(ops)
    assert sorted(order.tolist()) == list(range(n))                       # a permutation
    pos_of = {int(order[p]): p for p in range(n)}
    writer = {int(o[0]): k for k, o in enumerate(ops)}
    sub_of, phase_of = {}, {}
    covered = np.zeros(n, dtype=int)
    for ph in range(len(phases) - 1):
        for s in range(phases[ph], phases[ph + 1]):
            b, e = subs[s]
            covered[b:e] += 1
            for p in range(b, e):
                sub_of[p], phase_of[p] = s, ph
    assert (covered == 1).all()                                           # subtrees tile the positions exactly
    for k, o in enumerate(ops):
        for src in (int(o[3]), int(o[5])):
            w = writer.get(src)
            if w is None:
                continue
            pk, pw = pos_of[k], pos_of[w]
            assert phase_of[pw] <= phase_of[pk]
            if phase_of[pw] == phase_of[pk]:                              # same launch: must be the same walk, earlier
                assert sub_of[pw] == sub_of[pk] and pw < pk


@pytest.mark.parametrize("tips,seed,traversal", [(2, 1, "POST_ORDER"), (3, 2, "POST_ORDER"), (50, 3, "REVERSE_LEVEL_ORDER"),
                                                 (400, 4, "POST_ORDER"), (1000, 5, "REVERSE_LEVEL_ORDER")])
@pytest.mark.parametrize("want,minT", [(1, 4), (8, 4), (64, 1)])
def test_post_order_plans_are_valid(lib, tips, seed, traversal, want, minT):
    tree, ops = tree_ops(tips, seed, traversal)
    order, subs, phases = plan(lib, ops, 2 * tips, want=want, minT=minT)
    check_valid(ops, order, subs, phases)
    if tips >= 400 and want >= 8:
        assert len(phases) - 1 >= 2 and (subs[:, 1] - subs[:, 0]).max() < len(ops) // 2     # real parallelism was found
        first = subs[phases[0]:phases[1]]
        lens = first[:, 1] - first[:, 0]
        assert (np.diff(lens) <= 0).all()                                 # longest subtree walks are dispatched first


def test_caterpillar_and_single_op(lib):
    n = 40                                                                # comb tree: no parallelism exists
    ops = []
    for k in range(n):
        ops += [100 + k, -1, -1, (99 + k) if k else 0, 0, k + 1, 0]
    ops = np.array(ops, dtype=np.int32)
    order, subs, phases = plan(lib, ops, 200, want=16, minT=4)
    check_valid(ops, order, subs, phases)
    assert (order == np.arange(n)).all()
    o1 = np.array([5, -1, -1, 0, 0, 1, 1], dtype=np.int32)
    order, subs, phases = plan(lib, o1, 10)
    assert order.tolist() == [0] and subs.tolist() == [[0, 1]]


def test_hazard_lists_keep_caller_order(lib):
    base = [10, -1, -1, 0, 0, 1, 1, 11, -1, -1, 10, 2, 2, 3]
    double_write = np.array(base + [10, -1, -1, 2, 4, 3, 5], dtype=np.int32)        # buffer 10 written twice
    read_before_write = np.array([11, -1, -1, 10, 2, 2, 3, 10, -1, -1, 0, 0, 1, 1], dtype=np.int32)
    shared_result = np.array(base + [12, -1, -1, 10, 4, 3, 5], dtype=np.int32)        # buffer 10 consumed by two ops
    for ops in (double_write, read_before_write, shared_result):
        order, subs, phases = plan(lib, ops, 20)
        assert order.tolist() == list(range(len(ops) // 7)) and len(subs) == 1 and len(phases) == 2


def test_pre_order_levels(lib):
    tree = em.Tree.coalescent(60, 1.0, 9)
    off = tree.nodeCount                                                  # pre-order partial of node k lives at off + k
    ops, stack = [], [(tree.root, -1, -1)]
    while stack:
        node, parent, sib = stack.pop()
        if parent >= 0:
            ops += [off + node, -1, -1, off + parent, node, sib, sib]
        if not tree.isExternal(node):
            c1, c2 = int(tree.child[node][0]), int(tree.child[node][1])
            stack += [(c2, node, c1), (c1, node, c2)]
    ops = np.array(ops, dtype=np.int32)
    order, subs, phases = plan(lib, ops, 2 * off, pre=2)                  # pre=2: the level plan (fallback / B200_PRE_PHASES=0)
    check_valid(ops, order, subs, phases)
    assert len(phases) - 1 == tree.depth()                                # one launch per depth level
    assert (subs[:, 1] - subs[:, 0] == 1).all()


def preorder_ops(tree):
    off = tree.nodeCount
    ops, stack = [], [(tree.root, -1, -1)]
    while stack:
        node, parent, sib = stack.pop()
        if parent >= 0:
            ops += [off + node, -1, -1, off + parent, node, sib, sib]
        if not tree.isExternal(node):
            c1, c2 = int(tree.child[node][0]), int(tree.child[node][1])
            stack += [(c2, node, c1), (c1, node, c2)]
    return np.array(ops, dtype=np.int32), 2 * off


@pytest.mark.parametrize("tips,seed", [(2, 1), (3, 2), (60, 9), (500, 4), (1000, 6)])
@pytest.mark.parametrize("want,minT", [(1, 4), (8, 4), (64, 1)])
def test_pre_order_phased_subtrees(lib, tips, seed, want, minT):
    """Pre-order lists as phased subtree walks: valid (parents before children, a launch's walks independent), far
    fewer launches than depth levels, and inside a walk the op after a node is one of its children whenever it has one
    there (that is what register forwarding of pre[parent] relies on)."""
    tree = em.Tree.coalescent(tips, 1.0, seed)
    ops, nbuf = preorder_ops(tree)
    order, subs, phases = plan(lib, ops, nbuf, want=want, minT=minT, pre=1)
    check_valid(ops, order, subs, phases)
    if tips >= 60:
        assert len(phases) - 1 < tree.depth()
    o = ops.reshape(-1, 7)
    forwarded = 0
    for b, e in subs:
        for p in range(b + 1, e):
            if o[order[p]][3] == o[order[p - 1]][0]:
                forwarded += 1
    internal_nonroot = sum(1 for k in range(len(o)) if (o[:, 3] == o[k][0]).any())
    if want == 1:                                                         # one walk: every node with children forwards once
        assert forwarded == internal_nonroot + (0 if len(o) < 2 else 0) or forwarded >= internal_nonroot - 1
    # hazards fall back to the level plan, still valid
    bad = ops.copy(); bad[7 + 0] = bad[0]                                 # second op writes the first op's destination
    order, subs, phases = plan(lib, bad, nbuf, pre=1)
    assert sorted(order.tolist()) == list(range(len(bad) // 7))
