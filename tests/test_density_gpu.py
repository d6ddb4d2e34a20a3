"""GPU parity tests for piece 1 (csrc/k1_density.cu) through the C-ABI, against the oracle.
Tolerance: |logP_gpu - logP_oracle| <= 1e-9 * |logP_oracle| (north_star: 1e-9 relative, fp64)."""
import numpy as np
import pytest

import multitypetree_b200 as mtt
from oracle import oracle as orc
from tests.util import concat_batches, load_density_fixture

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _check(tb, m, N, ctx, **kw):
    """both wire layouts of the change lists (per-change owners, per-node counts) against the
    oracle; they must also agree with each other bit for bit (same kernels behind both)"""
    want = orc.sc_logp_batch(tb, m, N, variant=orc.FAITHFUL)
    got = mtt.sc_logp_batch(tb, m, N, ctx=ctx, layout="owner", **kw)
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=0)
    alt = mtt.sc_logp_batch(tb, m, N, ctx=ctx, layout="counts", **kw)
    assert np.array_equal(got, alt, equal_nan=True)
    return got


def test_reference_kat(ctx):
    """SCLikelihoodTest.java:43-73 through the plugin mirror: -16.52831 within 1e-5."""
    kat = ("(((A[&state=1]:0.25)[&state=0]:0.25,B[&state=0]:0.5)[&state=0]:1.5,"
           "(C[&state=0]:1.0,D[&state=0]:1.0)[&state=0]:1.0)[&state=0]:0.0;")
    mtTree = mtt.MultiTypeTreeFromNewick(kat, "state")
    migrationModel = mtt.SCMigrationModel([2.0, 1.0], [5.0, 10.0], mtt.TypeSet("A", "B"))
    likelihood = mtt.StructuredCoalescentTreeDensity(migrationModel=migrationModel,
                                                    multiTypeTree=mtTree, ctx=ctx)
    result = likelihood.calculateLogP()
    assert abs(result - (-16.52831)) < 1e-5
    assert abs(result - (-11.7 - 3 * np.log(5))) < 1e-12


@pytest.mark.parametrize("name", ["kat", "true128", "true64", "ewing3"])
def test_golden_fixtures(ctx, name):
    tb, m, N, expected = load_density_fixture(name)
    got = mtt.sc_logp_batch(tb, m, N, ctx=ctx)
    tol = 1e-5 if name == "kat" else RTOL * abs(expected)
    assert abs(got[0] - expected) <= tol
    got_v, valid = mtt.sc_logp_batch(tb, m, N, checkValidity=True, ctx=ctx, want_valid=True)
    assert valid[0] and got_v[0] == got[0]


@pytest.mark.parametrize("n,d,B", [(2, 2, 64), (3, 2, 64), (5, 3, 256), (50, 2, 512),
                                   (100, 2, 4096), (128, 2, 512), (200, 4, 256),
                                   (250, 3, 128), (100, 10, 256), (60, 16, 128)])
def test_simulated_small_trees(ctx, n, d, B):
    """warp-per-tree path (E <= 512), incl. config 2 (100 tips, 2 demes, 4096 trees)."""
    m, N = mtt.synthetic_model(d)
    tb = mtt.simulate_trees(B, n, d, m, N, seed=1000 * 2 + n)
    _check(tb, m, N, ctx)
    _check(tb, m, N, ctx, checkValidity=True)


@pytest.mark.parametrize("n,d,B", [(300, 2, 64), (400, 3, 200), (350, 10, 64), (1000, 4, 32),
                                   (5000, 8, 8), (7000, 2, 2)])
def test_simulated_large_trees(ctx, n, d, B):
    """1024-event warp stage (512 < E <= 1024) and the CTA-per-tree path (E <= 16384):
    configs 3 and 4 tree sizes."""
    m, N = mtt.synthetic_model(d)
    tb = mtt.simulate_trees(B, n, d, m, N, seed=3000 + n)
    _check(tb, m, N, ctx)
    _check(tb, m, N, ctx, checkValidity=True)


def test_mixed_sizes_and_ragged_batch(ctx):
    """Small and large trees interleaved in one batch (warp kernel + queue + block kernel)."""
    m, N = mtt.synthetic_model(3)
    parts = [mtt.simulate_trees(b, n, 3, m, N, seed=77 + n)
             for n, b in ((4, 5), (400, 3), (90, 40), (1500, 2), (127, 9), (2, 3), (170, 20))]
    tb = concat_batches(parts)
    assert tb.n_events().min() < 16 and tb.n_events().max() > 3000
    _check(tb, m, N, ctx)
    _check(tb, m, N, ctx, checkValidity=True)


def test_per_tree_models(ctx):
    """nModels == B: one migration model per tree (batched proposals / chains)."""
    rng = np.random.default_rng(5)
    d, B = 4, 300
    m0, N0 = mtt.synthetic_model(d)
    tb = mtt.simulate_trees(B, 80, d, m0, N0, seed=9)
    m = m0[None] * rng.uniform(0.5, 2.0, size=(B, d, d))
    N = N0[None] * rng.uniform(0.5, 2.0, size=(B, d))
    got = mtt.sc_logp_batch(tb, m, N, ctx=ctx)
    want = orc.sc_logp_batch(tb, m, N)
    np.testing.assert_allclose(got, want, rtol=RTOL)
    big = mtt.simulate_trees(6, 700, d, m0, N0, seed=10)
    got = mtt.sc_logp_batch(big, m[:6], N[:6], ctx=ctx)
    np.testing.assert_allclose(got, orc.sc_logp_batch(big, m[:6], N[:6]), rtol=RTOL)


def test_zero_rate_gives_minus_inf(ctx):
    """BSSVS: a switched-off rate on an observed migration gives log(0) = -inf."""
    tb, m, N, _ = load_density_fixture("kat")
    m = m.copy()
    m[1, 0] = 0.0
    got = mtt.sc_logp_batch(tb, m, N, ctx=ctx)
    assert got[0] == -np.inf and orc.sc_logp_batch(tb, m, N)[0] == -np.inf


def test_invalid_trees(ctx):
    """MultiTypeTree.isValid (MultiTypeTree.java:396-429) + checkValidity (:90-91)."""
    tb, m, N, _ = load_density_fixture("true128")
    good = tb.tree(0)
    rng = np.random.default_rng(3)
    trees, expect = [], []
    for k in range(40):
        t = good.copy()
        kind = k % 4
        if kind == 0:
            expect.append(True)
        elif kind == 1:      # node type differs from what its children's branches end in
            node = int(rng.integers(t.getLeafNodeCount(), t.getNodeCount()))
            t.nodeType[node] = 1 - t.nodeType[node]
            expect.append(False)
        elif kind == 2:      # a change above its parent's height
            node = next(i for i in range(t.getNodeCount())
                        if t.getChangeCount(i) > 0 and t.parent[i] >= 0)
            t.changeTimes[node][-1] = float(t.height[t.parent[node]]) + 1.0
            expect.append(False)
        else:                # child older than parent
            node = int(rng.integers(0, t.getLeafNodeCount()))
            t.height[node] = float(t.height[t.parent[node]]) + 0.5
            expect.append(False)
        trees.append(t)
    b = mtt.TreeBatch.from_trees(trees)
    lp, valid = mtt.sc_logp_batch(b, m, N, checkValidity=True, ctx=ctx, want_valid=True)
    want = orc.sc_logp_batch(b, m, N, checkValidity=True)
    assert list(valid) == expect
    assert np.array_equal(np.isneginf(lp), np.isneginf(want))
    ok = np.array(expect)
    np.testing.assert_allclose(lp[ok], want[ok], rtol=RTOL)
    assert good.isValid(ctx)
    assert not trees[1].isValid(ctx)


def test_quantisation_ties_are_repaired(ctx):
    """Events closer than the 24-bit time quantum of the packed sort key must still be ordered
    exactly: a huge root height squeezes all other events into a few quanta."""
    m, N = mtt.synthetic_model(2)
    tb = mtt.simulate_trees(64, 100, 2, m, N, seed=4242)
    tb.height = tb.height.copy()
    nN = 199
    roots = np.arange(64) * nN + nN - 1
    tb.height[roots] += 1e7          # quantum becomes ~0.6 time units
    _check(tb, m, N, ctx)
    # same for the CTA path (38-bit quantum): root at 1e13
    big = mtt.simulate_trees(4, 600, 2, m, N, seed=4243)
    big.height = big.height.copy()
    big.height[np.arange(4) * 1199 + 1198] += 1e13
    _check(big, m, N, ctx)


def test_root_branch_changes_ignored(ctx):
    """StructuredCoalescentTreeDensity.java:161-162: changes above the root do not count."""
    tb, m, N, expected = load_density_fixture("kat")
    t = tb.tree(0)
    root = t.getRoot()
    b = mtt.TreeBatch.from_trees([t])
    # hand-build a batch that DOES carry root changes (from_trees never exports them)
    cnt = b.chgCount.copy()
    cnt[root] = 2
    b2 = mtt.TreeBatch(b.nodeOff, b.height, b.parent, b.nodeType, cnt,
                       np.array([0, 3], np.int64), np.append(b.chgTime, [2.5, 3.0]),
                       np.append(b.chgType, [1, 0]).astype(np.uint8))
    got = mtt.sc_logp_batch(b2, m, N, ctx=ctx, with_parent=True)
    assert abs(got[0] - expected) < 1e-5


def test_malformed_change_lists(ctx):
    """Counts that do not add up to chgBase, or owners outside the tree: never a crash; the
    tree is reported invalid and -inf under checkValidity, in both layouts."""
    tb, m, N, _ = load_density_fixture("true64")
    bad = mtt.TreeBatch(tb.nodeOff, tb.height, tb.parent, tb.nodeType, tb.chgCount.copy(),
                        tb.chgBase, tb.chgTime, tb.chgType)
    bad.chgCount[5] += 3
    lp, valid = mtt.sc_logp_batch(bad, m, N, checkValidity=True, ctx=ctx, want_valid=True,
                                  layout="counts")
    assert not valid[0] and lp[0] == -np.inf
    with pytest.raises(ValueError):
        bad.chgNode
    own = mtt.TreeBatch(tb.nodeOff, tb.height, tb.parent, tb.nodeType, tb.chgCount, tb.chgBase,
                        tb.chgTime, tb.chgType)
    o = tb.chgNode.copy()
    o[3] = 60000
    own.set_owner_list(o)
    lp, valid = mtt.sc_logp_batch(own, m, N, checkValidity=True, ctx=ctx, want_valid=True)
    assert not valid[0] and lp[0] == -np.inf
    assert np.isfinite(mtt.sc_logp_batch(own, m, N, ctx=ctx)[0])     # ignored change, no crash


def test_batch_size_independence_and_determinism(ctx):
    m, N = mtt.synthetic_model(2)
    tb = mtt.simulate_trees(20000, 100, 2, m, N, seed=31337)
    a = mtt.sc_logp_batch(tb, m, N, ctx=ctx)
    b = mtt.sc_logp_batch(tb, m, N, ctx=ctx)
    assert np.array_equal(a, b)                       # bit-reproducible run to run
    part = mtt.sc_logp_batch(tb.slice(777, 1500), m, N, ctx=ctx)
    assert np.array_equal(part, a[777:1500])          # independent of batch composition
    want = orc.sc_logp_batch(tb, m, N, variant=orc.SORTED, nThreads=8)
    np.testing.assert_allclose(a, want, rtol=RTOL)


def test_api_errors(ctx):
    tb, m, N, _ = load_density_fixture("kat")
    with pytest.raises(mtt.MttError):
        mtt.sc_logp_batch(tb, np.zeros((17, 17)), np.ones(17), ctx=ctx)      # d > MTT_MAX_DEMES
    with pytest.raises(mtt.MttError):
        mtt.sc_logp_batch(tb, m, N, checkValidity=True, ctx=ctx, with_parent=False)
    assert mtt.sc_logp_batch(tb.slice(0, 0), m, N, ctx=ctx).shape == (0,)   # empty batch


def _pad_changes(tree, target_M):
    """Add back-and-forth migration pairs on the longest branches until the tree has exactly
    target_M changes (parity permitting); keeps the tree valid."""
    t = tree.copy()
    d_other = lambda c: 1 - c if c < 2 else 0
    M = t.getTotalNumberOfChanges()
    order = np.argsort([-(t.height[t.parent[i]] - t.height[i]) if t.parent[i] >= 0 else 1
                        for i in range(t.getNodeCount())])
    k = 0
    while M + 2 <= target_M:
        b = int(order[k % 20])
        k += 1
        lo = t.changeTimes[b][-1] if t.changeTimes[b] else float(t.height[b])
        hi = float(t.height[t.parent[b]])
        cur = t.getFinalType(b)
        t.addChange(b, d_other(cur), lo + (hi - lo) * 0.3)
        t.addChange(b, cur, lo + (hi - lo) * 0.6)
        M += 2
    return t


@pytest.mark.parametrize("n,targets", [(128, (255, 256, 257, 258)), (250, (510, 511, 512, 513, 514)),
                                       (64, (127, 128, 129)), (190, (381, 382, 383, 384)),
                                       (190, (383, 384, 385, 386)), (150, (299, 300, 383, 384)),
                                       (450, (1021, 1022, 1023, 1024)),
                                       (450, (1023, 1024, 1025, 1026)),
                                       (300, (599, 767, 768, 769, 770))])
def test_event_count_boundaries(ctx, n, targets):
    """E = nodes + changes right at the 256 / 384 / 512 / 768 / 1024 limits of the warp path (one
    run, two runs + merge in the 384- and 512-event stages, three and four runs with two merge
    levels in the 1024-event stage, hand-over to the CTA kernel)."""
    m, N = mtt.synthetic_model(2)
    pool = mtt.simulate_trees(64, n, 2, m * 0.01, N, seed=600 + n)      # few migrations
    trees, want_E = [], []
    for E in targets:
        for parity in (0, 1):
            for b in range(pool.B):
                t = pool.tree(b)
                M0 = t.getTotalNumberOfChanges()
                need = E - (2 * n - 1)
                if need >= M0 and (need - M0) % 2 == 0:
                    trees.append(_pad_changes(t, need))
                    want_E.append(E)
                    break
            break
    assert trees
    tb = mtt.TreeBatch.from_trees(trees)
    assert list(tb.n_events()) == want_E
    _check(tb, m, N, ctx)
    _check(tb, m, N, ctx, checkValidity=True)


def test_ties_and_zero_length_branches(ctx):
    """Equal event times (contemporaneous tips, zero-length branches, a change exactly at a
    node height): interval terms vanish (delta_t > 0 is false, :108) but every event term and
    count change must still be applied."""
    nw = ("((A[&s=0]:1.0,B[&s=0]:1.0)[&s=0]:0.0,((C[&s=1]:0.5)[&s=0]:0.5,D[&s=0]:1.0)[&s=0]:0.0)"
          "[&s=0]:0.0;")
    t = mtt.MultiTypeTreeFromNewick(nw, "s")
    m = np.array([[0, 0.7], [0.4, 0]])
    N = np.array([3.0, 2.0])
    tb = mtt.TreeBatch.from_trees([t] * 5)
    got = _check(tb, m, N, ctx, checkValidity=True)
    assert np.all(got == got[0]) and np.isfinite(got[0])
    want = orc.sc_logp_batch(tb, m, N, variant=orc.SUFFSTAT)
    np.testing.assert_allclose(got, want, rtol=1e-12)


def test_single_deme(ctx):
    """d = 1 (the two-deme specialisation with an empty second deme): Kingman's coalescent."""
    m, N = np.zeros((1, 1)), np.array([1.7])
    tb = mtt.simulate_trees(64, 40, 1, m, N, seed=11)
    assert tb.chgTime.size == 0
    got = _check(tb, m, N, ctx)
    assert np.all(np.isfinite(got))
    _check(tb, m, N, ctx, checkValidity=True)


def test_types_out_of_range_and_self_changes(ctx):
    """A node / change type >= d is malformed input: NaN, never a silently wrong number.  A
    change into the type it leaves has rate m(i,i) = 0 in the reference
    (SCMigrationModel.java:253-255): log(0) = -infinity."""
    m, N = mtt.synthetic_model(2)
    pool = mtt.simulate_trees(8, 30, 2, m, N, seed=5)
    trees = [pool.tree(i) for i in range(8)]
    trees[1].nodeType[3] = 7                                   # leaf type outside the model
    trees[2].nodeType[40] = 2                                  # internal node type = d
    b3 = next(i for i in range(59) if trees[3].getChangeCount(i) > 0)
    trees[3].changeTypes[b3][0] = 9                            # change type outside the model
    b4 = next(i for i in range(59) if trees[4].getChangeCount(i) > 0)
    trees[4].changeTypes[b4][0] = int(trees[4].nodeType[b4])   # self change
    tb = mtt.TreeBatch.from_trees(trees)
    for chk in (False, True):
        lp = mtt.sc_logp_batch(tb, m, N, checkValidity=chk, ctx=ctx)
        assert np.isnan(lp[1]) and np.isnan(lp[2]) and np.isnan(lp[3])
        assert lp[4] == -np.inf
        ok = [0, 5, 6, 7]
        np.testing.assert_allclose(lp[ok], orc.sc_logp_batch(pool, m, N)[ok], rtol=RTOL)
    assert orc.sc_logp_batch(tb.slice(4, 5), m, N)[0] == -np.inf
    # the same for more than two demes (the general descriptor path)
    m4, N4 = mtt.synthetic_model(4)
    pool = mtt.simulate_trees(4, 30, 4, m4, N4, seed=6)
    trees = [pool.tree(i) for i in range(4)]
    trees[1].nodeType[2] = 4
    b2 = next(i for i in range(59) if trees[2].getChangeCount(i) > 0)
    trees[2].changeTypes[b2][0] = int(trees[2].nodeType[b2])
    lp = mtt.sc_logp_batch(mtt.TreeBatch.from_trees(trees), m4, N4, ctx=ctx)
    assert np.isnan(lp[1]) and lp[2] == -np.inf
    np.testing.assert_allclose(lp[[0, 3]], orc.sc_logp_batch(pool, m4, N4)[[0, 3]], rtol=RTOL)


def test_event_bound_hint_is_a_promise(ctx):
    """mtt_sc_set_max_tree_events on the device-pointer entry: a tree beyond the promised bound
    is reported as NaN (never left unwritten), and the next call without a bound is unaffected."""
    import torch
    from multitypetree_b200 import _lib
    L = _lib.lib()
    m, N = mtt.synthetic_model(2)
    small = mtt.simulate_trees(30, 50, 2, m, N, seed=1)
    big = mtt.simulate_trees(2, 300, 2, m, N, seed=2)
    tb = mtt.TreeBatch.from_trees([small.tree(i) for i in range(15)] + [big.tree(0)] +
                                  [small.tree(i) for i in range(15, 30)] + [big.tree(1)])
    want = orc.sc_logp_batch(tb, m, N)

    def dev(a):
        t = torch.from_numpy(np.ascontiguousarray(a).ravel())
        out = torch.zeros(t.numel() + 64, dtype=t.dtype, device="cuda")
        out[:t.numel()] = t.cuda()
        return out
    arrs = [dev(x) for x in (m, N, tb.nodeOff, tb.height, tb.nodeType, tb.chgNode, tb.chgBase,
                             tb.chgTime, tb.chgType)]
    p = [a.data_ptr() for a in arrs]
    try:
        for hint in (200, 0, 300, 0):
            logP = torch.full((tb.B,), 123.0, dtype=torch.float64, device="cuda")
            ctx.check(L.mtt_sc_set_max_tree_events(ctx.handle, hint))
            ctx.check(L.mtt_sc_logp_owner_batch_dev(ctx.handle, tb.B, 2, 1, p[0], p[1], p[2], p[3],
                                                    None, p[4], p[5], p[6], p[7], p[8], 0,
                                                    logP.data_ptr(), None))
            ctx.synchronize()
            got = logP.cpu().numpy()
            beyond = tb.n_events() > hint if hint else np.zeros(tb.B, bool)
            assert beyond.sum() == (2 if hint else 0)
            assert np.all(np.isnan(got[beyond]))
            np.testing.assert_allclose(got[~beyond], want[~beyond], rtol=RTOL)
        # checkValidity without parent[] is refused by the device-pointer entry too
        rc = L.mtt_sc_logp_owner_batch_dev(ctx.handle, tb.B, 2, 1, p[0], p[1], p[2], p[3], None,
                                           p[4], p[5], p[6], p[7], p[8], 1, logP.data_ptr(), None)
        assert rc != 0
    finally:
        L.mtt_sc_set_max_tree_events(ctx.handle, 0)


def test_fixed_tips_entry_is_bit_identical(ctx):
    """mtt_sc_set_tips + mtt_sc_logp_tips_batch: trees over one taxon set, tips installed once;
    the results are the general entry's bit for bit (shared and per-tree models, validity)."""
    for n, d, B in ((100, 2, 3000), (30, 4, 500), (400, 3, 40), (1000, 4, 6)):
        m, N = mtt.synthetic_model(d)
        tb = mtt.simulate_trees(B, n, d, m, N, seed=40 + n)
        mtt.set_tips(tb.height[:n], tb.nodeType[:n], ctx=ctx)
        want, wv = mtt.sc_logp_batch(tb, m, N, checkValidity=True, ctx=ctx, want_valid=True)
        got, gv = mtt.sc_logp_tips_batch(tb, m, N, checkValidity=True, ctx=ctx, want_valid=True)
        assert np.array_equal(got, want) and np.array_equal(gv, wv)
        assert np.array_equal(mtt.sc_logp_tips_batch(tb, m, N, ctx=ctx),
                              mtt.sc_logp_batch(tb, m, N, ctx=ctx))
    # one model per tree
    m, N = mtt.synthetic_model(2)
    tb = mtt.simulate_trees(64, 50, 2, m, N, seed=3)
    mm = np.stack([m * (1 + 0.01 * i) for i in range(64)])
    NN = np.stack([N * (1 + 0.02 * i) for i in range(64)])
    mtt.set_tips(tb.height[:50], tb.nodeType[:50], ctx=ctx)
    assert np.array_equal(mtt.sc_logp_tips_batch(tb, mm, NN, ctx=ctx),
                          mtt.sc_logp_batch(tb, mm, NN, ctx=ctx))
    # a tree over other tips is refused by the mirror; a fresh context has no tips
    other = mtt.simulate_trees(2, 50, 2, m, N, leafTimes=np.arange(50) * 0.1, seed=4)
    with pytest.raises(ValueError):
        mtt.sc_logp_tips_batch(other, m, N, ctx=ctx)
    c2 = mtt.Context(0)
    with pytest.raises(ValueError):
        mtt.sc_logp_tips_batch(tb, m, N, ctx=c2)
    from multitypetree_b200 import _lib
    rc = _lib.lib().mtt_sc_logp_tips_batch(c2.handle, 1, 2, 1, _lib.ptr(m), _lib.ptr(N),
                                           _lib.ptr(tb.height), None, _lib.ptr(tb.nodeType),
                                           _lib.ptr(tb.chgNode), _lib.ptr(tb.chgBase),
                                           _lib.ptr(tb.chgTime), _lib.ptr(tb.chgType), 0,
                                           _lib.ptr(np.zeros(1)), None)
    assert rc != 0
    c2.close()


def _relabel(t, rng):
    """The same tree with leaves permuted among themselves and internal nodes among themselves
    (BEAST keeps leaves first; everything else about node numbers is arbitrary after a few
    topology moves: internal nodes are not sorted by height and the root is not the last)."""
    n, N = t.getLeafNodeCount(), t.getNodeCount()
    new = np.concatenate([rng.permutation(n), n + rng.permutation(N - n)])     # old -> new
    height, parent = np.zeros(N), np.full(N, -1, np.int32)
    nodeType = np.zeros(N, np.uint8)
    times, types = [None] * N, [None] * N
    for old in range(N):
        k = int(new[old])
        height[k], nodeType[k] = t.height[old], t.nodeType[old]
        parent[k] = -1 if t.parent[old] < 0 else int(new[t.parent[old]])
        times[k], types[k] = list(t.changeTimes[old]), list(t.changeTypes[old])
    return mtt.MultiTypeTree(height, parent, nodeType, times, types)


@pytest.mark.parametrize("n,d,B", [(100, 2, 256), (60, 5, 128), (300, 2, 24), (1000, 4, 6)])
def test_node_numbering_does_not_matter(ctx, n, d, B):
    """Register path, shared-memory path and CTA path on trees whose nodes are numbered at
    random: same logP as the simulator's numbering (internal nodes by height, root last) up to
    the order of the floating-point sums, and as the oracle."""
    m, N = mtt.synthetic_model(d)
    tb = mtt.simulate_trees(B, n, d, m, N, seed=900 + n)
    rng = np.random.default_rng(n)
    shuffled = mtt.TreeBatch.from_trees([_relabel(tb.tree(i), rng) for i in range(B)])
    roots = np.nonzero(shuffled.parent.reshape(B, -1) < 0)[1]
    assert np.any(roots != 2 * n - 2)
    base = mtt.sc_logp_batch(tb, m, N, checkValidity=True, ctx=ctx)
    got, valid = mtt.sc_logp_batch(shuffled, m, N, checkValidity=True, ctx=ctx, want_valid=True)
    assert valid.all()
    np.testing.assert_allclose(got, base, rtol=1e-12)
    np.testing.assert_allclose(got, orc.sc_logp_batch(shuffled, m, N, checkValidity=True),
                               rtol=RTOL)
    np.testing.assert_allclose(mtt.sc_logp_batch(shuffled, m, N, ctx=ctx), base, rtol=1e-12)


def test_degenerate_trees_inside_a_batch(ctx):
    """A single-leaf tree (one SAMPLE event, which is event 0 and skipped: logP = 0) and an empty
    tree (no nodes) between ordinary trees: no crash, no effect on the neighbours."""
    m, N = mtt.synthetic_model(2)
    tb = mtt.simulate_trees(4, 20, 2, m, N, seed=8)
    want = orc.sc_logp_batch(tb, m, N)
    nN = 39
    one = np.array([0.3])
    nodeOff = np.array([0, nN, nN + 1, nN + 1, 2 * nN + 1, 3 * nN + 1, 4 * nN + 1], np.int64)
    height = np.concatenate([tb.height[:nN], one, tb.height[nN:]])
    parent = np.concatenate([tb.parent[:nN], [-1], tb.parent[nN:]]).astype(np.int32)
    nodeType = np.concatenate([tb.nodeType[:nN], [1], tb.nodeType[nN:]]).astype(np.uint8)
    chgCount = np.concatenate([tb.chgCount[:nN], [0], tb.chgCount[nN:]]).astype(np.uint16)
    cb = tb.chgBase
    chgBase = np.array([cb[0], cb[1], cb[1], cb[1], cb[2], cb[3], cb[4]], np.int64)
    mixed = mtt.TreeBatch(nodeOff, height, parent, nodeType, chgCount, chgBase, tb.chgTime,
                          tb.chgType)
    for chk in (False, True):
        got = mtt.sc_logp_batch(mixed, m, N, checkValidity=chk, ctx=ctx)
        assert got[1] == 0.0 and got[2] == 0.0
        np.testing.assert_allclose(got[[0, 3, 4, 5]], want, rtol=RTOL)
