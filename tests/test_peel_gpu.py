"""GPU parity tests for piece 2 (csrc/k2_peel.cu) through the C-ABI, against the oracle
(BeerLikelihoodCore semantics).  Tolerance: 1e-9 relative on logL (north_star)."""
import numpy as np
import pytest

import multitypetree_b200 as mtt
from multitypetree_b200.likelihood import (Alignment, HKY, SiteModel, StrictClockModel,
                                           TreeLikelihood, simulate_alignment)
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _trees(B, n, d=2, seed=1):
    m, N = mtt.synthetic_model(d)
    tb = mtt.simulate_trees(B, n, d, m, N, seed=seed)
    nN = 2 * n - 1
    return tb.parent.reshape(B, nN), tb.height.reshape(B, nN), tb


def _oracle(al, par, hgt, sm, clock=1.0, nThreads=8):
    rates, props = sm.category_rates_and_proportions()
    return orc.tree_likelihood_batch(al.patterns, al.patternWeight, par, hgt, rates, props,
                                     sm.substModel.frequencies, sm.substModel.kappa,
                                     branchRate=clock, pInv=sm.proportionInvariant,
                                     useScaling=2, nThreads=nThreads)


@pytest.mark.parametrize("n,sites,gamma,pInv,missing,B", [
    (2, 50, 0, 0.0, 0.0, 4), (3, 200, 0, 0.0, 0.0, 8), (5, 300, 4, 0.0, 0.1, 16),
    (17, 1000, 0, 0.2, 0.0, 16), (64, 2000, 4, 0.1, 0.05, 8), (128, 2000, 0, 0.0, 0.0, 8),
    (100, 700, 2, 0.0, 0.0, 8), (33, 400, 8, 0.3, 0.2, 4), (200, 513, 3, 0.0, 0.0, 5),
    (40, 300, 5, 0.0, 0.1, 3), (40, 300, 6, 0.1, 0.0, 3), (300, 260, 7, 0.0, 0.0, 3)])
def test_logl_matches_oracle(ctx, n, sites, gamma, pInv, missing, B):
    par, hgt, _ = _trees(B, n, seed=10 + n)
    f = np.array([.1, .2, .3, .4])
    tips = simulate_alignment(par[0], hgt[0], sites, 2.5, f, 0.05, seed=n, missing=missing)
    al = Alignment(states=tips)
    sm = SiteModel(HKY(2.5, f), mutationRate=0.05, gammaCategoryCount=gamma, shape=0.6,
                   proportionInvariant=pInv)
    tl = TreeLikelihood(data=al, siteModel=sm, ctx=ctx)
    got = tl.calculateLogPBatch(par, hgt)
    want = _oracle(al, par, hgt, sm)
    assert np.all(np.isfinite(want))
    np.testing.assert_allclose(got, want, rtol=RTOL)
    tl.close()


def test_config1_shape_tree_likelihood_object(ctx):
    """config 1 shape (128 taxa, 2000 sites, HKY, 1 category, strict clock rate 1) through
    the TreeLikelihood mirror with a MultiTypeTree as its `tree` Input."""
    par, hgt, tb = _trees(1, 128, seed=128)
    tree = tb.tree(0)
    f = np.array([.25] * 4)
    tips = simulate_alignment(par[0], hgt[0], 2000, 3.0, f, 0.01, seed=3)
    al = Alignment(states=tips)
    assert al.getPatternCount() < 2000
    sm = SiteModel(HKY(3.0, f), mutationRate=0.01)
    tl = TreeLikelihood(data=al, tree=tree, siteModel=sm, ctx=ctx)
    lp = tl.calculateLogP()
    want = _oracle(al, par, hgt, sm)[0]
    assert lp == pytest.approx(want, rel=RTOL)
    pl = tl.getPatternLogLikelihoods()
    rates, props = sm.category_rates_and_proportions()
    _, pl_want = orc.tree_likelihood(al.patterns, al.patternWeight, par[0], hgt[0], rates, props,
                                     f, 3.0, want_patterns=True)
    np.testing.assert_allclose(pl, pl_want, rtol=RTOL)
    assert (pl * al.patternWeight).sum() == pytest.approx(lp, rel=1e-12)
    # a parameter move: kappa and clock rate change, same handle
    sm.substModel.kappa = 5.0
    tl.branchRateModel = StrictClockModel(1.7)
    tl.update_models()
    assert tl.calculateLogP() == pytest.approx(_oracle(al, par, hgt, sm, clock=1.7)[0], rel=RTOL)
    tl.close()


@pytest.mark.parametrize("n,P,rate", [(1000, 1500, 0.02), (1000, 300, 50.0), (3000, 256, 0.01),
                                      (5000, 130, 0.005)])
def test_large_trees_need_rescaling(ctx, n, P, rate):
    """configs 3/4 tree sizes: underflow is certain without per-node rescaling."""
    par, hgt, _ = _trees(2, n, d=4, seed=n)
    f = np.array([.3, .2, .2, .3])
    tips = simulate_alignment(par[0], hgt[0], P, 2.0, f, rate, seed=5)
    al = Alignment(states=tips, compress=False)
    sm = SiteModel(HKY(2.0, f), mutationRate=rate)
    tl = TreeLikelihood(data=al, siteModel=sm, ctx=ctx)
    got = tl.calculateLogPBatch(par, hgt)
    want = _oracle(al, par, hgt, sm)
    assert np.all(np.isfinite(want))
    np.testing.assert_allclose(got, want, rtol=RTOL)
    tl.close()


@pytest.mark.parametrize("n,rate,gamma", [(300, 50.0, 0), (420, 50.0, 4), (1000, 50.0, 0),
                                          (1000, 0.02, 2), (600, 50.0, 0)])
def test_proportion_invariant_with_rescaling(ctx, n, rate, gamma):
    """pInv > 0 on trees whose partials need rescaling, with constant and all-missing
    patterns.  BEAST returns the evaluation WITHOUT rescaling when it is finite (n = 300-420
    at saturation: ~4^-n per pattern, below the 1e-100 rescaling threshold but above the double
    range's floor) and the rescaled one -- which adds pInv to the already rescaled root
    partials -- only after a -infinity (n = 1000).  The kernel always rescales and must
    return BEAST's number in both regimes (oracle: useScaling=2)."""
    par, hgt, _ = _trees(2, n, d=2, seed=7 * n)
    f = np.array([.3, .2, .2, .3])
    tips = simulate_alignment(par[0], hgt[0], 120, 2.0, f, rate, seed=n, missing=0.02)
    const = np.repeat(np.arange(4, dtype=np.int32)[None, :], n, axis=0)      # AAAA.., CCCC.., ...
    holes = const.copy()
    holes[::3] = 4                                                           # constant + gaps
    allmiss = np.full((n, 1), 4, np.int32)
    al = Alignment(states=np.concatenate([tips, const, holes, allmiss], axis=1), compress=False)
    sm = SiteModel(HKY(2.0, f), mutationRate=rate, gammaCategoryCount=gamma, shape=0.5,
                   proportionInvariant=0.25)
    tl = TreeLikelihood(data=al, siteModel=sm, ctx=ctx)
    got = tl.calculateLogPBatch(par, hgt)
    want = _oracle(al, par, hgt, sm)
    assert np.all(np.isfinite(want))
    np.testing.assert_allclose(got, want, rtol=RTOL)
    # per-pattern values of the first tree follow the same choice
    tl.tree = type("T", (), {"parent": par[0], "height": hgt[0]})()
    rates, props = sm.category_rates_and_proportions()
    _, pat = orc.tree_likelihood(al.patterns, al.patternWeight, par[0], hgt[0], rates, props, f,
                                 2.0, pInv=0.25, useScaling=2, want_patterns=True)
    np.testing.assert_allclose(tl.getPatternLogLikelihoods(), pat, rtol=RTOL, atol=1e-12)
    tl.close()


def test_unbalanced_and_balanced_topologies(ctx):
    """Caterpillar (depth n) and perfectly balanced trees: the traversal planner must keep the
    number of live partial vectors within its log2(n)+1 bound for both."""
    f = np.array([.25] * 4)
    for n, kind in ((257, "cat"), (256, "bal"), (1000, "cat")):
        N = 2 * n - 1
        parent = np.full(N, -1, np.int32)
        height = np.zeros(N)
        if kind == "cat":
            prev = 0
            for i in range(1, n):
                node = n + i - 1
                parent[prev] = node
                parent[i] = node
                height[node] = 0.01 * i
                prev = node
        else:
            level = list(range(n))
            nxt, hh = n, 0.02
            while len(level) > 1:
                new = []
                for a, b in zip(level[::2], level[1::2]):
                    parent[a] = parent[b] = nxt
                    height[nxt] = hh
                    new.append(nxt)
                    nxt += 1
                level, hh = new, hh + 0.02
        tips = simulate_alignment(parent, height, 300, 2.0, f, 1.0, seed=n)
        al = Alignment(states=tips)
        sm = SiteModel(HKY(2.0, f))
        tl = TreeLikelihood(data=al, siteModel=sm, ctx=ctx)
        got = tl.calculateLogPBatch(parent[None], height[None])
        want = _oracle(al, parent[None], height[None], sm)
        np.testing.assert_allclose(got, want, rtol=RTOL)
        tl.close()


def test_deep_trees_in_sub_batches(ctx):
    """Balanced 256-tip trees park 8 partial vectors: two in shared memory, six in the global
    scratch.  With 20 k patterns and 700 trees that scratch passes 1 GB, so the batch runs in
    sub-batches (of 276 or 331 trees, for 6 or 5 scratch slots); trees on both sides of every
    possible seam must still match the oracle."""
    n, B, sites = 256, 700, 20000
    N = 2 * n - 1
    parent = np.full(N, -1, np.int32)
    height = np.zeros(N)
    level, nxt, hh = list(range(n)), n, 0.02
    while len(level) > 1:
        new = []
        for a, b in zip(level[::2], level[1::2]):
            parent[a] = parent[b] = nxt
            height[nxt] = hh
            new.append(nxt)
            nxt += 1
        level, hh = new, hh + 0.02
    f = np.array([.25] * 4)
    rng = np.random.default_rng(8)
    states = rng.integers(0, 4, size=(n, sites)).astype(np.uint8)      # all columns distinct
    al = Alignment(states=states, compress=False)
    sm = SiteModel(HKY(2.0, f))
    tl = TreeLikelihood(data=al, siteModel=sm, ctx=ctx)
    par = np.repeat(parent[None], B, axis=0)
    hgt = height[None] * (1.0 + 0.001 * np.arange(B)[:, None])
    got = tl.calculateLogPBatch(par, hgt)
    pick = [0, 1, 274, 275, 276, 277, 329, 330, 331, 332, 550, 551, 552, 553, 661, 662, 663, 699]
    want = _oracle(al, par[pick], hgt[pick], sm)
    np.testing.assert_allclose(got[pick], want, rtol=RTOL)
    assert np.all(np.diff(got) != 0) and np.all(np.isfinite(got))
    tl.close()


def test_per_tree_substitution_parameters(ctx):
    """mtt_peel_logl_batch_params: one (kappa, clock rate) per tree -- B parameter proposals or
    B chains in a single call -- against the oracle evaluated tree by tree."""
    n, B = 60, 12
    par, hgt, _ = _trees(B, n, seed=91)
    f = np.array([.15, .35, .3, .2])
    tips = simulate_alignment(par[0], hgt[0], 600, 2.0, f, 0.05, seed=4, missing=0.03)
    al = Alignment(states=tips)
    for gamma in (0, 4):
        sm = SiteModel(HKY(2.0, f), mutationRate=0.05, gammaCategoryCount=gamma, shape=0.7)
        tl = TreeLikelihood(data=al, siteModel=sm, ctx=ctx)
        rng = np.random.default_rng(gamma)
        kappa = rng.uniform(0.5, 6.0, B)
        rate = rng.uniform(0.3, 3.0, B)
        got = tl.calculateLogPBatchParams(par, hgt, kappa, rate)
        rates, props = sm.category_rates_and_proportions()
        want = np.array([orc.tree_likelihood(al.patterns, al.patternWeight, par[b], hgt[b],
                                             rates, props, f, kappa[b], branchRate=rate[b])
                         for b in range(B)])
        np.testing.assert_allclose(got, want, rtol=RTOL)
        # rate omitted = the handle's clock rate; a scalar kappa broadcasts
        same = tl.calculateLogPBatchParams(par, hgt, 2.0)
        np.testing.assert_allclose(same, tl.calculateLogPBatch(par, hgt), rtol=1e-14)
        tl.close()


def test_explicit_matrices_mode(ctx):
    """mtt_peel_logl_batch_mats: caller-provided per-node transition matrices (any model)."""
    n, B = 40, 6
    par, hgt, _ = _trees(B, n, seed=77)
    f = np.array([.1, .2, .3, .4])
    tips = simulate_alignment(par[0], hgt[0], 500, 2.0, f, 0.1, seed=1)
    al = Alignment(states=tips)
    sm = SiteModel(HKY(2.0, f), mutationRate=0.1, gammaCategoryCount=4, shape=0.8)
    rates, props = sm.category_rates_and_proportions()
    N = 2 * n - 1
    mats = np.zeros((B, N, 4, 16))
    for b in range(B):
        for v in range(N):
            if par[b, v] >= 0:
                for c in range(4):
                    mats[b, v, c] = orc.hky_matrix(2.0, f, (hgt[b, par[b, v]] - hgt[b, v])
                                                   * rates[c]).ravel()
    tl = TreeLikelihood(data=al, siteModel=sm, ctx=ctx)
    got = tl.calculateLogPBatchMats(par, mats)
    want = np.array([orc.tree_likelihood_mats(al.patterns, al.patternWeight, par[b], mats[b],
                                              props, f) for b in range(B)])
    np.testing.assert_allclose(got, want, rtol=RTOL)
    np.testing.assert_allclose(got, tl.calculateLogPBatch(par, hgt), rtol=1e-12)
    tl.close()


def test_pattern_shards_sum_to_whole(ctx):
    """config 4 decomposition: logL over pattern shards adds up to the unsharded logL (the
    all-reduce of SURVEY 8e, emulated on one GPU as a sum over shard handles)."""
    n, B = 50, 4
    par, hgt, _ = _trees(B, n, seed=5)
    f = np.array([.25] * 4)
    al = Alignment(states=simulate_alignment(par[0], hgt[0], 3000, 2.0, f, 0.1, seed=9))
    sm = SiteModel(HKY(2.0, f), mutationRate=0.1, gammaCategoryCount=4, shape=0.5)
    whole = TreeLikelihood(data=al, siteModel=sm, ctx=ctx)
    full = whole.calculateLogPBatch(par, hgt)
    acc = np.zeros(B)
    for r in range(8):
        sh = TreeLikelihood(data=al.shard(r, 8), siteModel=sm, ctx=ctx)
        acc += sh.calculateLogPBatch(par, hgt)
        sh.close()
    np.testing.assert_allclose(acc, full, rtol=1e-12)
    np.testing.assert_allclose(full, _oracle(al, par, hgt, sm), rtol=RTOL)
    whole.close()


def test_relaxed_clock_branch_rates(ctx):
    """mtt_peel_logl_batch_rates: one clock rate per branch (UCRelaxedClockModel of
    fxtemplates/MTTClockModels.xml).  Equal rates reproduce the strict clock bit for bit; random
    rates agree with the oracle fed the corresponding HKY matrices."""
    n, B = 40, 6
    par, hgt, _ = _trees(B, n, seed=12)
    f = np.array([.15, .35, .3, .2])
    al = Alignment(states=simulate_alignment(par[0], hgt[0], 600, 2.2, f, 0.1, seed=4, missing=0.05))
    sm = SiteModel(HKY(2.2, f), mutationRate=0.1, gammaCategoryCount=4, shape=0.7,
                   proportionInvariant=0.1)
    strict = TreeLikelihood(data=al, siteModel=sm, branchRateModel=StrictClockModel(1.7), ctx=ctx)
    base = strict.calculateLogPBatch(par, hgt)
    assert np.array_equal(strict.calculateLogPBatchRates(par, hgt, np.full(hgt.shape, 1.7)), base)
    rng = np.random.default_rng(3)
    rates = rng.lognormal(0.0, 0.5, hgt.shape)
    got = strict.calculateLogPBatchRates(par, hgt, rates)
    catRates, props = sm.category_rates_and_proportions()
    N = 2 * n - 1
    for b in range(B):
        mats = np.zeros((N, len(catRates), 16))
        for v in range(N):
            if par[b, v] < 0:
                continue
            for c, cr in enumerate(catRates):
                dist = (hgt[b, par[b, v]] - hgt[b, v]) * cr * rates[b, v]
                mats[v, c] = orc.hky_matrix(2.2, f, dist).ravel()
        want = orc.tree_likelihood_mats(al.patterns, al.patternWeight, par[b], mats, props, f,
                                        pInv=0.1)
        assert got[b] == pytest.approx(want, rel=RTOL)
    assert np.all(np.abs(got - base) > 1e-3)
    strict.close()


def test_pattern_shards_with_invariant_sites(ctx):
    """pInv > 0 on a tree whose unscaled evaluation underflows for SOME patterns only: the
    choice between BEAST's two evaluations is global, so the shards return both sums
    (mtt_peel_logl_parts_batch) and the choice is made after adding them up."""
    n, B = 1000, 2
    par, hgt, _ = _trees(B, n, seed=77)
    f = np.array([.3, .2, .2, .3])
    tips = simulate_alignment(par[0], hgt[0], 200, 2.0, f, 50.0, seed=2)
    const = np.repeat(np.arange(4, dtype=np.int32)[None, :], n, axis=0)
    al = Alignment(states=np.concatenate([const, tips, const], axis=1), compress=False)
    sm = SiteModel(HKY(2.0, f), mutationRate=50.0, proportionInvariant=0.2)
    want = _oracle(al, par, hgt, sm)
    assert np.all(np.isfinite(want))
    parts = np.zeros((3, B))
    for r in range(4):
        sh = TreeLikelihood(data=al.shard(r, 4), siteModel=sm, ctx=ctx)
        parts += sh.calculateLogPParts(par, hgt)
        sh.close()
    assert np.all(parts[2] > 0)                     # some shard saw the underflow
    got = np.where(parts[2] > 0, parts[1], parts[0])
    np.testing.assert_allclose(got, want, rtol=RTOL)
    whole = TreeLikelihood(data=al, siteModel=sm, ctx=ctx)
    np.testing.assert_allclose(whole.calculateLogPBatch(par, hgt), want, rtol=RTOL)
    whole.close()
    # a shard of constant patterns alone would have picked the other evaluation
    first = TreeLikelihood(data=Alignment(states=const, compress=False), siteModel=sm, ctx=ctx)
    p0 = first.calculateLogPParts(par, hgt)
    assert np.all(p0[2] == 0) and np.all(np.abs(p0[0] - p0[1]) > 1.0)
    first.close()


def test_malformed_tree_is_reported_not_silently_wrong(ctx):
    n = 10
    par, hgt, _ = _trees(2, n, seed=3)
    par = par.copy()
    par[1, 0] = par[1, 1]            # a node with three children / another with one
    f = np.array([.25] * 4)
    al = Alignment(states=simulate_alignment(par[0], hgt[0], 100, 2.0, f, 0.1, seed=9))
    tl = TreeLikelihood(data=al, siteModel=SiteModel(HKY(2.0, f)), ctx=ctx)
    got = tl.calculateLogPBatch(par, hgt)
    assert np.isfinite(got[0]) and np.isnan(got[1])
    with pytest.raises(mtt.MttError):
        _ = mtt.lib().mtt_peel_logl_batch  # noqa
        bad = TreeLikelihood.__new__(TreeLikelihood)
        bad.ctx, bad._h = ctx, mtt._lib.VP()
        import ctypes
        ctx.check(mtt.lib().mtt_peel_create(ctx.handle, 4, 10, 9, ctypes.byref(bad._h)))
    tl.close()


def test_determinism_and_batch_independence(ctx):
    n, B = 100, 12
    par, hgt, _ = _trees(B, n, seed=1)
    f = np.array([.25] * 4)
    al = Alignment(states=simulate_alignment(par[0], hgt[0], 1200, 2.0, f, 0.05, seed=2))
    tl = TreeLikelihood(data=al, siteModel=SiteModel(HKY(2.0, f), 0.05), ctx=ctx)
    a = tl.calculateLogPBatch(par, hgt)
    b = tl.calculateLogPBatch(par, hgt)
    c = tl.calculateLogPBatch(par[3:7], hgt[3:7])
    assert np.array_equal(a, b) and np.array_equal(a[3:7], c)
    tl.close()


@pytest.mark.parametrize("gamma", [0, 4])
def test_random_topologies(ctx, gamma):
    """Arbitrary (non-coalescent) topologies with dated tips and random data incl. missing
    states: every tree of a batch has its own traversal plan."""
    from tests.test_property_density import random_typed_tree
    rng = np.random.default_rng(17 + gamma)
    n, B = 37, 60
    trees = [random_typed_tree(rng, n, 2, 0.0) for _ in range(B)]
    par = np.stack([t.parent for t in trees])
    hgt = np.stack([t.height for t in trees])
    f = np.array([.15, .35, .3, .2])
    tips = rng.integers(0, 5, size=(n, 333)).astype(np.int32)      # 4 = missing
    al = Alignment(states=tips)
    sm = SiteModel(HKY(4.0, f), mutationRate=0.3, gammaCategoryCount=gamma, shape=0.4,
                   proportionInvariant=0.1)
    tl = TreeLikelihood(data=al, siteModel=sm, ctx=ctx)
    got = tl.calculateLogPBatch(par, hgt)
    want = _oracle(al, par, hgt, sm)
    np.testing.assert_allclose(got, want, rtol=RTOL)
    tl.close()
