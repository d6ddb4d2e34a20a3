"""GPU tests of the resident chains (csrc/k2_update.cu, mtt_peel_chains_*): incremental
dirty-node peeling with accept / reject must return what a full re-evaluation of the same
proposal returns (mtt_peel_logl_batch*, itself checked against the oracle in test_peel_gpu.py)
-- to rounding: the log scale factors are summed per subtree instead of along the traversal.
Tolerance 1e-12 relative against the full evaluation, 1e-9 against the oracle (north_star)."""
import numpy as np
import pytest

import multitypetree_b200 as mtt
from multitypetree_b200.likelihood import (Alignment, HKY, SiteModel, TreeLikelihood,
                                           simulate_alignment)
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
RTOL_FULL = 1e-12
RTOL = 1e-9


def _trees(B, n, d=2, seed=1):
    m, N = mtt.synthetic_model(d)
    tb = mtt.simulate_trees(B, n, d, m, N, seed=seed)
    nN = 2 * n - 1
    return tb.parent.reshape(B, nN).copy(), tb.height.reshape(B, nN).copy()


def _kids(par, v):
    return [int(c) for c in np.nonzero(par == v)[0]]


def _propose(rng, par, hgt, n):
    """One BEAST-like tree move on (par, hgt) in place; returns the nodes it touched."""
    N = 2 * n - 1
    root = int(np.nonzero(par < 0)[0][0])
    move = rng.integers(0, 5)
    if move == 0:                                   # node height (MultiTypeUniform)
        v = int(rng.integers(n, N))
        lo = max(hgt[c] for c in _kids(par, v))
        hi = hgt[par[v]] if par[v] >= 0 else lo + 2.0
        hgt[v] = lo + (hi - lo) * rng.uniform(0.05, 0.95)
        return [v]
    if move == 1:                                   # scale everything (MultiTypeTreeScale)
        f = rng.uniform(0.8, 1.25)
        old = hgt.copy()
        hgt[n:] = hgt[n:] * f
        ok = all(hgt[v] < hgt[par[v]] for v in range(N) if par[v] >= 0)
        if not ok:                                  # a tip above its parent: no move
            hgt[:] = old
            return []
        return list(range(n, N))
    if move == 2:                                   # exchange a node with its uncle
        for _ in range(20):
            i = int(rng.integers(0, N))
            p = int(par[i])
            if p < 0 or par[p] < 0:
                continue
            g = int(par[p])
            u = [c for c in _kids(par, g) if c != p][0]
            if hgt[u] < hgt[p]:
                par[i], par[u] = g, p
                return [i, u]
        return []
    if move == 3:                                   # new root: the old root's internal child
        kids = _kids(par, root)
        for a in kids:
            if a < n:
                continue
            b = [c for c in kids if c != a][0]
            x, y = _kids(par, a)
            if hgt[a] > hgt[b]:
                par[a], par[root], par[y] = -1, a, root
                hgt[a], hgt[root] = hgt[root], hgt[a]
                return [a, root, y, b, x]
        return []
    return []                                       # nothing


@pytest.mark.parametrize("n,sites,gamma,pInv,useDirty", [(40, 500, 0, 0.0, False),
                                                         (40, 500, 4, 0.1, True),
                                                         (17, 300, 3, 0.0, True),
                                                         (24, 200, 8, 0.0, True),
                                                         (24, 200, 6, 0.1, False),
                                                         (2, 60, 0, 0.0, True),
                                                         (3, 60, 2, 0.1, False),
                                                         (150, 260, 0, 0.2, False)])
def test_chains_follow_full_evaluation(ctx, n, sites, gamma, pInv, useDirty):
    K, steps = 6, 40
    rng = np.random.default_rng(n + gamma)
    par, hgt = _trees(K, n, seed=3 + n)
    f = np.array([.1, .2, .3, .4])
    tips = simulate_alignment(par[0], hgt[0], sites, 2.5, f, 0.05, seed=n, missing=0.05)
    al = Alignment(states=tips)
    sm = SiteModel(HKY(2.5, f), mutationRate=0.05, gammaCategoryCount=gamma, shape=0.6,
                   proportionInvariant=pInv)
    tl = TreeLikelihood(data=al, siteModel=sm, ctx=ctx)
    ch = tl.chains(K)
    kap = np.full(K, 2.5)
    got = ch.update(par, hgt, kappa=kap)
    assert ch.last_entries == K * (n - 1)
    want = tl.calculateLogPBatchParams(par, hgt, kap)
    np.testing.assert_allclose(got, want, rtol=RTOL_FULL)
    ch.accept()
    accLogL = got.copy()
    sawPartial = False
    for step in range(steps):
        ppar, phgt, pkap = par.copy(), hgt.copy(), kap.copy()
        dirty = np.zeros(par.shape, np.uint8)
        for k in range(K):
            if rng.random() < 0.1:
                pkap[k] = rng.uniform(1.0, 4.0)     # substitution parameter move: all dirty
            else:
                for v in _propose(rng, ppar[k], phgt[k], n):
                    dirty[k, v] = 1
        got = ch.update(ppar, phgt, kappa=pkap, dirty=dirty if useDirty else None)
        want = tl.calculateLogPBatchParams(ppar, phgt, pkap)
        np.testing.assert_allclose(got, want, rtol=RTOL_FULL)
        sawPartial |= 0 < ch.last_entries < K * (n - 1)
        same = [k for k in range(K) if np.array_equal(ppar[k], par[k]) and
                np.array_equal(phgt[k], hgt[k]) and pkap[k] == kap[k]]
        assert np.array_equal(got[same], accLogL[same])     # untouched chains: the accepted value
        acc = rng.random(K) < 0.5
        if step % 7 == 3:
            acc[:] = False                          # (and let the next update imply the reject)
        else:
            ch.accept(acc)
        par[acc], hgt[acc], kap[acc], accLogL[acc] = ppar[acc], phgt[acc], pkap[acc], got[acc]
    assert sawPartial or n <= 3
    # the accepted states survived all of that
    got = ch.update(par, hgt, kappa=kap)
    assert ch.last_entries == 0
    np.testing.assert_allclose(got, tl.calculateLogPBatchParams(par, hgt, kap), rtol=RTOL_FULL)
    rates, props = sm.category_rates_and_proportions()
    want = orc.tree_likelihood_batch(al.patterns, al.patternWeight, par, hgt, rates, props, f,
                                     2.5, branchRate=1.0, pInv=pInv, useScaling=2, nThreads=8)
    ok = kap == 2.5
    np.testing.assert_allclose(got[ok], want[ok], rtol=RTOL)
    ch.close()
    tl.close()


@pytest.mark.parametrize("n,P,rate,pInv", [(1000, 300, 50.0, 0.0), (420, 260, 50.0, 0.2),
                                           (1000, 1500, 0.02, 0.0)])
def test_chains_with_rescaling(ctx, n, P, rate, pInv):
    """Large / saturated trees: per-node rescaling, cumulative scale factors per subtree, and the
    choice between BEAST's two evaluations under a proportion of invariant sites."""
    K = 2
    rng = np.random.default_rng(n)
    par, hgt = _trees(K, n, d=4, seed=n)
    f = np.array([.3, .2, .2, .3])
    tips = simulate_alignment(par[0], hgt[0], P, 2.0, f, rate, seed=5)
    al = Alignment(states=tips, compress=False)
    sm = SiteModel(HKY(2.0, f), mutationRate=rate, proportionInvariant=pInv)
    tl = TreeLikelihood(data=al, siteModel=sm, ctx=ctx)
    ch = tl.chains(K)
    got = ch.update(par, hgt)
    want = _oracle(al, par, hgt, sm)
    assert np.all(np.isfinite(want))
    np.testing.assert_allclose(got, want, rtol=RTOL)
    np.testing.assert_allclose(got, tl.calculateLogPBatch(par, hgt), rtol=RTOL_FULL)
    ch.accept()
    for step in range(6):
        ppar, phgt = par.copy(), hgt.copy()
        for k in range(K):
            _propose(rng, ppar[k], phgt[k], n)
        got = ch.update(ppar, phgt)
        np.testing.assert_allclose(got, tl.calculateLogPBatch(ppar, phgt), rtol=RTOL_FULL)
        np.testing.assert_allclose(got, _oracle(al, ppar, phgt, sm), rtol=RTOL)
        acc = np.array([step % 2 == 0, step % 3 == 0])
        ch.accept(acc)
        par[acc], hgt[acc] = ppar[acc], phgt[acc]
    ch.close()
    tl.close()


def _oracle(al, par, hgt, sm, clock=1.0):
    rates, props = sm.category_rates_and_proportions()
    return orc.tree_likelihood_batch(al.patterns, al.patternWeight, par, hgt, rates, props,
                                     sm.substModel.frequencies, sm.substModel.kappa,
                                     branchRate=clock, pInv=sm.proportionInvariant,
                                     useScaling=2, nThreads=8)


def test_chains_relaxed_clock_and_model_changes(ctx):
    n, K = 30, 3
    N = 2 * n - 1
    rng = np.random.default_rng(8)
    par, hgt = _trees(K, n, seed=21)
    f = np.array([.25, .25, .25, .25])
    tips = simulate_alignment(par[0], hgt[0], 400, 2.0, f, 0.1, seed=2)
    al = Alignment(states=tips)
    sm = SiteModel(HKY(2.0, f), mutationRate=0.1)
    tl = TreeLikelihood(data=al, siteModel=sm, ctx=ctx)
    ch = tl.chains(K)
    rates = rng.lognormal(0.0, 0.3, size=(K, N))
    got = ch.update(par, hgt, branchRates=rates)
    np.testing.assert_allclose(got, tl.calculateLogPBatchRates(par, hgt, rates), rtol=RTOL_FULL)
    ch.accept()
    # one branch rate per chain changes: only the path above it is recomputed
    r2 = rates.copy()
    for k in range(K):
        r2[k, rng.integers(0, n)] *= 1.7
    got = ch.update(par, hgt, branchRates=r2)
    assert 0 < ch.last_entries < K * (n - 1)
    np.testing.assert_allclose(got, tl.calculateLogPBatchRates(par, hgt, r2), rtol=RTOL_FULL)
    ch.accept(np.array([1, 0, 1]))
    rates[[0, 2]] = r2[[0, 2]]
    # back to a strict clock: every branch rate differs
    got = ch.update(par, hgt, clockRate=np.array([1.0, 1.3, 0.7]))
    assert ch.last_entries == K * (n - 1)
    np.testing.assert_allclose(got, tl.calculateLogPBatchParams(par, hgt, 2.0, [1.0, 1.3, 0.7]),
                               rtol=RTOL_FULL)
    ch.accept()
    got2 = ch.update(par, hgt, clockRate=np.array([1.0, 1.3, 0.7]))
    assert ch.last_entries == 0 and np.array_equal(got, got2)
    # a setter of the handle (new kappa for everybody) makes every resident partial stale
    sm.substModel.kappa = 3.5
    tl.update_models()
    got = ch.update(par, hgt, clockRate=np.array([1.0, 1.3, 0.7]))
    assert ch.last_entries == K * (n - 1)
    np.testing.assert_allclose(got, tl.calculateLogPBatchParams(par, hgt, 3.5, [1.0, 1.3, 0.7]),
                               rtol=RTOL_FULL)
    ch.close()
    tl.close()


def test_chains_malformed_proposal_leaves_the_chain_alone(ctx):
    n, K = 12, 4
    N = 2 * n - 1
    par, hgt = _trees(K, n, seed=5)
    f = np.array([.25, .25, .25, .25])
    tips = simulate_alignment(par[0], hgt[0], 300, 2.0, f, 0.1, seed=2)
    tl = TreeLikelihood(data=Alignment(states=tips), siteModel=SiteModel(HKY(2.0, f), 0.1), ctx=ctx)
    ch = tl.chains(K)
    # a malformed tree as the very first proposal of a chain: NaN, and the chain starts over
    first = par.copy()
    first[2, 0] = 0                                 # a node as its own parent
    got = ch.update(first, hgt)
    assert np.isnan(got[2]) and np.all(np.isfinite(np.delete(got, 2)))
    ch.accept()
    base = ch.update(par, hgt)
    assert ch.last_entries == n - 1                 # chain 2 had nothing accepted yet
    np.testing.assert_allclose(base, tl.calculateLogPBatch(par, hgt), rtol=RTOL_FULL)
    ch.accept()
    bad = par.copy()
    root = [int(np.nonzero(par[k] < 0)[0][0]) for k in range(K)]
    bad[0, 0] = 1                                   # a tip as a parent
    bad[1, root[1]] = n                             # no root any more (and a cycle)
    bad[2, 5] = N + 3                               # out of range
    # chain 3: three children for one node
    p3 = int(par[3, 0])
    other = [i for i in range(N) if par[3, i] != p3 and par[3, i] >= 0][0]
    bad[3, other] = p3
    got = ch.update(bad, hgt)
    assert np.all(np.isnan(got))
    ch.accept()                                     # nothing to accept
    again = ch.update(par, hgt)
    assert ch.last_entries == 0 and np.array_equal(again, base)
    # a cycle between internal nodes, detached from the root
    cyc = par.copy()
    k = 0
    inner = [v for v in range(n, N) if par[k, v] >= 0 and par[k, par[k, v]] >= 0]
    v = inner[0]
    p = int(par[k, v])
    cyc[k, p] = v                                   # v <-> p
    x = int(np.nonzero(par[k] == v)[0][0])
    cyc[k, x] = int(par[k, p])                      # every node keeps two children
    got = ch.update(cyc, hgt)
    assert np.isnan(got[0]) and np.array_equal(got[1:], base[1:])
    again = ch.update(par, hgt)
    assert np.array_equal(again, base)
    ch.close()
    tl.close()


def test_chains_outliving_their_handle_fail_loudly(ctx):
    n, K = 8, 2
    par, hgt = _trees(K, n, seed=9)
    f = np.array([.25, .25, .25, .25])
    tips = simulate_alignment(par[0], hgt[0], 100, 2.0, f, 0.1, seed=2)
    tl = TreeLikelihood(data=Alignment(states=tips), siteModel=SiteModel(HKY(2.0, f), 0.1), ctx=ctx)
    ch = tl.chains(K)
    ch.update(par, hgt)
    tl.close()                                      # frees the chains' device side too
    with pytest.raises(mtt.MttError):
        ch.update(par, hgt)
    with pytest.raises(mtt.MttError):
        ch.accept()
    ch.close()


@pytest.mark.parametrize("n,rate,pInv", [(40, 0.05, 0.0), (1000, 50.0, 0.2)])
def test_chains_on_pattern_shards_sum_to_the_whole(ctx, n, rate, pInv):
    """config-4 regime: each rank holds a contiguous slice of the patterns and the chains' trees;
    the per-shard parts, summed, give the whole alignment's value -- including BEAST's switch to
    the rescaled evaluation, which no shard can decide alone (the 1000-tip case underflows)."""
    from multitypetree_b200 import parallel
    K, world = 3, 3
    rng = np.random.default_rng(n)
    par, hgt = _trees(K, n, d=4, seed=n + 1)
    f = np.array([.3, .2, .2, .3])
    tips = simulate_alignment(par[0], hgt[0], 700, 2.0, f, rate, seed=5)
    al = Alignment(states=tips, compress=False)
    sm = SiteModel(HKY(2.0, f), mutationRate=rate, proportionInvariant=pInv)
    whole = TreeLikelihood(data=al, siteModel=sm, ctx=ctx)
    shards = [TreeLikelihood(data=al.shard(r, world), siteModel=sm, ctx=ctx) for r in range(world)]
    chains = [tl.chains(K) for tl in shards]

    def sharded(par, hgt):
        parts = sum(ch.update(par, hgt, parts=True) for ch in chains)      # the all-reduce
        return np.where(parts[2] > 0, parts[1], parts[0])

    got = sharded(par, hgt)
    np.testing.assert_allclose(got, whole.calculateLogPBatch(par, hgt), rtol=RTOL_FULL)
    np.testing.assert_allclose(got, _oracle(al, par, hgt, sm), rtol=RTOL)
    for ch in chains:
        ch.accept()
    for step in range(5):
        ppar, phgt = par.copy(), hgt.copy()
        for k in range(K):
            _propose(rng, ppar[k], phgt[k], n)
        got = sharded(ppar, phgt)
        assert sum(ch.last_entries for ch in chains) < world * K * (n - 1)
        np.testing.assert_allclose(got, whole.calculateLogPBatch(ppar, phgt), rtol=RTOL_FULL)
        acc = rng.random(K) < 0.5
        for ch in chains:
            ch.accept(acc)
        par[acc], hgt[acc] = ppar[acc], phgt[acc]
    # the helper of parallel.py outside torch.distributed is the single-shard case
    one = whole.chains(K)
    np.testing.assert_allclose(parallel.pattern_sharded_chains_update(one, par, hgt),
                               whole.calculateLogPBatch(par, hgt), rtol=RTOL_FULL)
    for x in chains + [one]:
        x.close()
    for tl in shards + [whole]:
        tl.close()


def test_chains_verify_mode_catches_an_unflagged_change(ctx):
    n, K = 10, 2
    par, hgt = _trees(K, n, seed=12)
    f = np.array([.25, .25, .25, .25])
    tips = simulate_alignment(par[0], hgt[0], 100, 2.0, f, 0.1, seed=2)
    tl = TreeLikelihood(data=Alignment(states=tips), siteModel=SiteModel(HKY(2.0, f), 0.1), ctx=ctx)
    ch = tl.chains(K)
    base = ch.update(par, hgt)
    ch.accept()
    h2 = hgt.copy()
    v = int(par[1, 0])
    up = h2[1, par[1, v]] if par[1, v] >= 0 else h2[1, v] + 1.0
    h2[1, v] = 0.5 * (h2[1, v] + up)
    wrong = np.zeros(par.shape, np.uint8)           # the move is not flagged
    stale = ch.update(par, h2, dirty=wrong)
    assert np.array_equal(stale, base)              # unnoticed, as documented ...
    ch.set_verify(True)
    with pytest.raises(mtt.MttError, match="not flagged dirty"):
        ch.update(par, h2, dirty=wrong)             # ... unless the binding is under test
    right = wrong.copy()
    right[1, v] = 1
    got = ch.update(par, h2, dirty=right)
    np.testing.assert_allclose(got, tl.calculateLogPBatch(par, h2), rtol=RTOL_FULL)
    ch.close()
    tl.close()


def test_chains_sparse_updates(ctx):
    """The proposal as a list of changed nodes (what a BEAST shim has in hand) gives what the
    full-array form gives: same moves, same accept / reject sequence, two sets of chains."""
    n, K, steps = 40, 5, 40
    N = 2 * n - 1
    rng = np.random.default_rng(99)
    par, hgt = _trees(K, n, seed=77)
    f = np.array([.1, .2, .3, .4])
    tips = simulate_alignment(par[0], hgt[0], 400, 2.5, f, 0.05, seed=3)
    tl = TreeLikelihood(data=Alignment(states=tips), siteModel=SiteModel(HKY(2.5, f), 0.05), ctx=ctx)
    dense, sparse = tl.chains(K), tl.chains(K)
    empty = np.zeros(0)
    with pytest.raises(mtt.MttError, match="full update"):      # nothing accepted yet
        sparse.update_sparse(np.zeros(K + 1, np.int64), empty, empty, empty)
    a = dense.update(par, hgt)
    b = sparse.update(par, hgt)
    assert np.array_equal(a, b)
    dense.accept()
    sparse.accept()
    for step in range(steps):
        ppar, phgt = par.copy(), hgt.copy()
        off, node = [0], []
        for k in range(K):
            touched = _propose(rng, ppar[k], phgt[k], n)
            if step % 5 == 0 and touched:
                touched = touched + touched[:1]                 # a node named twice
            node += touched
            off.append(len(node))
        node = np.array(node, np.int64)
        chain = np.repeat(np.arange(K), np.diff(off))
        a = dense.update(ppar, phgt)
        b = sparse.update_sparse(off, node, ppar[chain, node], phgt[chain, node])
        assert np.array_equal(a, b)
        assert dense.last_entries == sparse.last_entries
        acc = rng.random(K) < 0.5
        dense.accept(acc)
        sparse.accept(acc)
        par[acc], hgt[acc] = ppar[acc], phgt[acc]
    # a kappa move cannot be said with a list
    with pytest.raises(mtt.MttError, match="full update"):
        sparse.update_sparse(np.zeros(K + 1, np.int64), empty, empty, empty, kappa=3.0)
    # ... and leaves the chains as they were; an out-of-range node is a malformed proposal
    bad = sparse.update_sparse(np.array([0, 1, 1, 1, 1, 1]), [N + 4], [n], [1.0])
    assert np.isnan(bad[0]) and np.array_equal(bad[1:], dense.update(par, hgt)[1:])
    np.testing.assert_allclose(sparse.update(par, hgt), tl.calculateLogPBatch(par, hgt),
                               rtol=RTOL_FULL)
    assert sparse.last_entries == 0
    for x in (dense, sparse):
        x.close()
    tl.close()
