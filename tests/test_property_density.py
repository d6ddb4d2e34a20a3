"""Property tests (hypothesis) on random small typed trees: the three oracle variants agree on
CPU; on the GPU the kernel agrees with the oracle for the same random trees, including trees
with many exactly-equal event times."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import multitypetree_b200 as mtt
from oracle import oracle as orc


def random_typed_tree(rng, n, d, tie_prob):
    """Random binary tree with random types and valid random change lists.  With probability
    tie_prob a time is snapped to a coarse grid so that many events tie exactly."""
    def snap(x):
        return round(x * 4) / 4 if rng.random() < tie_prob else x
    N = 2 * n - 1
    height = np.zeros(N)
    parent = np.full(N, -1, np.int32)
    nodeType = rng.integers(0, d, N).astype(np.uint8)
    for i in range(n):
        height[i] = snap(rng.uniform(0, 1.0)) if rng.random() < 0.5 else 0.0
    active = list(range(n))
    for v in range(n, N):
        a, b = rng.choice(len(active), 2, replace=False)
        ca, cb = active[a], active[b]
        height[v] = max(height[ca], height[cb]) + snap(rng.exponential(0.5))
        parent[ca] = parent[cb] = v
        active = [x for x in active if x not in (ca, cb)] + [v]
    times, types = [[] for _ in range(N)], [[] for _ in range(N)]
    for v in range(N - 1):
        p = parent[v]
        lo, hi = height[v], height[p]
        k = int(rng.integers(0, 4))
        cur = int(nodeType[v])
        ts = sorted(snap(rng.uniform(lo, hi)) for _ in range(k)) if hi > lo else [lo] * k
        ts = [min(max(t, lo), hi) for t in ts]
        for t in ts:
            nxt = int((cur + 1 + rng.integers(0, d - 1)) % d)
            times[v].append(float(t))
            types[v].append(nxt)
            cur = nxt
        if cur != nodeType[p]:                       # make the branch end in the parent's type
            times[v].append(float(hi))
            types[v].append(int(nodeType[p]))
    return mtt.MultiTypeTree(height, parent, nodeType, times, types)


def random_model(rng, d):
    m = rng.uniform(0.05, 2.0, (d, d))
    np.fill_diagonal(m, 0.0)
    return m, rng.uniform(0.5, 5.0, d)


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(seed=st.integers(0, 2**31 - 1), n=st.integers(2, 40), d=st.integers(2, 6),
       ties=st.sampled_from([0.0, 0.5, 1.0]))
def test_oracle_variants_agree_on_random_trees(seed, n, d, ties):
    rng = np.random.default_rng(seed)
    t = random_typed_tree(rng, n, d, ties)
    m, N = random_model(rng, d)
    tb = mtt.TreeBatch.from_trees([t])
    a = orc.sc_logp_batch(tb, m, N, variant=orc.FAITHFUL, checkValidity=True)[0]
    b = orc.sc_logp_batch(tb, m, N, variant=orc.SORTED)[0]
    c = orc.sc_logp_batch(tb, m, N, variant=orc.SUFFSTAT)[0]
    assert np.isfinite(a)                            # the generator only makes valid trees
    assert b == pytest.approx(a, rel=1e-12, abs=1e-12)
    assert c == pytest.approx(a, rel=1e-10, abs=1e-10)


@pytest.mark.gpu
@pytest.mark.parametrize("ties", [0.0, 0.5, 1.0])
def test_gpu_matches_oracle_on_random_trees(ctx, ties):
    rng = np.random.default_rng(int(ties * 10) + 99)
    for d in (2, 3, 5, 9):
        trees = [random_typed_tree(rng, int(rng.integers(2, 120)), d, ties) for _ in range(150)]
        m, N = random_model(rng, d)
        tb = mtt.TreeBatch.from_trees(trees)
        want = orc.sc_logp_batch(tb, m, N, variant=orc.FAITHFUL, checkValidity=True)
        got, valid = mtt.sc_logp_batch(tb, m, N, checkValidity=True, ctx=ctx, want_valid=True)
        assert valid.all() and np.all(np.isfinite(want))
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(mtt.sc_logp_batch(tb, m, N, ctx=ctx), want, rtol=1e-9,
                                   atol=1e-12)
