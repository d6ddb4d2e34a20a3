"""CPU tests: the oracle for piece 2.  No reference test pins peeling (it is BEAST 2 core code,
absent from the reference tree) => "parity unpinned"; the restatement is checked against
brute-force enumeration of internal states with scipy.linalg.expm transition matrices."""
import itertools

import numpy as np
import pytest
import scipy.linalg

import multitypetree_b200 as mtt
from multitypetree_b200.likelihood import SiteModel, HKY, simulate_alignment
from oracle import oracle as orc


def hky_Q(kappa, f):
    Q = np.array([[0 if i == j else f[j] * (kappa if (i % 2) == (j % 2) else 1.0)
                   for j in range(4)] for i in range(4)], float)
    np.fill_diagonal(Q, -Q.sum(1))
    return Q / -(f * np.diag(Q)).sum()          # one expected substitution per unit time


@pytest.mark.parametrize("kappa,f,t", [(2.5, (.1, .2, .3, .4), 0.37), (1.0, (.25,) * 4, 1.3),
                                       (7.0, (.4, .1, .4, .1), 0.002), (2.0, (.3, .2, .2, .3), 50)])
def test_hky_closed_form_matches_expm(kappa, f, t):
    f = np.array(f)
    got = orc.hky_matrix(kappa, f, t)
    np.testing.assert_allclose(got, scipy.linalg.expm(hky_Q(kappa, f) * t), atol=1e-14)


def brute_force(tipStates, weights, parent, height, rates, props, f, kappa, pInv=0.0):
    """Sum over all internal-state assignments; no pruning, no rescaling."""
    n, P = tipStates.shape
    N = 2 * n - 1
    root = int(np.nonzero(parent < 0)[0][0])
    Q = hky_Q(kappa, np.asarray(f))
    total = 0.0
    for p in range(P):
        site = 0.0
        for r, pr in zip(rates, props):
            mats = {v: scipy.linalg.expm(Q * (height[parent[v]] - height[v]) * r)
                    for v in range(N) if v != root}
            lik = 0.0
            for assign in itertools.product(range(4), repeat=n - 1):
                st = {n + k: s for k, s in enumerate(assign)}
                term = f[st[root]]
                for v in range(N):
                    if v == root:
                        continue
                    if v < n:
                        s = tipStates[v, p]
                        if s >= 4:
                            continue              # missing data: sums to 1
                    else:
                        s = st[v]
                    term *= mats[v][st[parent[v]], s]
                lik += term
            site += pr * lik
        if pInv > 0:
            obs = set(int(s) for s in tipStates[:, p] if s < 4)
            if len(obs) <= 1:
                for k in range(4):
                    if not obs or k in obs:
                        site += pInv * f[k]
        total += weights[p] * np.log(site)
    return total


def small_tree(n, seed):
    m, N = mtt.synthetic_model(2)
    tb = mtt.simulate_trees(1, n, 2, m, N, seed=seed, nThreads=1)
    return tb.parent.copy(), tb.height.copy()


@pytest.mark.parametrize("n,gamma,pInv,missing", [(2, 0, 0.0, 0.0), (3, 0, 0.0, 0.0),
                                                  (4, 4, 0.0, 0.0), (5, 0, 0.2, 0.0),
                                                  (4, 3, 0.1, 0.3), (5, 0, 0.0, 0.4)])
def test_oracle_matches_brute_force(n, gamma, pInv, missing):
    rng = np.random.default_rng(n * 10 + gamma)
    parent, height = small_tree(n, 40 + n)
    P = 12
    tips = rng.integers(0, 4, size=(n, P)).astype(np.int32)
    tips[:, 0] = 2                                  # a constant pattern
    if missing:
        tips[rng.random(tips.shape) < missing] = 4
        tips[:, 1] = 4                              # an all-missing pattern
    w = rng.integers(1, 5, size=P).astype(np.int32)
    f = np.array([.1, .2, .3, .4])
    kappa = 2.5
    sm = SiteModel(HKY(kappa, f), mutationRate=0.7, gammaCategoryCount=gamma, shape=0.5,
                   proportionInvariant=pInv)
    rates, props = sm.category_rates_and_proportions()
    want = brute_force(tips, w, parent, height, rates, props, f, kappa, pInv)
    for scaling in (0, 1, 2):
        got = orc.tree_likelihood(tips, w, parent, height, rates, props, f, kappa,
                                  pInv=pInv, useScaling=scaling)
        assert got == pytest.approx(want, rel=1e-12)


def test_site_model_rates_average_to_one():
    for gamma, pInv in ((4, 0.0), (4, 0.3), (8, 0.1), (0, 0.25)):
        sm = SiteModel(HKY(), 1.0, gamma, 0.7, pInv)
        r, p = sm.category_rates_and_proportions()
        assert (r * p).sum() == pytest.approx(1.0, rel=1e-12)
        assert p.sum() == pytest.approx(1 - pInv)


def test_scaling_rescues_underflow_and_agrees_when_finite():
    m, N = mtt.synthetic_model(4)
    tb = mtt.simulate_trees(1, 1500, 4, m, N, seed=8)
    f = np.array([.25] * 4)
    tips = simulate_alignment(tb.parent, tb.height, 40, 2.0, f, 50.0, seed=1)   # saturated
    w = np.ones(40, np.int32)
    a = orc.tree_likelihood(tips, w, tb.parent, tb.height, [1.0], [1.0], f, 2.0,
                            branchRate=50.0, useScaling=0)
    b = orc.tree_likelihood(tips, w, tb.parent, tb.height, [1.0], [1.0], f, 2.0,
                            branchRate=50.0, useScaling=1)
    c = orc.tree_likelihood(tips, w, tb.parent, tb.height, [1.0], [1.0], f, 2.0,
                            branchRate=50.0, useScaling=2)
    assert a == -np.inf and np.isfinite(b) and c == b       # BEAST turns scaling on after -inf
    small = mtt.simulate_trees(1, 60, 4, m, N, seed=9)
    tips = simulate_alignment(small.parent, small.height, 40, 2.0, f, 0.05, seed=2)
    a = orc.tree_likelihood(tips, w, small.parent, small.height, [1.0], [1.0], f, 2.0,
                            branchRate=0.05, useScaling=0)
    b = orc.tree_likelihood(tips, w, small.parent, small.height, [1.0], [1.0], f, 2.0,
                            branchRate=0.05, useScaling=1)
    assert np.isfinite(a) and a == pytest.approx(b, rel=1e-13)


def test_alignment_pattern_compression():
    from multitypetree_b200.likelihood import Alignment
    seqs = {"a": "ACGTACGTNN--", "b": "ACGTACGARY-A", "c": "ACGAACGACCCC"}
    al = Alignment(seqs)
    assert al.siteCount == 12 and al.patternWeight.sum() == 12
    assert al.getPatternCount() < 12                       # ACG columns repeat
    rebuilt = al.patterns[:, al.patternIndex]
    want = np.stack([np.minimum(np.array([{"A": 0, "C": 1, "G": 2, "T": 3}.get(ch, 4)
                                          for ch in seqs[t]]), 4) for t in al.taxa])
    assert np.array_equal(rebuilt, want)
    sh = [al.shard(r, 3) for r in range(3)]
    assert sum(s.getPatternCount() for s in sh) == al.getPatternCount()
