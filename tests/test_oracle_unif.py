"""CPU tests: the oracle for piece 3 (parity unpinned by the reference; checked against
scipy / numpy and against theory)."""
import numpy as np
import pytest
import scipy.linalg

from oracle import oracle as orc


def _model(d, seed=0, scale=1.0):
    rng = np.random.default_rng(seed)
    m = rng.uniform(0.05, 1.0, size=(d, d)) * scale
    np.fill_diagonal(m, 0.0)
    return m


@pytest.mark.parametrize("d", [2, 3, 4, 8, 10, 16])
def test_update_matrices(d):
    m = _model(d, d)
    for sym in (False, True):
        um = orc.UnifModel(m, sym)
        Q = (0.5 * (m + m.T) if sym else m).copy()
        np.fill_diagonal(Q, -Q.sum(1))
        np.testing.assert_allclose(um.Q, Q, rtol=1e-14, atol=1e-16)
        assert um.mu == pytest.approx(np.max(-np.diag(Q)), rel=1e-14)
        np.testing.assert_allclose(um.R, Q / um.mu + np.eye(d), rtol=1e-13, atol=1e-15)
        assert np.all(um.R >= -1e-15) and np.allclose(um.R.sum(1), 1.0)


@pytest.mark.parametrize("d", [2, 4, 10])
def test_rpow_matches_matrix_power_and_steady(d):
    um = orc.UnifModel(_model(d, 7 + d))
    for n in (0, 1, 2, 5, 17, 40):
        np.testing.assert_allclose(um.rpow(n), np.linalg.matrix_power(um.R, n), rtol=1e-12,
                                   atol=1e-15)
    um.rpow(3000)
    s = um.steady_n()
    assert s > 0 and (s - 1) % 10 == 0          # found at a multiple of 10 (:435)
    assert um.rpow_size() == s
    assert np.array_equal(um.rpow(s - 1), um.rpow(s + 123))   # n >= size returns the last (:422-426)
    assert np.array_equal(um.rpow(s - 1), um.rpow(s - 2))     # that is what "steady" means


@pytest.mark.parametrize("d,L", [(2, 0.0), (2, 0.3), (4, 1.7), (10, 0.01), (10, 25.0),
                                 (16, 3.0)])
def test_expm_matches_scipy(d, L):
    um = orc.UnifModel(_model(d, 3 * d))
    got = orc.expm(um.Q * L)
    want = scipy.linalg.expm(um.Q * L)
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(got.sum(1), 1.0, rtol=1e-11)


def test_branch_type_prob_formula():
    d = 3
    um = orc.UnifModel(_model(d, 11))
    t0, t1 = 0.5, 2.5
    ct, cty = [0.9, 1.4], [2, 1]
    st, lp = um.branch_type_prob(t0, t1, 0, ct, cty, 1)
    Q = um.Q
    want = (Q[0, 0] * 0.4 + np.log(Q[0, 2]) + Q[2, 2] * 0.5 + np.log(Q[2, 1]) + Q[1, 1] * 1.1
            - np.log(scipy.linalg.expm(Q * 2.0)[0, 1]))
    assert st == 0 and lp == pytest.approx(want, rel=1e-12)


def test_retype_consumes_uniforms_in_reference_order():
    d = 4
    um = orc.UnifModel(_model(d, 5))
    rng = np.random.default_rng(1)
    u = rng.random(4096)
    r = um.retype_branch(0.0, 3.0, 1, 2, u)
    assert r["status"] == 0
    assert r["uUsed"] == 1 + 2 * r["nVirt"]          # 1 count draw, nVirt times, nVirt FB draws
    assert r["nChg"] <= r["nVirt"]
    times = r["chgTime"]
    assert np.all(np.diff(times) >= 0) and (len(times) == 0 or (times[0] >= 0 and times[-1] <= 3))
    ty = [1] + list(r["chgType"])
    assert all(a != b for a, b in zip(ty[:-1], ty[1:])) and ty[-1] == 2
    # the path probability returned equals getBranchTypeProb of the produced path
    st, lp = um.branch_type_prob(0.0, 3.0, 1, r["chgTime"], r["chgType"], 2)
    assert lp == pytest.approx(r["logProb"], rel=1e-12)
    # too few uniforms is reported, not read past
    assert um.retype_branch(0.0, 3.0, 1, 2, u[:1])["status"] in (0, 5)


def test_retype_statistics_match_theory():
    """P(no change | a -> a over L) = exp(Q_aa L) / P_aa(L); mean number of changes by
    numerical integration of the conditioned jump intensity -- the conditioned-path checks of
    XML/badness_eradication/conditionedPathGen.R, done analytically."""
    d = 3
    um = orc.UnifModel(_model(d, 21))
    L, a, b = 1.3, 0, 0
    rng = np.random.default_rng(2)
    n, zero, tot = 20000, 0, 0
    for _ in range(n):
        r = um.retype_branch(0.0, L, a, b, rng.random(512))
        assert r["status"] == 0
        zero += r["nChg"] == 0
        tot += r["nChg"]
    P = scipy.linalg.expm(um.Q * L)
    p0 = np.exp(um.Q[a, a] * L) / P[a, b]
    assert abs(zero / n - p0) < 4 * np.sqrt(p0 * (1 - p0) / n)
    # E[#jumps | a,b] = (1/P_ab) * int_0^L sum_{i != j} P_ai(s) Q_ij P_jb(L-s) ds
    ss = np.linspace(0, L, 2001)
    Qoff = um.Q - np.diag(np.diag(um.Q))
    vals = [scipy.linalg.expm(um.Q * s)[a] @ Qoff @ scipy.linalg.expm(um.Q * (L - s))[:, b]
            for s in ss]
    mean = np.trapezoid(vals, ss) / P[a, b]
    assert abs(tot / n - mean) < 0.05


def test_no_valid_path():
    m = np.array([[0, 1.0, 0], [0, 0, 0.5], [0, 0, 0]])     # 0->1->2 only; state 2 absorbing
    # mu = max(-Q_ii) = 1; R is still stochastic
    um = orc.UnifModel(m)
    u = np.random.default_rng(0).random(256)
    assert um.retype_branch(0.0, 1.0, 2, 0, u)["status"] == 1     # Pba == 0
    r = um.retype_branch(0.0, 1.0, 0, 2, u)
    assert r["status"] == 0 and list(r["chgType"][-1:]) == [2]


def test_randomizer_stream_properties():
    """MersenneTwisterFast restatement: determinism, range, nextInt bounds, peek/skip."""
    from multitypetree_b200.uniformization import Randomizer
    a, b = Randomizer(42), Randomizer(42)
    xs = a.nextDoubles(1000)
    assert np.array_equal(xs, b.nextDoubles(1000)) and xs.min() >= 0 and xs.max() < 1
    assert abs(xs.mean() - 0.5) < 0.05
    pk = a.peekDoubles(10)
    assert np.array_equal(pk, a.nextDoubles(10))
    b.skipDoubles(10)
    assert a.nextDouble() == b.nextDouble()
    ints = [a.nextInt(7) for _ in range(2000)] + [a.nextInt(8) for _ in range(2000)]
    assert min(ints) == 0 and max(ints[:2000]) == 6 and max(ints[2000:]) == 7
    assert Randomizer(1).nextDouble() != Randomizer(2).nextDouble()


def test_randomizer_rejects_zero_seed():
    """A zero seed (BEAST: 'Non zero random seed required') would freeze the Knuth-seeded
    generator at 0 forever; seeds with zero low 32 bits are the same state."""
    import pytest
    from multitypetree_b200.uniformization import Randomizer
    for bad in (0, 1 << 32, 5 << 32):
        with pytest.raises(ValueError):
            Randomizer(bad)
    r = Randomizer(7)
    with pytest.raises(ValueError):
        r.setSeed(0)
    assert 0.0 < r.nextDouble() < 1.0
