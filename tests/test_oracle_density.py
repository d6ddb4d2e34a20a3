"""CPU tests: the oracle for piece 1 against the reference's golden vector and against
independent formulas.  (No GPU.)"""
import math

import numpy as np
import pytest

import multitypetree_b200 as mtt
from oracle import oracle as orc
from tests.util import load_density_fixture


def test_reference_kat_all_variants():
    """SCLikelihoodTest.java:43-73: -16.52831 within 1e-5 (exactly -11.7 - 3 ln 5)."""
    tb, m, N, expected = load_density_fixture("kat")
    for variant in (orc.FAITHFUL, orc.SORTED, orc.SUFFSTAT):
        got = orc.sc_logp_batch(tb, m, N, variant=variant)[0]
        assert abs(got - expected) < 1e-5
        assert abs(got - (-11.7 - 3 * math.log(5))) < 1e-12


def test_kat_from_newick_matches_fixture():
    kat = ("(((A[&state=1]:0.25)[&state=0]:0.25,B[&state=0]:0.5)[&state=0]:1.5,"
           "(C[&state=0]:1.0,D[&state=0]:1.0)[&state=0]:1.0)[&state=0]:0.0;")
    t = mtt.MultiTypeTreeFromNewick(kat, "state")
    assert t.getLeafNodeCount() == 4 and t.getTotalNumberOfChanges() == 1
    assert t.getNodeType(0) == 1 and t.getChangeType(0, 0) == 0
    assert t.getChangeTime(0, 0) == pytest.approx(0.25)
    tb, *_ = load_density_fixture("kat")
    tb2 = mtt.TreeBatch.from_trees([t])
    for f in ("height", "parent", "nodeType", "chgCount", "chgTime", "chgType"):
        assert np.array_equal(getattr(tb, f), getattr(tb2, f))


def test_kat_mpmath():
    """Independent high-precision evaluation of the KAT intervals (SURVEY 8c derivation)."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    N = [mp.mpf(5), mp.mpf(10)]
    mrate = {(0, 1): mp.mpf(2), (1, 0): mp.mpf(1)}
    intervals = [(mp.mpf("0.25"), (3, 1)), (mp.mpf("0.25"), (4, 0)), (mp.mpf("0.5"), (3, 0)),
                 (mp.mpf(1), (2, 0))]
    lp = mp.mpf(0)
    for dt, k in intervals:
        lam = sum(k[c] * (k[c] - 1) / (2 * N[c]) for c in range(2))
        lam += k[0] * mrate[(0, 1)] + k[1] * mrate[(1, 0)]
        lp -= dt * lam
    lp += mp.log(mrate[(1, 0)]) + 3 * mp.log(1 / N[0])
    tb, m, Nn, _ = load_density_fixture("kat")
    got = orc.sc_logp_batch(tb, m, Nn)[0]
    assert abs(got - float(lp)) < 1e-13


@pytest.mark.parametrize("name", ["true128", "true64", "ewing3"])
def test_fixtures_variants_agree(name):
    tb, m, N, expected = load_density_fixture(name)
    a = orc.sc_logp_batch(tb, m, N, variant=orc.FAITHFUL, checkValidity=True)[0]
    b = orc.sc_logp_batch(tb, m, N, variant=orc.SORTED)[0]
    c = orc.sc_logp_batch(tb, m, N, variant=orc.SUFFSTAT)[0]
    assert a == pytest.approx(expected, rel=1e-14)
    assert b == pytest.approx(a, rel=1e-13)
    assert c == pytest.approx(a, rel=1e-12)


@pytest.mark.parametrize("n,d", [(5, 2), (100, 2), (64, 3), (300, 4), (1000, 8)])
def test_simulated_trees_variants_agree(n, d):
    m, N = mtt.synthetic_model(d)
    tb = mtt.simulate_trees(16, n, d, m, N, seed=100 * n + d, nThreads=2)
    a = orc.sc_logp_batch(tb, m, N, variant=orc.FAITHFUL, checkValidity=True)
    b = orc.sc_logp_batch(tb, m, N, variant=orc.SORTED)
    c = orc.sc_logp_batch(tb, m, N, variant=orc.SUFFSTAT)
    assert np.all(np.isfinite(a))
    np.testing.assert_allclose(b, a, rtol=1e-13)
    np.testing.assert_allclose(c, a, rtol=1e-11)


def test_rate_encodings():
    """SCMigrationModel.java:119-133, 324-345: the three rateMatrix encodings, scale factors,
    flags and forward-rate mode resolve to the same dense matrix in the host mirror and in
    the oracle."""
    d = 3
    full = np.array([[0, 1, 2], [3, 0, 4], [5, 6, 0]], float)
    offdiag = [1, 2, 3, 4, 5, 6]
    sym = [1.5, 2.5, 3.5]            # (1,0), (2,0), (2,1)
    pops = [2.0, 3.0, 4.0]
    ts = mtt.TypeSet("a", "b", "c")
    m_sq, _ = mtt.SCMigrationModel(full.ravel(), pops, ts).dense()
    m_od, _ = mtt.SCMigrationModel(offdiag, pops, ts).dense()
    assert np.array_equal(m_sq, full) and np.array_equal(m_od, full)
    m_sy, _ = mtt.SCMigrationModel(sym, pops, ts).dense()
    assert m_sy[1, 0] == m_sy[0, 1] == 1.5 and m_sy[2, 0] == 2.5 and m_sy[1, 2] == 3.5
    for kw in (dict(), dict(rateScale=0.5, popScale=2.0),
               dict(rateFlags=[1, 0, 1, 1, 0, 1]), dict(useForward=True)):
        mo, No = orc.resolve_dense(d, offdiag, pops, **kw)
        mm = mtt.SCMigrationModel(
            offdiag, pops, ts, rateMatrixScaleFactor=kw.get("rateScale"),
            popSizesScaleFactor=kw.get("popScale"), rateMatrixFlags=kw.get("rateFlags"),
            useForwardMigrationRateMatrix=kw.get("useForward", False))
        mh, Nh = mm.dense()
        np.testing.assert_allclose(mh, mo, rtol=0, atol=0)
        np.testing.assert_allclose(Nh, No, rtol=0, atol=0)
    with pytest.raises(ValueError):
        mtt.SCMigrationModel([1, 2, 3, 4], pops, ts)


def test_invalid_tree_detected():
    tb, m, N, _ = load_density_fixture("kat")
    t = tb.tree(0)
    bad_type = t.copy()
    bad_type.nodeType[0] = 0          # leaf A now type 0 but its branch still ends in type 0->ok
    bad_type.changeTypes[0] = [1]     # final type 1 != parent type 0
    bad_time = t.copy()
    bad_time.changeTimes[0] = [0.75]  # change above the parent's height 0.5
    for bad in (bad_type, bad_time):
        b = mtt.TreeBatch.from_trees([bad])
        assert orc.sc_logp_batch(b, m, N, checkValidity=True)[0] == -np.inf
    assert np.isfinite(orc.sc_logp_batch(tb, m, N, checkValidity=True)[0])
