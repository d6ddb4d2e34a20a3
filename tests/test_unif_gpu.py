"""GPU parity tests for piece 3 (csrc/k3_unif.cu) through the C-ABI, against the oracle.

Bars (north_star): R^k, expm and sampled type paths bit-exact for the same uniform stream;
log-probabilities within 1e-9 relative (they pass through libm log/exp, <= 1 ulp each side)."""
import numpy as np
import pytest
import scipy.linalg

import multitypetree_b200 as mtt
from multitypetree_b200.operators import NodeRetype
from multitypetree_b200.uniformization import (Randomizer, UniformizationModels,
                                               ST_NEED_MORE_UNIFORMS, ST_NEEDS_POISSON_TAIL,
                                               ST_NO_VALID_PATH, ST_OK, ST_SKIPPED)
from oracle import oracle as orc
from tests.util import load_density_fixture, oracle_node_retype

pytestmark = pytest.mark.gpu


def _model(d, seed=0, scale=1.0):
    rng = np.random.default_rng(seed)
    m = rng.uniform(0.05, 1.0, size=(d, d)) * scale
    np.fill_diagonal(m, 0.0)
    return m


@pytest.mark.parametrize("d", [2, 3, 4, 8, 10, 16])
@pytest.mark.parametrize("sym", [False, True])
def test_update_matrices_bit_exact(ctx, d, sym):
    ms = np.stack([_model(d, 10 * d + k) for k in range(5)])
    um = UniformizationModels(ms, sym=sym, ctx=ctx)
    for k in range(5):
        o = orc.UnifModel(ms[k], sym)
        Q, R, mu = um.matrices(k)
        assert np.array_equal(Q, o.Q) and np.array_equal(R, o.R) and mu == o.mu


@pytest.mark.parametrize("d", [2, 4, 10, 16])
def test_rpow_bit_exact_and_steady(ctx, d):
    """SCMigrationModel.getRpowN / RpowSteadyN (:404-473): the chain R^i = R^(i-1) R."""
    ms = np.stack([_model(d, 100 + d), _model(d, 200 + d, scale=0.1)])
    um = UniformizationModels(ms, ctx=ctx)
    for k in range(2):
        o = orc.UnifModel(ms[k])
        pows, steady = um.rpow(k, 25)
        assert steady == -1 or steady <= 26
        for n in range(26):
            assert np.array_equal(pows[n], o.rpow(n)), (k, n)
        np.testing.assert_allclose(pows[17], np.linalg.matrix_power(o.R, 17), rtol=1e-12,
                                   atol=1e-15)
    # extend far enough to hit the steady state; table grows in place
    for k in range(2):
        o = orc.UnifModel(ms[k])
        o.rpow(4000)
        pows, steady = um.rpow(k, 4000)
        # (a 2x2 chain may flip its last bit forever and never become steady: -1 on both sides)
        assert steady == o.steady_n()
        probe = [0, 1, 26, 100, 1234, 4000]
        if steady > 0:
            probe += [steady - 2, steady - 1, steady, steady + 5]
        for n in probe:
            assert np.array_equal(pows[n], o.rpow(n)), (k, n)


@pytest.mark.parametrize("d", [2, 3, 4, 8, 10, 16])
def test_expm_bit_exact_vs_oracle_and_close_to_scipy(ctx, d):
    m = _model(d, 3 * d)
    um = UniformizationModels(m, ctx=ctx)
    o = orc.UnifModel(m)
    Ls = np.array([0.0, 1e-6, 0.013, 0.3, 0.77, 1.0, 2.5, 9.0, 40.0, 333.0])
    got = um.expm(Ls)
    for i, L in enumerate(Ls):
        want = orc.expm(o.Q * L)
        np.testing.assert_allclose(got[i], want, rtol=1e-13, atol=1e-300)
        np.testing.assert_allclose(got[i], scipy.linalg.expm(o.Q * L), rtol=1e-9, atol=1e-12)
    # bit-exactness away from power-of-two scaling boundaries
    same = sum(np.array_equal(got[i], orc.expm(o.Q * L)) for i, L in enumerate(Ls))
    assert same >= len(Ls) - 1


def test_branch_logp_matches_oracle(ctx):
    """getBranchTypeProb (:233-275) on every branch of the reference's 128-tip true tree."""
    tb, m, N, _ = load_density_fixture("true128")
    t = tb.tree(0)
    for sym in (False, True):
        um = UniformizationModels(m, sym=sym, ctx=ctx)
        o = orc.UnifModel(m, sym)
        bs = [b for b in range(t.getNodeCount()) if t.parent[b] >= 0]
        t0 = [t.height[b] for b in bs]
        t1 = [t.height[t.parent[b]] for b in bs]
        a = [t.getNodeType(b) for b in bs]
        bb = [t.getNodeType(int(t.parent[b])) for b in bs]
        off = np.cumsum([0] + [t.getChangeCount(b) for b in bs])
        ct = [x for b in bs for x in t.changeTimes[b]]
        cty = [x for b in bs for x in t.changeTypes[b]]
        lp, st = um.branch_logp(t0, t1, a, bb, off, ct, cty)
        want = np.array([o.branch_type_prob(t0[i], t1[i], a[i], t.changeTimes[b],
                                            t.changeTypes[b], bb[i])[1]
                         for i, b in enumerate(bs)])
        assert np.all(st == 0)
        np.testing.assert_allclose(lp, want, rtol=1e-9, atol=1e-12)


def _random_jobs(rng, nJobs, d, maxBr=3, Lscale=2.0):
    brOff = np.concatenate([[0], np.cumsum(rng.integers(1, maxBr + 1, nJobs))])
    nBr = int(brOff[-1])
    t0 = rng.uniform(0, 5, nBr)
    t1 = t0 + rng.exponential(Lscale, nBr)
    a = rng.integers(0, d, nBr)
    b = rng.integers(0, d, nBr)
    return brOff, t0, t1, a, b


@pytest.mark.parametrize("d,Lscale,sym", [(2, 1.0, False), (4, 2.0, False), (10, 2.0, False),
                                          (10, 8.0, True), (16, 0.5, False)])
def test_retype_paths_bit_exact(ctx, d, Lscale, sym):
    """retypeBranch (:116-225): same uniform stream in -> identical (nVirt, nChanges, types,
    times) out, logProb within 1e-9, uniforms consumed identical."""
    rng = np.random.default_rng(1000 + d)
    m = _model(d, 77 + d)
    um = UniformizationModels(m, sym=sym, ctx=ctx)
    o = orc.UnifModel(m, sym)
    nJobs, W = 400, 2048
    brOff, t0, t1, a, b = _random_jobs(rng, nJobs, d, Lscale=Lscale)
    u = rng.random(nJobs * W)
    uOff = np.arange(nJobs + 1) * W
    r = um.retype(brOff, t0, t1, a, b, u, uOff, outCap=256)
    nOK = 0
    for j in range(nJobs):
        pos = 0
        for br in range(brOff[j], brOff[j + 1]):
            w = o.retype_branch(t0[br], t1[br], int(a[br]), int(b[br]), u[j * W + pos:(j + 1) * W],
                                cap=256)
            assert r["status"][br] == w["status"], (j, br)
            if w["status"] != ST_OK:
                break
            nOK += 1
            pos += w["uUsed"]
            assert r["nVirt"][br] == w["nVirt"]
            n = w["nChg"]
            assert r["nChg"][br] == n
            assert np.array_equal(r["chgType"][br, :n], w["chgType"].astype(np.uint8))
            assert np.array_equal(r["chgTime"][br, :n], w["chgTime"])          # bit-exact
            assert r["logProb"][br] == pytest.approx(w["logProb"], rel=1e-9, abs=1e-12)
        else:
            assert r["uUsed"][j] == pos
    assert nOK > nJobs


def test_retype_many_models_one_per_chain(ctx):
    """config 5 shape: 64 chains, d = 10, each with its own migration model."""
    d, nCh = 10, 64
    rng = np.random.default_rng(5)
    ms = np.stack([_model(d, 900 + c, scale=rng.uniform(0.2, 2.0)) for c in range(nCh)])
    um = UniformizationModels(ms, ctx=ctx)
    os_ = [orc.UnifModel(ms[c]) for c in range(nCh)]
    brOff = np.arange(nCh + 1) * 3
    t0 = rng.uniform(0, 3, 3 * nCh)
    t1 = t0 + rng.exponential(3.0, 3 * nCh)
    a = rng.integers(0, d, 3 * nCh)
    b = rng.integers(0, d, 3 * nCh)
    model = np.repeat(np.arange(nCh), 3)
    W = 4096
    u = rng.random(nCh * W)
    r = um.retype(brOff, t0, t1, a, b, u, np.arange(nCh + 1) * W, outCap=512, model=model)
    for c in range(nCh):
        pos = 0
        for br in range(3 * c, 3 * c + 3):
            w = os_[c].retype_branch(t0[br], t1[br], int(a[br]), int(b[br]),
                                     u[c * W + pos:(c + 1) * W], cap=512)
            assert r["status"][br] == w["status"]
            if w["status"] != ST_OK:          # long branch: steady-state tail, job stops
                assert w["status"] == ST_NEEDS_POISSON_TAIL
                assert r["nVirt"][br] == w["nVirt"]
                pos = None
                break
            pos += w["uUsed"]
            n = w["nChg"]
            assert r["nChg"][br] == n and np.array_equal(r["chgTime"][br, :n], w["chgTime"])
            assert np.array_equal(r["chgType"][br, :n], w["chgType"].astype(np.uint8))
        if pos is not None:
            assert r["uUsed"][c] == pos


def test_retype_status_codes(ctx):
    m = np.array([[0, 1.0, 0], [0, 0, 0.5], [0, 0, 0]])     # 0 -> 1 -> 2 only
    um = UniformizationModels(m, ctx=ctx)
    u = np.random.default_rng(0).random(3 * 256)
    # job 0: impossible transition first, then a fine branch -> NO_VALID_PATH, SKIPPED
    # job 1: fine; job 2: starved of uniforms
    r = um.retype([0, 2, 3, 4], [0, 0, 0, 0], [1, 1, 1, 30], [2, 0, 0, 0], [0, 2, 2, 2], u,
                  [0, 256, 512, 514], outCap=64)
    assert list(r["status"][:3]) == [ST_NO_VALID_PATH, ST_SKIPPED, ST_OK]
    assert r["status"][3] == ST_NEED_MORE_UNIFORMS
    assert r["logProb"][0] == -np.inf and r["chgType"][2, r["nChg"][2] - 1] == 2
    o = orc.UnifModel(m)
    assert o.retype_branch(0, 1, 2, 0, u)["status"] == ST_NO_VALID_PATH


def test_retype_degenerate_branches(ctx):
    """Zero-length and very short / very long branches: L = 0 gives expm = I (jblas: j = 0),
    so equal end types yield an empty path and different end types NO_VALID_PATH."""
    d = 4
    m = _model(d, 31)
    um = UniformizationModels(m, ctx=ctx)
    o = orc.UnifModel(m)
    t0 = np.array([1.0, 1.0, 0.0, 0.0, 2.0, 0.0])
    t1 = np.array([1.0, 1.0, 1e-12, 1e-300, 2.0 + 1e-9, 60.0])
    a = np.array([2, 2, 1, 3, 0, 1])
    b = np.array([2, 0, 1, 3, 0, 2])
    W = 4096
    u = np.random.default_rng(4).random(len(t0) * W)
    r = um.retype(np.arange(len(t0) + 1), t0, t1, a, b, u, np.arange(len(t0) + 1) * W, outCap=512)
    for i in range(len(t0)):
        w = o.retype_branch(t0[i], t1[i], int(a[i]), int(b[i]), u[i * W:(i + 1) * W], cap=512)
        assert r["status"][i] == w["status"], i
        if w["status"] == ST_OK:
            n = w["nChg"]
            assert r["nChg"][i] == n and r["nVirt"][i] == w["nVirt"]
            assert np.array_equal(r["chgTime"][i, :n], w["chgTime"])
            assert r["logProb"][i] == pytest.approx(w["logProb"], rel=1e-9, abs=1e-12)
    assert r["status"][0] == ST_OK and r["nChg"][0] == 0
    assert r["status"][1] == ST_NO_VALID_PATH
    E = um.expm([0.0, 1e-300])
    assert np.array_equal(E[0], np.eye(d)) and np.allclose(E[1], np.eye(d), atol=1e-290)
    lp, st = um.branch_logp([1.0], [1.0], [2], [2], [0, 0], [], [])
    assert st[0] == 0 and lp[0] == 0.0


def test_retype_steady_state_tail_and_forced_count(ctx):
    """Huge mu*L: the count draw runs past the steady index -> NEEDS_POISSON_TAIL with the
    threshold in nVirt (:94-102); resubmitting with nVirtForced samples the path."""
    d = 3
    for seed in range(100):          # a model whose R^k chain does reach a bit-exact fixpoint
        m = _model(d, seed, scale=5.0)
        probe = orc.UnifModel(m)
        probe.rpow(3000)
        if probe.steady_n() > 0:
            break
    else:
        pytest.skip("no steady model found")
    um = UniformizationModels(m, ctx=ctx)
    o = orc.UnifModel(m)
    L = 400.0
    rng = np.random.default_rng(9)
    u = rng.random(8192)
    r = um.retype([0, 1], [0.0], [L], [0], [1], u, [0, 8192], outCap=16)
    w = o.retype_branch(0.0, L, 0, 1, u)
    assert r["status"][0] == w["status"] == ST_NEEDS_POISSON_TAIL
    assert r["nVirt"][0] == w["nVirt"] == o.steady_n()
    nv = 2200
    r = um.retype([0, 1], [0.0], [L], [0], [1], u[1:], [0, 8191], outCap=4096,
                  nVirtForced=[nv])
    w = o.retype_branch(0.0, L, 0, 1, u[1:], nVirtForced=nv, cap=4096)
    assert r["status"][0] == w["status"] == ST_OK and r["nVirt"][0] == nv
    n = w["nChg"]
    assert r["nChg"][0] == n and n > 100
    assert np.array_equal(r["chgTime"][0, :n], w["chgTime"])
    assert np.array_equal(r["chgType"][0, :n], w["chgType"].astype(np.uint8))
    assert r["logProb"][0] == pytest.approx(w["logProb"], rel=1e-9)


def test_retype_statistics_match_theory(ctx):
    """Same analytic checks as the oracle's CPU test, on 2^17 GPU-sampled paths."""
    d = 3
    m = _model(d, 21)
    um = UniformizationModels(m, ctx=ctx)
    Q = orc.UnifModel(m).Q
    L, n, W = 1.3, 1 << 17, 64
    rng = np.random.default_rng(3)
    u = rng.random(n * W)
    r = um.retype(np.arange(n + 1), np.zeros(n), np.full(n, L), np.zeros(n, int),
                  np.zeros(n, int), u, np.arange(n + 1) * W, outCap=32)
    ok = r["status"] == ST_OK
    assert ok.mean() > 0.999
    P = scipy.linalg.expm(Q * L)
    p0 = np.exp(Q[0, 0] * L) / P[0, 0]
    z = (r["nChg"][ok] == 0).mean()
    assert abs(z - p0) < 5 * np.sqrt(p0 * (1 - p0) / ok.sum())


def test_node_retype_operator_matches_oracle_stream(ctx):
    """operators/NodeRetype.java:31-63 as a whole: 200 successive proposals on the reference's
    64-tip tree, GPU mirror vs CPU oracle driven by the same Randomizer seed -> identical trees
    (types and change times bit for bit) and Hastings ratios within 1e-9."""
    tb, _, _, _ = load_density_fixture("true64")
    d = 2
    model = mtt.SCMigrationModel([0.3, 0.2], [5.0, 10.0], mtt.TypeSet("0", "1"), ctx=ctx)
    m, _ = model.dense()
    tg, to = tb.tree(0), tb.tree(0)
    op = NodeRetype(multiTypeTree=tg, migrationModel=model, randomizer=Randomizer(53), ctx=ctx)
    o = orc.UnifModel(m)
    rng_o = Randomizer(53)
    finite = 0
    for step in range(200):
        hr_g = op.proposal()
        hr_o = oracle_node_retype(to, o, d, rng_o)
        assert np.array_equal(tg.nodeType, to.nodeType), step
        assert tg.changeTimes == to.changeTimes and tg.changeTypes == to.changeTypes, step
        if np.isfinite(hr_o):
            finite += 1
            assert hr_g == pytest.approx(hr_o, rel=1e-9, abs=1e-10)
        else:
            assert hr_g == hr_o
        assert op.rng.nextDouble() == rng_o.nextDouble()      # streams stay in lock step
    assert finite > 150
    # the retyped tree is still a valid multi-type tree whose density the GPU evaluates
    assert tg.isValid(ctx)
    dens = mtt.StructuredCoalescentTreeDensity(model, tg, checkValidity=True, ctx=ctx)
    want = orc.sc_logp_batch(mtt.TreeBatch.from_trees([to]), *model.dense())[0]
    assert dens.calculateLogP() == pytest.approx(want, rel=1e-9)


def test_migration_model_device_accessors(ctx):
    """SCMigrationModel.getQ/getR/getMu/getRpowN mirror (device-backed)."""
    model = mtt.SCMigrationModel([2.0, 1.0], [5.0, 10.0], mtt.TypeSet("A", "B"), ctx=ctx)
    Q = model.getQ(False)
    assert np.array_equal(Q, [[-2.0, 2.0], [1.0, -1.0]]) and model.getMu(False) == 2.0
    assert np.array_equal(model.getR(False), [[0.0, 1.0], [0.5, 0.5]])
    np.testing.assert_allclose(model.getRpowN(5), np.linalg.matrix_power(model.getR(), 5),
                               rtol=1e-14)
    assert np.array_equal(model.getQ(True), [[-1.5, 1.5], [1.5, -1.5]])
    model.setRateMatrix([4.0, 1.0])
    assert model.getMu(False) == 4.0        # dirty -> matrices rebuilt, cached powers dropped


def test_retype_large_batch_chunked_path_equals_one_piece_path(ctx):
    """Batches of >= 16384 jobs go through the chunked three-stream path (upload, kernels and
    download of eight chunks overlap): it must return exactly what the one-piece path returns for
    the same jobs, and what the oracle returns."""
    d = 6
    rng = np.random.default_rng(4242)
    m = _model(d, 5)
    um = UniformizationModels(m, ctx=ctx)
    o = orc.UnifModel(m)
    nJobs, W = 20000, 64
    brOff, t0, t1, a, b = _random_jobs(rng, nJobs, d, Lscale=1.0)
    u = rng.random(nJobs * W)
    uOff = np.arange(nJobs + 1) * W
    whole = um.retype(brOff, t0, t1, a, b, u, uOff, outCap=32)
    whole = {k: np.array(v) for k, v in whole.items()}
    lo = 0
    for hi in (7000, 14000, nJobs):                     # three one-piece calls
        bs, be = brOff[lo], brOff[hi]
        part = um.retype(brOff[lo:hi + 1] - bs, t0[bs:be], t1[bs:be], a[bs:be], b[bs:be],
                         u[lo * W:hi * W], uOff[lo:hi + 1] - lo * W, outCap=32)
        for k in ("status", "nVirt", "nChg", "logProb"):
            assert np.array_equal(np.asarray(part[k]), whole[k][bs:be], equal_nan=True), (k, lo)
        live = np.arange(32)[None, :] < np.minimum(whole["nChg"][bs:be], 32)[:, None]
        for k in ("chgTime", "chgType"):                # (entries past nChg are not written)
            assert np.array_equal(np.asarray(part[k])[live], whole[k][bs:be][live]), (k, lo)
        assert np.array_equal(np.asarray(part["uUsed"]), whole["uUsed"][lo:hi])
        lo = hi
    for j in rng.integers(0, nJobs, 150):               # and against the oracle
        pos = 0
        for br in range(brOff[j], brOff[j + 1]):
            w = o.retype_branch(t0[br], t1[br], int(a[br]), int(b[br]), u[j * W + pos:(j + 1) * W],
                                cap=32)
            assert whole["status"][br] == w["status"]
            if w["status"] != ST_OK:
                break
            pos += w["uUsed"]
            n = w["nChg"]
            assert whole["nVirt"][br] == w["nVirt"] and whole["nChg"][br] == n
            assert np.array_equal(whole["chgTime"][br, :n], w["chgTime"])
