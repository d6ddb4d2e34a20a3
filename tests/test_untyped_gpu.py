"""StructuredCoalescentUntypedTreeDensity on the device (csrc/k4_untyped.cu)
against the CPU oracle (oracle/orc_untyped.c) particle by particle for the same leaf orders and
uniform windows, plus the oracle itself against the exact two-leaf marginal density."""
import numpy as np
import pytest
import scipy.linalg as sl

import multitypetree_b200 as mtt
from multitypetree_b200.uniformization import Randomizer, UniformizationModels
from multitypetree_b200.untyped import (StructuredCoalescentUntypedTreeDensity, shuffled,
                                        untyped_logp_batch)
from oracle import oracle as orc


def _exact_two_leaf(m, N, a, b, T):
    """log density of a two-leaf untyped tree with coalescence time T: the pair process
    (type of lineage 1, type of lineage 2) killed at rate 1/N_c while both sit in deme c."""
    d = len(N)
    Q = np.array(m, float)
    for i in range(d):
        Q[i, i] = -(Q[i].sum() - Q[i, i])
    G = np.kron(Q, np.eye(d)) + np.kron(np.eye(d), Q)
    for c in range(d):
        G[c * d + c, c * d + c] -= 1.0 / N[c]
    E = sl.expm(G * T)
    return np.log(sum(E[a * d + b, c * d + c] / N[c] for c in range(d)))


def test_oracle_particle_estimator_converges_to_exact_two_leaf_density():
    m = np.array([[0, 0.3], [0.2, 0]])
    N = np.array([5.0, 10.0])
    um = orc.UnifModel(m)
    parent = np.array([2, 2, -1], np.int32)
    rng = np.random.default_rng(7)
    for a, b, T in ((0, 0, 3.0), (0, 1, 3.0), (1, 0, 8.0)):
        height = np.array([0.0, 0.0, T])
        ws = []
        for _ in range(20000):
            r = orc.untyped_particle(um, N, parent, height, [a, b, 0], rng.permutation(2),
                                     rng.random(512))
            assert r["status"] == 0
            ws.append(r["logW"])
        ws = np.array(ws)
        w = np.exp(ws - ws.max())
        se = w.std() / w.mean() / np.sqrt(len(w))
        assert abs(orc.untyped_combine(ws) - _exact_two_leaf(m, N, a, b, T)) < 5 * se + 1e-12


def _inputs(trees, P, win, seed):
    rng = np.random.default_rng(seed)
    orders = [rng.permutation(t.getLeafNodeCount()).astype(np.int32) for t in trees
              for _ in range(P)]
    u = rng.random(len(trees) * P * win)
    uOff = np.arange(len(trees) * P + 1, dtype=np.int64) * win
    return orders, u, uOff


@pytest.mark.gpu
@pytest.mark.parametrize("d,n,scale", [(2, 8, 1.0), (2, 40, 1.0), (3, 25, 0.5), (4, 60, 2.0),
                                       (10, 30, 1.0), (16, 12, 1.0),
                                       (2, 10, 30.0)])     # many migrations: the stage grows
def test_particles_match_oracle(ctx, d, n, scale):
    """Every particle: same log weight (1e-9), same number of uniforms consumed -- i.e. the
    same sampled typing -- and the combined estimate, for random trees and models."""
    m, N = mtt.synthetic_model(d)
    m = m * scale
    trees = [mtt.simulate_trees(3, n, d, m, N, seed=90 + d).tree(b) for b in range(3)]
    P, win = 12, 4096
    orders, u, uOff = _inputs(trees, P, win, 5 + d)
    um = UniformizationModels(m, ctx=ctx)
    r = untyped_logp_batch(um, N, trees, P, np.concatenate(orders), u, uOff, ctx=ctx)
    o = orc.UnifModel(m)
    for b, t in enumerate(trees):
        ws = []
        for p in range(P):
            q = b * P + p
            ro = orc.untyped_particle(o, N, t.parent, t.height, t.nodeType, orders[q],
                                      u[uOff[q]:uOff[q + 1]])
            assert ro["status"] == r["status"][b, p] == 0
            assert ro["uUsed"] == r["uUsed"][b, p], (b, p)
            assert r["logW"][b, p] == pytest.approx(ro["logW"], rel=1e-9, abs=1e-9), (b, p)
            ws.append(ro["logW"])
        assert r["logP"][b] == pytest.approx(orc.untyped_combine(ws), rel=1e-9)


@pytest.mark.gpu
def test_per_tree_models_and_size_limit(ctx):
    """model[] selects the migration model per tree; a tree beyond the documented node limit is
    refused loudly (MTT_ERR_TOO_LARGE), never truncated."""
    d = 3
    m0, N = mtt.synthetic_model(d)
    ms = np.stack([m0, 0.3 * m0])
    Ns = np.stack([N, 2.0 * N])
    trees = [mtt.simulate_trees(2, 15, d, m0, N, seed=40).tree(b) for b in range(2)]
    P, win = 5, 2048
    orders, u, uOff = _inputs(trees, P, win, 9)
    um = UniformizationModels(ms, ctx=ctx)
    r = untyped_logp_batch(um, Ns, trees, P, np.concatenate(orders), u, uOff, model=[1, 0],
                           ctx=ctx)
    for b, t in enumerate(trees):
        mdl = [1, 0][b]
        o = orc.UnifModel(ms[mdl])
        for p in range(P):
            q = b * P + p
            ro = orc.untyped_particle(o, Ns[mdl], t.parent, t.height, t.nodeType, orders[q],
                                      u[uOff[q]:uOff[q + 1]])
            assert r["logW"][b, p] == pytest.approx(ro["logW"], rel=1e-9, abs=1e-9)
    big = mtt.simulate_trees(1, 600, d, m0, N, seed=41).tree(0)          # 1199 nodes > 1024
    orders, u, uOff = _inputs([big], 1, 64, 2)
    with pytest.raises(mtt.MttError, match="nodes"):
        untyped_logp_batch(um, Ns, [big], 1, np.concatenate(orders), u, uOff, ctx=ctx)


@pytest.mark.gpu
def test_window_too_short_is_reported(ctx):
    m, N = mtt.synthetic_model(2)
    t = mtt.simulate_trees(1, 30, 2, m, N, seed=4).tree(0)
    orders, u, uOff = _inputs([t], 4, 8, 1)
    um = UniformizationModels(m, ctx=ctx)
    r = untyped_logp_batch(um, N, [t], 4, np.concatenate(orders), u, uOff, ctx=ctx)
    assert np.all(r["status"] == 5)
    o = orc.UnifModel(m)
    ro = orc.untyped_particle(o, N, t.parent, t.height, t.nodeType, orders[0], u[:8])
    assert ro["status"] == 5


@pytest.mark.gpu
def test_disconnected_model_gives_reference_infinities(ctx):
    """No path between the types (Pba == 0): the Java subtracts -Infinity from the particle
    weight (:136, :311-313); the device reproduces the resulting +inf / NaN pattern."""
    m = np.array([[0.0, 0.0], [0.0, 0.0]]) + np.array([[0, 1e-300], [1e-300, 0]]) * 0
    m = np.array([[0.0, 0.5], [0.0, 0.0]])          # 1 -> 0 impossible backward in time
    N = np.array([1.0, 1.0])
    t = mtt.MultiTypeTreeFromNewick("((a[&t=0]:1,b[&t=1]:1)[&t=0]:1,c[&t=0]:2)[&t=0]:0;", "t")
    orders, u, uOff = _inputs([t], 6, 256, 3)
    um = UniformizationModels(m, ctx=ctx)
    r = untyped_logp_batch(um, N, [t], 6, np.concatenate(orders), u, uOff, ctx=ctx)
    o = orc.UnifModel(m)
    for p in range(6):
        ro = orc.untyped_particle(o, N, t.parent, t.height, t.nodeType, orders[p],
                                  u[uOff[p]:uOff[p + 1]])
        a, b = r["logW"][0, p], ro["logW"]
        assert (np.isnan(a) and np.isnan(b)) or a == b or a == pytest.approx(b, rel=1e-9)


@pytest.mark.gpu
def test_estimator_on_device_converges_to_exact_two_leaf_density(ctx):
    m = np.array([[0, 0.3], [0.2, 0]])
    N = np.array([5.0, 10.0])
    model = mtt.SCMigrationModel([0.3, 0.2], N, mtt.TypeSet("0", "1"), ctx=ctx)
    t = mtt.MultiTypeTreeFromNewick("(a[&t=0]:3,b[&t=1]:3)[&t=0]:0;", "t")
    dens = StructuredCoalescentUntypedTreeDensity(model, t, nParticles=50000,
                                                  randomizer=Randomizer(11), ctx=ctx)
    est = dens.calculateLogP()
    assert est == pytest.approx(_exact_two_leaf(m, N, 0, 1, 3.0), abs=0.02)
    assert dens.isStochastic() and dens.requiresRecalculation()


def test_shuffled_is_a_permutation():
    rng = Randomizer(3)
    for n in (1, 2, 7, 100):
        assert sorted(shuffled(rng, n)) == list(range(n))


def _py_particle(um, N, parent, height, leafType, order, u):
    """Pure-Python restatement of one particle (StructuredCoalescentUntypedTreeDensity.java
    :121-189, :216-316, :452-502) on top of the oracle's branch sampler: an independent check of
    orc_untyped.c's typing walk, event list, stable sort and density scan."""
    d = um.d
    Q = um.Q
    n = len(parent)
    nLeaves = (n + 1) // 2
    nodeType = [int(leafType[i]) if i < nLeaves else -1 for i in range(n)]
    events = []          # (time, seq, kind, type, dest)
    pos, w = 0, 0.0

    def add(kind, ty, dest, t):
        events.append((float(t), len(events), kind, ty, dest))

    for li, leaf in enumerate(order):
        leaf = int(leaf)
        if li == 0:
            time, ty, nxt = float(height[leaf]), nodeType[leaf], int(parent[leaf])
            add("S", ty, 0, time)
            lp = 0.0
            while True:
                aTot = 0.0
                for c in range(d):
                    if c != ty:
                        aTot += Q[ty, c]
                newTime = time + (-1.0 * np.log(1 - u[pos]) / aTot)
                pos += 1
                while nxt >= 0 and height[nxt] < newTime:
                    nodeType[nxt] = ty
                    add("C", ty, 0, height[nxt])
                    lp += -aTot * (height[nxt] - time)
                    time = float(height[nxt])
                    nxt = int(parent[nxt])
                if nxt < 0:
                    break
                lp += -aTot * (newTime - time)
                time = newTime
                uu = u[pos] * aTot
                pos += 1
                new = None
                for c in range(d):
                    if c == ty:
                        continue
                    uu -= Q[ty, c]
                    if uu < 0.0:
                        new = c
                        break
                assert new is not None
                add("M", ty, new, time)
                lp += np.log(Q[ty, new])
                ty = new
            w -= lp
            continue
        add("S", nodeType[leaf], 0, height[leaf])
        anc = int(parent[leaf])
        while nodeType[anc] < 0:
            anc = int(parent[anc])
        r = um.retype_branch(float(height[leaf]), float(height[anc]), nodeType[leaf],
                             nodeType[anc], u[pos:], cap=1024)
        assert r["status"] == 0
        pos += r["uUsed"]
        prevType, prevNode = nodeType[leaf], leaf
        for k in range(r["nChg"]):
            add("M", prevType, int(r["chgType"][k]), r["chgTime"][k])
            while height[prevNode] < r["chgTime"][k]:
                if prevNode >= nLeaves:
                    nodeType[prevNode] = prevType
                    add("C", prevType, 0, height[prevNode])
                prevNode = int(parent[prevNode])
            prevType = int(r["chgType"][k])
        while height[prevNode] < height[anc]:
            if prevNode >= nLeaves:
                nodeType[prevNode] = prevType
                add("C", prevType, 0, height[prevNode])
            prevNode = int(parent[prevNode])
        w -= r["logProb"]
    events.sort(key=lambda e: (e[0], e[1]))
    cnt = [1 if events[0][3] == c else 0 for c in range(d)]
    for i in range(1, len(events)):
        t, _, kind, ty, dest = events[i]
        dt = t - events[i - 1][0]
        if dt > 0:
            lam = 0.0
            for c in range(d):
                k = cnt[c]
                lam += k * (k - 1) / (2.0 * N[c])
                for cp in range(d):
                    if cp != c:
                        lam += k * Q[c, cp]
            w += -dt * lam
        if kind == "C":
            w += np.log(1.0 / N[ty])
            cnt[ty] -= 1
        elif kind == "M":
            w += np.log(Q[ty, dest])
            cnt[ty] -= 1
            cnt[dest] += 1
        else:
            cnt[ty] += 1
    return w, pos, nodeType, len(events)


@pytest.mark.parametrize("d,n", [(2, 6), (3, 12), (5, 20)])
def test_oracle_particle_matches_python_restatement(d, n):
    m, N = mtt.synthetic_model(d)
    um = orc.UnifModel(m)
    rng = np.random.default_rng(d)
    for b in range(3):
        t = mtt.simulate_trees(3, n, d, m, N, seed=500 + d).tree(b)
        for _ in range(6):
            order = rng.permutation(n)
            u = rng.random(2048)
            want = orc.untyped_particle(um, N, t.parent, t.height, t.nodeType, order, u)
            got = _py_particle(um, N, t.parent, t.height, t.nodeType, order, u)
            assert want["status"] == 0
            assert got[1] == want["uUsed"] and got[3] == want["nEvents"]
            assert list(got[2]) == list(want["nodeType"])
            assert got[0] == pytest.approx(want["logW"], rel=1e-12, abs=1e-12)


@pytest.mark.gpu
def test_large_call_chunked_upload_equals_small_calls(ctx):
    """Calls with >= 4096 particles upload their uniforms in chunks on a second stream and start
    each chunk's particles as it lands: same log weights and uniform counts as the same
    particles evaluated tree by tree in small calls, and as the oracle on a sample."""
    d, n = 3, 12
    m, N = mtt.synthetic_model(d)
    T, P, win = 20, 256, 512                       # 5120 particles
    trees = [mtt.simulate_trees(T, n, d, m, N, seed=77).tree(b) for b in range(T)]
    orders, u, uOff = _inputs(trees, P, win, 31)
    um = UniformizationModels(m, ctx=ctx)
    big = untyped_logp_batch(um, N, trees, P, np.concatenate(orders), u, uOff, ctx=ctx)
    assert np.all(big["status"] == 0)
    for b in range(T):                             # one tree per call: 256 particles, one piece
        lo, hi = b * P, (b + 1) * P
        small = untyped_logp_batch(um, N, trees[b:b + 1], P, np.concatenate(orders[lo:hi]),
                                   u[uOff[lo]:uOff[hi]], uOff[lo:hi + 1] - uOff[lo], ctx=ctx)
        assert np.array_equal(small["logW"][0], big["logW"][b])
        assert np.array_equal(small["uUsed"][0], big["uUsed"][b])
        assert small["logP"][0] == big["logP"][b]
    o = orc.UnifModel(m)
    rng = np.random.default_rng(3)
    for q in rng.integers(0, T * P, 40):
        b, p = divmod(int(q), P)
        t = trees[b]
        ro = orc.untyped_particle(o, N, t.parent, t.height, t.nodeType, orders[q],
                                  u[uOff[q]:uOff[q + 1]])
        assert ro["status"] == 0 and ro["uUsed"] == big["uUsed"][b, p]
        assert big["logW"][b, p] == pytest.approx(ro["logW"], rel=1e-9, abs=1e-9)


# ------------------------------------------------------------------ the typings themselves
def _check_typed_tree(t, d):
    """MultiTypeTree.isValid (evolution/tree/MultiTypeTree.java:396-429) in plain Python: every
    node typed, change times inside the branch and increasing, no change into the type it
    leaves, the type at the top of every branch = the parent's type."""
    n = t.getNodeCount()
    for v in range(n):
        assert 0 <= t.nodeType[v] < d
        p = int(t.parent[v])
        if p < 0:
            assert t.getChangeCount(v) == 0
            continue
        prev_t, prev_ty = t.height[v], int(t.nodeType[v])
        for k in range(t.getChangeCount(v)):
            ct, cy = t.getChangeTime(v, k), t.getChangeType(v, k)
            assert prev_t <= ct <= t.height[p]
            assert cy != prev_ty and 0 <= cy < d
            prev_t, prev_ty = ct, cy
        assert prev_ty == int(t.nodeType[p]), v


def _root_path(t, leaf):
    path, v = [], int(leaf)
    while v >= 0:
        path.append(v)
        v = int(t.parent[v])
    return np.array(path)


def test_oracle_typed_tree_is_a_valid_typing_of_the_untyped_tree():
    from multitypetree_b200.untyped import typed_tree_from_sample
    for d, n in ((2, 20), (4, 35), (7, 12)):
        m, N = mtt.synthetic_model(d)
        t = mtt.simulate_trees(1, n, d, m, N, seed=300 + d).tree(0)
        o = orc.UnifModel(m)
        rng = np.random.default_rng(d)
        for _ in range(10):
            order = rng.permutation(n).astype(np.int32)
            u = rng.random(8192)
            r = orc.untyped_particle_tree(o, N, t.parent, t.height, t.nodeType, order, u)
            r0 = orc.untyped_particle(o, N, t.parent, t.height, t.nodeType, order, u)
            assert r["status"] == 0 and r["uUsed"] == r0["uUsed"] and r["logW"] == r0["logW"]
            assert np.array_equal(r["nodeType"], r0["nodeType"])
            assert np.array_equal(r["nodeType"][:n], t.nodeType[:n])      # leaves keep their types
            ty = typed_tree_from_sample(t, r["nodeType"], r["chgNode"], r["chgType"], r["chgTime"])
            _check_typed_tree(ty, d)


@pytest.mark.gpu
@pytest.mark.parametrize("d,n,scale", [(2, 8, 1.0), (2, 100, 1.0), (3, 25, 0.5), (4, 60, 2.0),
                                       (10, 30, 1.0), (16, 12, 1.0), (2, 10, 30.0)])
def test_sampled_typings_match_oracle_bit_for_bit(ctx, d, n, scale):
    """mtt_untyped_sample_batch: node types, change owners, change types and change times of
    every particle identical to the restatement's for the same leaf order and uniform window
    (the unconditioned first lineage's times to the last place of log());
    weights as in the density entry; the returned trees are valid and the typed density of each
    (mtt_sc_logp_batch, the k1 kernel) agrees with the weight's density part."""
    from multitypetree_b200.untyped import typed_tree_from_sample, untyped_sample_batch
    m, N = mtt.synthetic_model(d)
    m = m * scale
    trees = [mtt.simulate_trees(3, n, d, m, N, seed=190 + d).tree(b) for b in range(3)]
    P, win = 6, 8192
    orders, u, uOff = _inputs(trees, P, win, 15 + d)
    um = UniformizationModels(m, ctx=ctx)
    r = untyped_sample_batch(um, N, trees, P, np.concatenate(orders), u, uOff, ctx=ctx)
    rl = untyped_logp_batch(um, N, trees, P, np.concatenate(orders), u, uOff, ctx=ctx)
    assert np.array_equal(r["logW"], rl["logW"]) and np.array_equal(r["uUsed"], rl["uUsed"])
    o = orc.UnifModel(m)
    typed = []
    for b, t in enumerate(trees):
        for p in range(P):
            q = b * P + p
            ro = orc.untyped_particle_tree(o, N, t.parent, t.height, t.nodeType, orders[q],
                                           u[uOff[q]:uOff[q + 1]])
            assert ro["status"] == r["status"][b, p] == 0
            c = ro["nChg"]
            assert r["nChg"][b, p] == c
            assert np.array_equal(r["nodeType"][b][p], ro["nodeType"].astype(np.uint8))
            assert np.array_equal(r["chgNode"][b, p, :c], ro["chgNode"])
            assert np.array_equal(r["chgType"][b, p, :c], ro["chgType"].astype(np.uint8))
            # times: bit-identical along every conditioned path (u * L + t0); the first
            # lineage's waiting times go through log() (Randomizer.nextExponential), where the
            # device's log and the host's may differ in the last place
            first = _root_path(t, orders[q][0])
            onFirst = np.isin(ro["chgNode"], first)
            gt = r["chgTime"][b, p, :c]
            assert np.array_equal(gt[~onFirst], ro["chgTime"][~onFirst])
            assert np.all(np.abs(gt - ro["chgTime"]) <= 4 * np.spacing(ro["chgTime"]))
            ty = typed_tree_from_sample(t, r["nodeType"][b][p], r["chgNode"][b, p, :c],
                                        r["chgType"][b, p, :c], r["chgTime"][b, p, :c])
            _check_typed_tree(ty, d)
            typed.append(ty)
    tb = mtt.TreeBatch.from_trees(typed)
    lp = mtt.sc_logp_batch(tb, m, N, checkValidity=True, ctx=ctx)
    want = orc.sc_logp_batch(tb, m, N, variant=orc.FAITHFUL)
    assert np.all(np.isfinite(lp)) and np.allclose(lp, want, rtol=1e-9, atol=0)


@pytest.mark.gpu
def test_sample_room_and_failures_are_reported(ctx):
    """nChg beyond chgCap is reported (never written past the room); a colouring without a valid
    path returns that lineage's status and unreached nodes as 255."""
    from multitypetree_b200.untyped import untyped_sample_batch
    m, N = mtt.synthetic_model(2)
    m = 20.0 * m
    t = mtt.simulate_trees(1, 10, 2, m, N, seed=4).tree(0)
    orders, u, uOff = _inputs([t], 3, 8192, 1)
    um = UniformizationModels(m, ctx=ctx)
    full = untyped_sample_batch(um, N, [t], 3, np.concatenate(orders), u, uOff, ctx=ctx)
    small = untyped_sample_batch(um, N, [t], 3, np.concatenate(orders), u, uOff, chgCap=4, ctx=ctx)
    assert np.array_equal(small["nChg"], full["nChg"]) and full["nChg"].min() > 4
    assert np.array_equal(small["chgTime"][0, :, :4], full["chgTime"][0, :, :4])
    m1 = np.array([[0.0, 0.5], [0.0, 0.0]])          # 1 -> 0 impossible backward in time
    t1 = mtt.MultiTypeTreeFromNewick("((a[&t=0]:1,b[&t=1]:1)[&t=0]:1,c[&t=0]:2)[&t=0]:0;", "t")
    um1 = UniformizationModels(m1, ctx=ctx)
    orders, u, uOff = _inputs([t1], 8, 256, 3)
    r = untyped_sample_batch(um1, np.array([1.0, 1.0]), [t1], 8, np.concatenate(orders), u, uOff,
                             ctx=ctx)
    o = orc.UnifModel(m1)
    for p in range(8):
        ro = orc.untyped_particle_tree(o, [1.0, 1.0], t1.parent, t1.height, t1.nodeType,
                                       orders[p], u[uOff[p]:uOff[p + 1]])
        assert ro["status"] == r["status"][0, p]
        got = r["nodeType"][0][p].astype(np.int32)
        got[got == 255] = -1
        assert np.array_equal(got, ro["nodeType"])
    assert np.any(r["status"] == 1)


@pytest.mark.gpu
def test_tree_initialisers(ctx):
    """MultiTypeTreeFromUntypedNewick and StructuredCoalescentUntypedTree through the device:
    valid typed trees with the Newick's topology, heights and leaf types; the same seed gives
    the same tree."""
    ts = mtt.TypeSet("A", "B", "C")
    model = mtt.SCMigrationModel([0.3, 0.1, 0.2, 0.4, 0.2, 0.1], [5.0, 3.0, 4.0], ts, ctx=ctx)
    nwk = "(((a:1.0,b:1.0):0.5,c:1.5):1.0,(d:0.7,e:0.2):1.3);"
    trait = dict(a="A", b="B", c="C", d="A", e="C")
    t1 = mtt.MultiTypeTreeFromUntypedNewick(nwk, model, trait, randomizer=Randomizer(11), ctx=ctx)
    t2 = mtt.MultiTypeTreeFromUntypedNewick(nwk, model, trait, randomizer=Randomizer(11), ctx=ctx)
    assert t1.getFlattenedNewick() == t2.getFlattenedNewick()
    assert t1.getLeafNodeCount() == 5 and list(t1.taxa) == list("abcde")
    assert [int(x) for x in t1.nodeType[:5]] == [0, 1, 2, 0, 2]
    assert t1.height[t1.getRoot()] == pytest.approx(2.5) and np.all(t1.height[:5] == 0.0)
    t3 = mtt.MultiTypeTreeFromUntypedNewick(nwk, model, trait, randomizer=Randomizer(11), ctx=ctx,
                                            adjustTipHeights=False)
    assert t3.height[:5] == pytest.approx([0.0, 0.0, 0.0, 0.5, 1.0])
    _check_typed_tree(t3, 3)
    _check_typed_tree(t1, 3)
    assert t1.isValid(ctx=ctx)
    back = mtt.MultiTypeTreeFromNewick(t1.getFlattenedNewick(), "type")
    assert back.getTotalNumberOfChanges() == t1.getTotalNumberOfChanges()
    with pytest.raises(ValueError, match="Trait set"):
        mtt.MultiTypeTreeFromUntypedNewick(nwk, model, None, ctx=ctx)
    # many trees in one call per round
    m, N = model.dense()
    many = [mtt.simulate_trees(16, 20, 3, m, N, seed=77).tree(b) for b in range(16)]
    for ty in mtt.colour_untyped_trees(model, many, Randomizer(5), ctx=ctx):
        _check_typed_tree(ty, 3)
    u = mtt.StructuredCoalescentUntypedTree(model, [0, 1, 2, 0, 1, 2, 0, 1], seed=9, ctx=ctx)
    assert u.getLeafNodeCount() == 8 and u.getTotalNumberOfChanges() == 0
    assert np.all(u.height[:8] == 0.0)
    for v in range(u.getNodeCount() - 1):
        assert u.height[u.parent[v]] > u.height[v]
    _check_typed_tree(mtt.colour_untyped_trees(model, [u], Randomizer(3), ctx=ctx)[0], 3)


def test_oracle_typed_trees_weighted_root_type_matches_exact_posterior():
    """The typing that comes back IS the one the weight belongs to: over many particles of a
    two-leaf tree the importance-weighted frequency of the root's type equals the exact posterior
    P(root type c | T) ~ expm(G T)[(a, b), (c, c)] / N_c of the killed pair process."""
    m = np.array([[0, 0.3], [0.2, 0]])
    N = np.array([5.0, 10.0])
    d = 2
    Q = np.array(m, float)
    for i in range(d):
        Q[i, i] = -Q[i].sum()
    G = np.kron(Q, np.eye(d)) + np.kron(np.eye(d), Q)
    for c in range(d):
        G[c * d + c, c * d + c] -= 1.0 / N[c]
    a, b, T = 0, 1, 4.0
    E = sl.expm(G * T)
    post = np.array([E[a * d + b, c * d + c] / N[c] for c in range(d)])
    post /= post.sum()
    um = orc.UnifModel(m)
    parent = np.array([2, 2, -1], np.int32)
    height = np.array([0.0, 0.0, T])
    rng = np.random.default_rng(11)
    ws, roots = [], []
    for _ in range(20000):
        r = orc.untyped_particle_tree(um, N, parent, height, [a, b, 0], rng.permutation(2),
                                      rng.random(512))
        assert r["status"] == 0
        # the changes lead from each leaf's type to the root's
        for leaf in (0, 1):
            ty = [a, b][leaf]
            for k in np.nonzero(r["chgNode"] == leaf)[0]:
                assert r["chgType"][k] != ty
                ty = r["chgType"][k]
            assert ty == r["nodeType"][2]
        ws.append(r["logW"])
        roots.append(r["nodeType"][2])
    w = np.exp(np.array(ws) - max(ws))
    roots = np.array(roots)
    est = np.array([w[roots == c].sum() for c in range(d)]) / w.sum()
    ess = w.sum() ** 2 / (w ** 2).sum()
    assert np.all(np.abs(est - post) < 5 * np.sqrt(post * (1 - post) / ess) + 1e-3), (est, post)


@pytest.mark.gpu
def test_sampled_typings_of_trees_of_different_sizes(ctx):
    """Ragged batch: trees of 8, 20 and 13 tips in one call, three particles each."""
    from multitypetree_b200.untyped import typed_tree_from_sample, untyped_sample_batch
    d = 3
    m, N = mtt.synthetic_model(d)
    trees = [mtt.simulate_trees(1, n, d, m, N, seed=500 + n).tree(0) for n in (8, 20, 13)]
    P, win = 3, 4096
    orders, u, uOff = _inputs(trees, P, win, 21)
    um = UniformizationModels(m, ctx=ctx)
    r = untyped_sample_batch(um, N, trees, P, np.concatenate(orders), u, uOff, ctx=ctx)
    o = orc.UnifModel(m)
    for b, t in enumerate(trees):
        assert r["nodeType"][b].shape == (P, t.getNodeCount())
        for p in range(P):
            q = b * P + p
            ro = orc.untyped_particle_tree(o, N, t.parent, t.height, t.nodeType, orders[q],
                                           u[uOff[q]:uOff[q + 1]])
            c = ro["nChg"]
            assert ro["status"] == r["status"][b, p] == 0 and r["nChg"][b, p] == c
            assert np.array_equal(r["nodeType"][b][p], ro["nodeType"].astype(np.uint8))
            assert np.array_equal(r["chgNode"][b, p, :c], ro["chgNode"])
            assert np.array_equal(r["chgType"][b, p, :c], ro["chgType"].astype(np.uint8))
            assert np.all(np.abs(r["chgTime"][b, p, :c] - ro["chgTime"])
                          <= 4 * np.spacing(ro["chgTime"]))
            _check_typed_tree(typed_tree_from_sample(t, r["nodeType"][b][p], r["chgNode"][b, p, :c],
                                                     r["chgType"][b, p, :c],
                                                     r["chgTime"][b, p, :c]), d)


def test_initialiser_host_logic_with_the_oracle_standing_in_for_the_device(monkeypatch):
    """MultiTypeTreeFromUntypedNewick's host side on CPU: Newick and trait handling,
    adjustTipHeights, the retry after a colouring without a valid path -- with the device call
    replaced (in this test only) by the oracle's particle."""
    import multitypetree_b200.untyped as un

    m = np.array([[0.0, 0.4, 0.1], [0.3, 0.0, 0.2], [0.2, 0.5, 0.0]])
    N = np.array([5.0, 3.0, 4.0])
    calls = []

    class _Dev:
        ctx = None

        def matrices(self, k):
            o = orc.UnifModel(m)
            return o.Q, o.R, o.mu

    class _Model:
        typeSet = mtt.TypeSet("A", "B", "C")

        def _ensure_device(self, sym):
            return _Dev()

        def dense(self):
            return m, N

    def fake_sample(um, N_, trees, P, order, u, uOff, chgCap=None, model=None, ctx=None):
        o = orc.UnifModel(m)
        B = len(trees)
        out = dict(nodeType=[], nChg=np.zeros((B, 1), np.int32), status=np.zeros((B, 1), np.int32),
                   chgNode=np.zeros((B, 1, 512), np.int32), chgType=np.zeros((B, 1, 512), np.uint8),
                   chgTime=np.zeros((B, 1, 512)))
        pos = 0
        for b, t in enumerate(trees):
            n = t.getLeafNodeCount()
            r = orc.untyped_particle_tree(o, N_, t.parent, t.height, t.nodeType,
                                          order[pos:pos + n], u[uOff[b]:uOff[b + 1]])
            pos += n
            st = r["status"]
            if len(calls) == 0 and b == 0:
                st = 1                      # first attempt: pretend there was no valid path
            out["status"][b, 0] = st
            out["nChg"][b, 0] = r["nChg"]
            out["nodeType"].append(r["nodeType"].astype(np.uint8)[None, :])
            c = r["nChg"]
            out["chgNode"][b, 0, :c] = r["chgNode"]
            out["chgType"][b, 0, :c] = r["chgType"]
            out["chgTime"][b, 0, :c] = r["chgTime"]
        calls.append(B)
        return out

    monkeypatch.setattr(un, "untyped_sample_batch", fake_sample)
    nwk = "(((a:1.0,b:1.0):0.5,(c:1.5)x:0.0):1.0,(d:0.7,e:0.2):1.3);"      # one single-child node
    trait = dict(a="A", b="B", c="C", d="A", e="C")
    t = un.MultiTypeTreeFromUntypedNewick(nwk, _Model(), trait, randomizer=Randomizer(3))
    assert calls == [1, 1]                                   # retried once
    assert t.getNodeCount() == 9 and list(t.taxa) == list("abcde")
    assert [int(x) for x in t.nodeType[:5]] == [0, 1, 2, 0, 2]
    assert np.all(t.height[:5] == 0.0) and t.height[8] == pytest.approx(2.5)
    _check_typed_tree(t, 3)
    t2 = un.MultiTypeTreeFromUntypedNewick(nwk, _Model(), trait, randomizer=Randomizer(3),
                                           adjustTipHeights=False)
    assert t2.height[:5] == pytest.approx([0.0, 0.0, 0.0, 0.5, 1.0])
    _check_typed_tree(t2, 3)
    back = mtt.MultiTypeTreeFromNewick(t2.getFlattenedNewick(), "type")
    assert back.getTotalNumberOfChanges() == t2.getTotalNumberOfChanges()
    assert np.allclose(np.sort(back.height), np.sort(t2.height), atol=1e-12)
    with pytest.raises(ValueError, match="Trait set"):
        un.MultiTypeTreeFromUntypedNewick(nwk, _Model(), None)
    with pytest.raises(KeyError):
        un.MultiTypeTreeFromUntypedNewick(nwk, _Model(), dict(a="A"))
