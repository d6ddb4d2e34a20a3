"""CPU tests of the tree-state bridge: flat Newick with single-child nodes
<-> flattened tree SoA (evolution/tree/MultiTypeTree.java:444-526 getFlattenedTree,
:538-665 initFromFlatTree), and the batch container."""
import numpy as np
import pytest

import multitypetree_b200 as mtt
from oracle import oracle as orc
from tests.util import load_density_fixture


def _same_tree(a, b, tol=1e-12):
    """Equal up to node renumbering: compare through a canonical signature."""
    def sig(t):
        out = []
        for v in range(t.getNodeCount()):
            out.append((round(float(t.height[v]), 9), int(t.nodeType[v]),
                        tuple(round(x, 9) for x in t.changeTimes[v]), tuple(t.changeTypes[v]),
                        round(float(t.height[t.parent[v]]), 9) if t.parent[v] >= 0 else -1.0))
        return sorted(out)
    return sig(a) == sig(b)


@pytest.mark.parametrize("name,label", [("kat", "state"), ("true128", "location"),
                                        ("true64", "location"), ("ewing3", "deme")])
def test_flat_newick_round_trip(name, label):
    tb, m, N, expected = load_density_fixture(name)
    t = tb.tree(0)
    nw = t.getFlattenedNewick(label)
    t2 = mtt.MultiTypeTreeFromNewick(nw, label)
    assert t2.getLeafNodeCount() == t.getLeafNodeCount()
    assert t2.getTotalNumberOfChanges() == t.getTotalNumberOfChanges()
    assert _same_tree(t, t2)
    got = orc.sc_logp_batch(mtt.TreeBatch.from_trees([t2]), m, N, checkValidity=True)[0]
    assert got == pytest.approx(expected, rel=1e-5 if name == "kat" else 1e-12)


def test_round_trip_simulated_and_batch_slices():
    m, N = mtt.synthetic_model(3)
    tb = mtt.simulate_trees(12, 40, 3, m, N, seed=5, nThreads=1)
    want = orc.sc_logp_batch(tb, m, N)
    trees = [mtt.MultiTypeTreeFromNewick(tb.tree(b).getFlattenedNewick("type"), "type")
             for b in range(tb.B)]
    tb2 = mtt.TreeBatch.from_trees(trees)
    np.testing.assert_allclose(orc.sc_logp_batch(tb2, m, N), want, rtol=1e-12)
    part = tb.slice(3, 9)
    assert part.B == 6 and part.nodeOff[0] == 0 and part.chgBase[0] == 0
    np.testing.assert_allclose(orc.sc_logp_batch(part, m, N), want[3:9], rtol=0)
    assert tb.nbytes() == sum(tb.slice(b, b + 1).nbytes() - 16 for b in range(tb.B)) + 16


def test_newick_parser_metadata_forms():
    """Quoted values, extra attributes, metadata before or after the branch length."""
    a = mtt.MultiTypeTreeFromNewick(
        '((x[&type="X",location="1",time=3.0]:1.0)[&location="0",reaction="Migration"]:1.0,'
        'y[&location="0"]:2.0)[&location="0"]:0.0;', "location")
    b = mtt.MultiTypeTreeFromNewick(
        "((x:1.0[&location=1]):1.0[&location=0],y:2.0[&location=0])[&location=0];", "location")
    for t in (a, b):
        assert t.getLeafNodeCount() == 2 and t.getNodeType(0) == 1
        assert t.changeTimes[0] == [1.0] and t.changeTypes[0] == [0]
    ts = mtt.TypeSet("HongKong", "NewZealand")
    c = mtt.MultiTypeTreeFromNewick(
        "(a[&loc=NewZealand]:1,b[&loc=HongKong]:1)[&loc=HongKong]:0;", "loc", typeSet=ts)
    assert list(c.nodeType) == [1, 0, 0]


def test_owner_list_layouts_agree():
    """chgNode (owner layout of mtt_sc_logp_owner_batch): the simulator's native export, the
    list derived from per-node counts, and slices of either are the same thing."""
    m, N = mtt.synthetic_model(3)
    tb = mtt.simulate_trees(200, 40, 3, m, N, seed=12)
    native = tb.chgNode.copy()
    tb.set_owner_list(None)
    derived = tb.chgNode
    assert np.array_equal(native, derived) and derived.dtype == np.uint16
    assert len(derived) == len(tb.chgTime) == int(tb.chgBase[-1])
    # every change's owner carries that many changes, in node order
    for b in (0, 57, 199):
        n0, c0, c1 = int(tb.nodeOff[b]), int(tb.chgBase[b]), int(tb.chgBase[b + 1])
        own = derived[c0:c1].astype(int)
        assert np.all(np.diff(own) >= 0)
        assert np.array_equal(np.bincount(own, minlength=79), tb.chgCount[n0:n0 + 79])
    part = tb.slice(50, 120)
    assert np.array_equal(part.chgNode, derived[int(tb.chgBase[50]):int(tb.chgBase[120])])
    rebuilt = mtt.TreeBatch.from_trees([tb.tree(b) for b in range(5)])
    assert np.array_equal(rebuilt.chgNode, derived[:int(tb.chgBase[5])])


def test_nexus_tree_log_round_trip(tmp_path):
    """Loggable init / log / close (MultiTypeTree.java:795-831): the NEXUS tree log BEAST writes
    can be read back tree by tree (translated leaf numbers are Nr + 1)."""
    m, N = mtt.synthetic_model(2)
    tb = mtt.simulate_trees(3, 12, 2, m, N, seed=9, nThreads=1)
    trees = [tb.tree(i) for i in range(3)]
    path = tmp_path / "run.trees"
    with open(path, "w") as f:
        trees[0].init(f, "location")
        for i, t in enumerate(trees):
            t._logLabel = "location"
            t.log(1000 * i, f)
        trees[0].close(f)
    txt = open(path).read()
    assert txt.startswith("#NEXUS") and "Dimensions ntax=12;" in txt and "tree STATE_2000 = " in txt
    back = mtt.read_nexus_trees(str(path))
    assert len(back) == 3
    for nw, t in zip(back, trees):
        t2 = mtt.MultiTypeTreeFromNewick(nw, "location")
        assert _same_tree(t, t2)
        assert sorted(t2.taxa) == sorted(str(i + 1) for i in range(12))
