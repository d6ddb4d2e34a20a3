"""Shared helpers for the tests."""
import os

import numpy as np

import multitypetree_b200 as mtt

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_FIELDS = ("nodeOff", "height", "parent", "nodeType", "chgCount", "chgBase", "chgTime", "chgType")


def load_density_fixture(name):
    z = np.load(os.path.join(GOLDEN, "density_fixtures.npz"))
    tb = mtt.TreeBatch(*[z["%s_%s" % (name, f)] for f in _FIELDS])
    return tb, z[name + "_m"], z[name + "_N"], float(z[name + "_expected"])


def concat_batches(batches):
    trees = []
    for b in batches:
        trees.extend(b.tree(i) for i in range(b.B))
    return mtt.TreeBatch.from_trees(trees)


def rel_err(a, b):
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def oracle_node_retype(tree, um, d, rng):
    """operators/NodeRetype.java:31-63 restated on top of the CPU oracle's retypeBranch /
    getBranchTypeProb, consuming `rng` (a multitypetree_b200 Randomizer, host-side MT stream)
    in the reference's order.  Mutates `tree`; returns the log Hastings ratio."""
    t = tree
    node = t.getLeafNodeCount() + rng.nextInt(t.getInternalNodeCount())
    kids = t.getChildren(node)

    def old_prob(b):
        p = int(t.parent[b])
        st, lp = um.branch_type_prob(float(t.height[b]), float(t.height[p]), t.getNodeType(b),
                                     t.changeTimes[b], t.changeTypes[b], t.getNodeType(p))
        return lp

    logHR = 0.0
    if t.parent[node] >= 0:
        logHR += old_prob(node)
    logHR += old_prob(kids[0]) + old_prob(kids[1])
    t.setNodeType(node, rng.nextInt(d))

    def retype(b):
        p = int(t.parent[b])
        window = 256
        while True:
            u = rng.peekDoubles(window)
            r = um.retype_branch(float(t.height[b]), float(t.height[p]), t.getNodeType(b),
                                 t.getNodeType(p), u, cap=window)
            if r["status"] == 5:        # need more uniforms
                window *= 4
                continue
            break
        rng.skipDoubles(r["uUsed"])
        if r["status"] == 1:
            raise LookupError("no valid path")
        if r["status"] != 0:
            return float("-inf")
        t.clearChanges(b)
        for k in range(r["nChg"]):
            t.addChange(b, int(r["chgType"][k]), float(r["chgTime"][k]))
        return r["logProb"]

    try:
        if t.parent[node] >= 0:
            logHR -= retype(node)
        logHR -= retype(kids[0]) + retype(kids[1])
    except LookupError:
        return float("-inf")
    return logHR
