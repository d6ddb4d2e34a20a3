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
