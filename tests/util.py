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


class OracleUnifModels:
    """TEST adapter: the batched interface of multitypetree_b200.uniformization
    .UniformizationModels (branch_logp / retype / matrices, same status protocol as
    mtt_unif_retype_batch) answered by the CPU oracle.  Lets the host logic of operators.py and
    mcmc.py run without a GPU in the ``not gpu`` tests; never used by the product."""

    def __init__(self, m, sym=False):
        from oracle import oracle as orc
        m = np.asarray(m, float)
        if m.ndim == 2:
            m = m[None]
        self.nModels, self.d = m.shape[0], m.shape[1]
        self.models = [orc.UnifModel(mm, sym) for mm in m]

    def matrices(self, model=0):
        um = self.models[model]
        return um.Q, um.R, um.mu

    def branch_logp(self, tBottom, tTop, typeBottom, typeTop, chgOff, chgTime, chgType,
                    model=None):
        n = len(tBottom)
        lp, st = np.empty(n), np.empty(n, np.int32)
        for i in range(n):
            um = self.models[0 if model is None else int(model[i])]
            lo, hi = int(chgOff[i]), int(chgOff[i + 1])
            st[i], lp[i] = um.branch_type_prob(float(tBottom[i]), float(tTop[i]),
                                               int(typeBottom[i]), chgTime[lo:hi],
                                               chgType[lo:hi], int(typeTop[i]))
        return lp, st

    def retype(self, brOff, tBottom, tTop, typeBottom, typeTop, uniforms, uOff, outCap=64,
               model=None, nVirtForced=None):
        nJobs, nBr = len(brOff) - 1, len(tBottom)
        uniforms = np.asarray(uniforms, float)
        out = dict(nChg=np.zeros(nBr, np.int32), chgTime=np.zeros((nBr, outCap)),
                   chgType=np.zeros((nBr, outCap), np.uint8), logProb=np.full(nBr, -np.inf),
                   status=np.full(nBr, 7, np.int32), nVirt=np.full(nBr, -1, np.int32),
                   uUsed=np.zeros(nJobs, np.int64))
        for j in range(nJobs):
            u = uniforms[int(uOff[j]):int(uOff[j + 1])]
            pos = 0
            for br in range(int(brOff[j]), int(brOff[j + 1])):
                um = self.models[0 if model is None else int(model[br])]
                forced = -1 if nVirtForced is None else int(nVirtForced[br])
                r = um.retype_branch(float(tBottom[br]), float(tTop[br]), int(typeBottom[br]),
                                     int(typeTop[br]), u[pos:], nVirtForced=forced, cap=outCap)
                out["status"][br] = r["status"]
                out["nVirt"][br] = r["nVirt"]
                pos += int(r["uUsed"])
                if r["status"] != 0:
                    break
                k = min(r["nChg"], outCap)
                out["nChg"][br] = r["nChg"]
                out["chgTime"][br, :k] = r["chgTime"][:k]
                out["chgType"][br, :k] = r["chgType"][:k]
                out["logProb"][br] = r["logProb"]
            out["uUsed"][j] = pos
        return out
