"""Probe: typed trees of the untyped bench batch against the oracle, field by field."""
import sys
sys.path.insert(0, ".")
import numpy as np
import multitypetree_b200 as mtt
from multitypetree_b200.uniformization import UniformizationModels
from multitypetree_b200.untyped import untyped_sample_batch
from oracle import oracle as orc

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
d, n, P, W = 2, 100, 256, 1024
m, N = mtt.synthetic_model(d)
tb = mtt.simulate_trees(B, n, d, m, N, seed=6000)
trees = [tb.tree(b) for b in range(B)]
rng = np.random.default_rng(6000)
orders = np.concatenate([rng.permutation(n).astype(np.int32) for _ in range(B * P)])
u = rng.random(B * P * W)
uOff = np.arange(B * P + 1, dtype=np.int64) * W
ctx = mtt.Context(0)
um = UniformizationModels(m, ctx=ctx)
r = untyped_sample_batch(um, N, trees, P, orders, u, uOff, chgCap=256, ctx=ctx)
o = orc.UnifModel(m)
bad = 0
for q in range(min(B * P, 4096)):
    t = trees[q // P]
    ro = orc.untyped_particle_tree(o, N, t.parent, t.height, t.nodeType, orders[q * n:(q + 1) * n],
                                   u[uOff[q]:uOff[q + 1]])
    b, p_, c = q // P, q % P, ro["nChg"]
    g = dict(nChg=int(r["nChg"][b, p_]), st=int(r["status"][b, p_]))
    f = []
    if g["nChg"] != c:
        f.append("nChg %d vs %d" % (g["nChg"], c))
    if not np.array_equal(r["nodeType"][b][p_], ro["nodeType"].astype(np.uint8)):
        f.append("nodeType at %s" % np.nonzero(r["nodeType"][b][p_] != ro["nodeType"])[0][:5])
    k = min(c, g["nChg"], 256)
    for name in ("chgNode", "chgType", "chgTime"):
        x, y = r[name][b, p_, :k], ro[name][:k]
        if name == "chgTime":
            y = np.where(np.abs(x - y) <= 4 * np.spacing(y), x, y)     # last place of log()
        if not np.array_equal(x, y):
            i = int(np.nonzero(x != y)[0][0])
            f.append("%s[%d] %r vs %r (of %d)" % (name, i, x[i], y[i], k))
    if f or g["st"] != ro["status"]:
        bad += 1
        if bad <= 8:
            print(q, "status", g["st"], ro["status"], "uUsed", int(r["uUsed"][b, p_]), ro["uUsed"], "; ".join(f))
print("differing particles:", bad)
