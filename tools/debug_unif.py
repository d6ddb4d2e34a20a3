import sys
import numpy as np
sys.path.insert(0, ".")
import multitypetree_b200 as mtt
from multitypetree_b200.uniformization import UniformizationModels
from oracle import oracle as orc
def _model(d, seed=0, scale=1.0):
    rng = np.random.default_rng(seed)
    m = rng.uniform(0.05, 1.0, size=(d, d)) * scale
    np.fill_diagonal(m, 0.0)
    return m
ctx = mtt.Context(0)
m = _model(3, 4, scale=5.0)
um = UniformizationModels(m, ctx=ctx)
o = orc.UnifModel(m)
Q, R, mu = um.matrices(0)
print("Q eq", np.array_equal(Q, o.Q), "R eq", np.array_equal(R, o.R), mu, o.mu)
pows, steady = um.rpow(0, 300)
o.rpow(300)
print("steady gpu", steady, "oracle", o.steady_n(), "size", o.rpow_size())
for n in range(301):
    if not np.array_equal(pows[n], o.rpow(n)):
        print("first mismatch at", n, pows[n] - o.rpow(n)); break
else:
    print("all 301 powers equal")
u = np.random.default_rng(9).random(8192)
r = um.retype([0, 1], [0.0], [400.0], [0], [1], u, [0, 8192], outCap=16)
w = o.retype_branch(0.0, 400.0, 0, 1, u)
print("gpu", r["status"], r["nVirt"], "oracle", w["status"], w["nVirt"], "steady now", um.rpow(0, 10)[1], o.steady_n())
