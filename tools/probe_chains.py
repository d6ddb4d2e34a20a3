"""Development probe: where the host time of mtt_peel_chains_update goes (cfg3 shape)."""
import sys
import time
sys.path.insert(0, ".")
import ctypes as C
import numpy as np
import multitypetree_b200 as mtt
from multitypetree_b200 import _lib
from multitypetree_b200.likelihood import (Alignment, HKY, SiteModel, TreeLikelihood,
                                           simulate_alignment)

K, n, d = 64, 1000, 4
m, N = mtt.synthetic_model(d)
tb = mtt.simulate_trees(K, n, d, m, N, seed=3000, nThreads=8)
nN = 2 * n - 1
par, hgt = tb.parent.reshape(K, nN).copy(), tb.height.reshape(K, nN).copy()
f = np.array([.25] * 4)
clock = 0.2 / float(tb.height.max())
al = Alignment(states=simulate_alignment(par[0], hgt[0], 10000, 2.0, f, clock, seed=3000))
ctx = mtt.Context(0)
tl = TreeLikelihood(data=al, siteModel=SiteModel(HKY(2.0, f), mutationRate=clock), ctx=ctx)
ch = tl.chains(K)
ch.set_timing(True)
ch.update(par, hgt)
ch.accept()
rng = np.random.default_rng(1)
dirty0 = np.zeros((K, nN), np.uint8)


def timeit(fn, reps=200):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e6


print("update, nothing dirty (mask given): %.1f us" % timeit(lambda: ch.update(par, hgt, dirty=dirty0)))
print("update, nothing dirty (full compare): %.1f us" % timeit(lambda: ch.update(par, hgt)))
L = _lib.lib()
out = np.empty(K)
pp, hp, dp, op = _lib.ptr(par), _lib.ptr(hgt), _lib.ptr(dirty0), _lib.ptr(out)
print("raw C call, nothing dirty (mask): %.1f us" % timeit(
    lambda: L.mtt_peel_chains_update(ch._s, pp, hp, None, None, None, dp, op)))
# one height change per chain (leaf-most internal nodes vs random)
kids = [[np.nonzero(par[k] == v)[0] for v in range(nN)] for k in range(K)]


def proposal():
    ph = hgt.copy()
    dirty = np.zeros((K, nN), np.uint8)
    for k in range(K):
        v = int(rng.integers(n, nN))
        lo = max(hgt[k, c] for c in kids[k][v])
        hi = hgt[k, par[k, v]] if par[k, v] >= 0 else lo + 1
        ph[k, v] = lo + (hi - lo) * 0.5
        dirty[k, v] = 1
    return ph, dirty


for timing in (True, False):
  ch.set_timing(timing)
  tot = dev = ent = 0.0
  reps = 100
  for _ in range(reps):
    ph, dirty = proposal()
    hp2, dp2 = _lib.ptr(ph), _lib.ptr(dirty)
    t0 = time.perf_counter()
    L.mtt_peel_chains_update(ch._s, pp, hp2, None, None, None, dp2, op)
    tot += time.perf_counter() - t0
    dev += ch.last_device_ms
    ent += ch.last_entries
  print("raw C call, one move per chain (events %s): %.1f us total, %.1f us device, %.1f entries/chain"
        % (timing, tot / reps * 1e6, dev / reps * 1e3, ent / reps / K))
acc = np.ones(K, np.uint8)
ap = _lib.ptr(acc)


def step():
    L.mtt_peel_chains_update(ch._s, pp, hp2, None, None, None, dp2, op)
    L.mtt_peel_chains_accept(ch._s, ap)


print("repeat the same accepted proposal (no entries) + accept: %.1f us" % timeit(step))
ch.close()
tl.close()
