import sys, time
sys.path.insert(0, ".")
import numpy as np
import multitypetree_b200 as mtt
from multitypetree_b200 import _lib
m, N = mtt.synthetic_model(2)
ctx = mtt.Context(0)
L = _lib.lib()
for B, n in ((1, 100), (64, 100), (64, 1000), (4096, 100)):
    tb = mtt.simulate_trees(B, n, 2, m, N, seed=1)
    host = [np.ascontiguousarray(x) for x in (m, N, tb.nodeOff, tb.height, tb.nodeType, tb.chgNode, tb.chgBase, tb.chgTime, tb.chgType)]
    out = np.empty(B)
    p = [_lib.ptr(a) for a in host]
    def call():
        ctx.check(L.mtt_sc_logp_owner_batch(ctx.handle, B, 2, 1, p[0], p[1], p[2], p[3], None, p[4], p[5], p[6], p[7], p[8], 0, _lib.ptr(out), None))
    for _ in range(20): call()
    t0 = time.perf_counter()
    for _ in range(200): call()
    print("B=%d n=%d: %.1f us per call" % (B, n, (time.perf_counter() - t0) / 200 * 1e6))

# the fixed-tips entry (what a BEAST plugin calls per step: tips installed once)
for B, n in ((1, 100), (64, 100)):
    tb = mtt.simulate_trees(B, n, 2, m, N, seed=1)
    mtt.set_tips(tb.height[:n], tb.nodeType[:n], ctx=ctx)
    mtt.sc_logp_tips_batch(tb, m, N, ctx=ctx)
    nN = 2 * n - 1
    intH = np.ascontiguousarray(tb.height.reshape(B, nN)[:, n:])
    intT = np.ascontiguousarray(tb.nodeType.reshape(B, nN)[:, n:])
    arrs = [np.ascontiguousarray(x) for x in (m, N, intH, intT, tb.chgNode, tb.chgBase, tb.chgTime, tb.chgType)]
    q = [_lib.ptr(a) for a in arrs]
    out = np.empty(B)
    def call_tips():
        ctx.check(L.mtt_sc_logp_tips_batch(ctx.handle, B, 2, 1, q[0], q[1], q[2], None, q[3], q[4], q[5], q[6], q[7], 0, _lib.ptr(out), None))
    for _ in range(20): call_tips()
    t0 = time.perf_counter()
    for _ in range(200): call_tips()
    print("tips entry B=%d n=%d: %.1f us per call" % (B, n, (time.perf_counter() - t0) / 200 * 1e6))
    assert np.array_equal(out, mtt.sc_logp_batch(tb, m, N, ctx=ctx))
