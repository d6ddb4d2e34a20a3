"""Quick device-resident timing of the density kernel (development probe, not the bench)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import multitypetree_b200 as mtt
from multitypetree_b200 import _lib


def dev(a):
    t = torch.from_numpy(np.ascontiguousarray(a).ravel()).cuda()
    pad = torch.zeros(t.numel() + 64, dtype=t.dtype, device="cuda")
    pad[:t.numel()] = t
    return pad


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    d = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    m, N = mtt.synthetic_model(d)
    t0 = time.time()
    tb = mtt.simulate_trees(B, n, d, m, N, seed=2000)
    print("simulated %d trees in %.2fs; E mean %.1f max %d; %.1f MB" % (
        B, time.time() - t0, tb.n_events().mean(), tb.n_events().max(), tb.nbytes() / 1e6))
    ctx = mtt.Context(0)
    L = _lib.lib()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    arrs = [dev(x) for x in (m, N, tb.nodeOff, tb.height, tb.parent, tb.nodeType, tb.chgNode,
                             tb.chgBase, tb.chgTime, tb.chgType)]
    logP = torch.zeros(B, dtype=torch.float64, device="cuda")
    for hint, with_parent in ((int(tb.n_events().max()), False), (0, False),
                              (int(tb.n_events().max()), True)):
        L.mtt_sc_set_max_tree_events(ctx.handle, hint)
        def run():
            p = [a.data_ptr() for a in arrs]
            par = p[4] if with_parent else None
            ctx.check(L.mtt_sc_logp_owner_batch_dev(ctx.handle, B, d, 1, p[0], p[1], p[2], p[3], par,
                                              p[5], p[6], p[7], p[8], p[9], int(with_parent),
                                              logP.data_ptr(), None))
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        alg = 12 * tb.n_events().sum() + 8 * B
        print("hint=%d parent=%d: %.3f ms/batch  %.1f M trees/s  alg %.1f GB/s (%.3f of 6531)  actual-bytes %.1f GB/s"
              % (hint, with_parent, ms, B / ms / 1e3, alg / ms / 1e6, alg / ms / 1e6 / 6531.3,
                 tb.nbytes() / ms / 1e6))
    from oracle import oracle as orc
    k = min(B, 4096)
    want = orc.sc_logp_batch(tb.slice(0, k), m, N, variant=1, nThreads=8)
    got = logP[:k].cpu().numpy()
    print("max rel err vs oracle:", np.max(np.abs(got - want) / np.abs(want)))


if __name__ == "__main__":
    main()
