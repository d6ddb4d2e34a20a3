"""Development probe: device-resident timing of the density kernel on config-2 trees
(one launch shape, hint = max events), optionally with another build of the library
(MTT_B200_LIB).  Not the bench."""
import sys
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
    tb = mtt.simulate_trees(B, n, d, m, N, seed=2000)
    ctx = mtt.Context(0)
    L = _lib.lib()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    arrs = [dev(x) for x in (m, N, tb.nodeOff, tb.height, tb.parent, tb.nodeType, tb.chgNode,
                             tb.chgBase, tb.chgTime, tb.chgType)]
    logP = torch.zeros(B, dtype=torch.float64, device="cuda")
    L.mtt_sc_set_max_tree_events(ctx.handle, int(tb.n_events().max()))
    p = [a.data_ptr() for a in arrs]

    def run():
        ctx.check(L.mtt_sc_logp_owner_batch_dev(ctx.handle, B, d, 1, p[0], p[1], p[2], p[3], None,
                                                p[5], p[6], p[7], p[8], p[9], 0,
                                                logP.data_ptr(), None))
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    alg = 12 * tb.n_events().sum() + 8 * B
    from oracle import oracle as orc
    k = min(B, 2048)
    want = orc.sc_logp_batch(tb.slice(0, k), m, N, variant=1, nThreads=8)
    got = logP[:k].cpu().numpy()
    err = np.max(np.abs(got - want) / np.abs(want))
    print("%s: %.3f ms  %.1f M trees/s  frac %.3f  err %.1e" % (
        _lib.LIB_PATH.split("/")[-1], best, B / best / 1e3, alg / best / 1e6 / 6531.3, err))


if __name__ == "__main__":
    main()
