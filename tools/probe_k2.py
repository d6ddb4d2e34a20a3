"""Device-resident timing of the peeling kernels (development probe)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import multitypetree_b200 as mtt
from multitypetree_b200 import _lib
from multitypetree_b200.likelihood import Alignment, HKY, SiteModel, TreeLikelihood, simulate_alignment


def dev(a):
    t = torch.from_numpy(np.ascontiguousarray(a).ravel()).cuda()
    pad = torch.zeros(t.numel() + 64, dtype=t.dtype, device="cuda")
    pad[:t.numel()] = t
    return pad


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    sites = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    gamma = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    d = 4
    m, N = mtt.synthetic_model(d)
    tb = mtt.simulate_trees(B, n, d, m, N, seed=3000)
    nN = 2 * n - 1
    par, hgt = tb.parent.reshape(B, nN), tb.height.reshape(B, nN)
    f = np.array([.25] * 4)
    clock = 0.2 / float(hgt[0].max())
    al = Alignment(states=simulate_alignment(par[0], hgt[0], sites, 2.0, f, clock, seed=3000))
    P = al.getPatternCount()
    ctx = mtt.Context(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    L = _lib.lib()
    tl = TreeLikelihood(data=al, siteModel=SiteModel(HKY(2.0, f), mutationRate=clock,
                                                     gammaCategoryCount=gamma), ctx=ctx)
    dpar, dhgt = dev(par), dev(hgt)
    out = torch.zeros(B, dtype=torch.float64, device="cuda")

    def run():
        ctx.check(L.mtt_peel_logl_batch_dev(tl.handle, B, dpar.data_ptr(), dhgt.data_ptr(),
                                            out.data_ptr()))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    C = max(gamma, 1)
    flop = (n - 1) * P * C * 68.0
    print("n=%d P=%d C=%d B=%d: %.3f ms/batch  %.1f evals/s  %.2f TFLOP/s fp64  %.1f us/tree" % (
        n, P, C, B, ms, B / ms * 1e3, flop * B / ms / 1e9, ms * 1e3 / B))
    from oracle import oracle as orc
    sm = tl.siteModel
    rates, props = sm.category_rates_and_proportions()
    k = min(B, 2)
    want = orc.tree_likelihood_batch(al.patterns, al.patternWeight, par[:k], hgt[:k], rates, props, f, 2.0)
    print("rel err", np.max(np.abs(out[:k].cpu().numpy() - want) / np.abs(want)))


if __name__ == "__main__":
    main()
