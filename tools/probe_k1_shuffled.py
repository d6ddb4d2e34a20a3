"""Development probe: density throughput when the nodes of every tree are numbered at random
(leaves among themselves, internal nodes among themselves) instead of the simulator's numbering
(internal nodes by height, root last)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import multitypetree_b200 as mtt
from multitypetree_b200 import _lib


def relabel_batch(tb, n, rng):
    B, N = tb.B, 2 * n - 1
    keyL = rng.random((B, n))
    keyI = rng.random((B, N - n))
    new = np.concatenate([np.argsort(np.argsort(keyL, axis=1), axis=1),
                          n + np.argsort(np.argsort(keyI, axis=1), axis=1)], axis=1)   # old -> new
    base = (np.arange(B) * N)[:, None]
    height = np.empty(B * N)
    nodeType = np.empty(B * N, np.uint8)
    parent = np.empty(B * N, np.int32)
    g = (base + new).ravel()
    height[g] = tb.height
    nodeType[g] = tb.nodeType
    par = tb.parent.reshape(B, N)
    newpar = np.where(par < 0, -1, np.take_along_axis(new, np.maximum(par, 0), axis=1))
    parent[g] = newpar.ravel()
    tree = np.searchsorted(tb.chgBase, np.arange(len(tb.chgTime)), side="right") - 1
    owner = new[tree, tb.chgNode.astype(np.int64)]
    order = np.lexsort((np.arange(len(owner)), owner, tree))        # stable: by tree, new owner
    chgCount = np.zeros(B * N, np.uint16)
    np.add.at(chgCount, tree * N + owner, 1)
    out = mtt.TreeBatch(tb.nodeOff.copy(), height, parent, nodeType, chgCount, tb.chgBase.copy(),
                        tb.chgTime[order], tb.chgType[order])
    out.set_owner_list(owner[order].astype(np.uint16))
    return out


def dev(a):
    t = torch.from_numpy(np.ascontiguousarray(a).ravel()).cuda()
    pad = torch.zeros(t.numel() + 64, dtype=t.dtype, device="cuda")
    pad[:t.numel()] = t
    return pad


def run(tb, m, N, d, ctx, L, label):
    B = tb.B
    arrs = [dev(x) for x in (m, N, tb.nodeOff, tb.height, tb.nodeType, tb.chgNode, tb.chgBase,
                             tb.chgTime, tb.chgType)]
    logP = torch.zeros(B, dtype=torch.float64, device="cuda")
    L.mtt_sc_set_max_tree_events(ctx.handle, int(tb.n_events().max()))
    p = [a.data_ptr() for a in arrs]

    def go():
        ctx.check(L.mtt_sc_logp_owner_batch_dev(ctx.handle, B, d, 1, p[0], p[1], p[2], p[3], None,
                                                p[4], p[5], p[6], p[7], p[8], 0, logP.data_ptr(), None))
    for _ in range(5):
        go()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            go()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    print("%-28s %.3f ms  %.1f M trees/s" % (label, best, B / best / 1e3))
    return logP.cpu().numpy()


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
    n, d = 100, 2
    m, N = mtt.synthetic_model(d)
    tb = mtt.simulate_trees(B, n, d, m, N, seed=2000)
    sh = relabel_batch(tb, n, np.random.default_rng(1))
    ctx = mtt.Context(0)
    L = _lib.lib()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    a = run(tb, m, N, d, ctx, L, "simulator numbering")
    b = run(sh, m, N, d, ctx, L, "random numbering")
    print("max rel diff between the two:", np.max(np.abs(a - b) / np.abs(a)))


if __name__ == "__main__":
    main()
