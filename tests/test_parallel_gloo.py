"""CPU tests of the N>1 path (world_size 2, gloo): tree sharding without a collective on the
data path, and pattern sharding with one all-reduce.  The per-rank compute is stood in for by the
oracle here (no GPU in this container); on GPUs the same helpers wrap the CUDA entry points."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

import multitypetree_b200 as mtt
from multitypetree_b200 import parallel
from multitypetree_b200.likelihood import Alignment, simulate_alignment


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, what, q):
    import torch.distributed as dist
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m, N = mtt.synthetic_model(2)
        tb = mtt.simulate_trees(37, 30, 2, m, N, seed=11, nThreads=1)
        if what == "trees":
            got = parallel.sharded_density(tb, lambda b: orc.sc_logp_batch(b, m, N))
            want = orc.sc_logp_batch(tb, m, N)
        else:
            par, hgt = tb.parent.reshape(37, -1)[:5], tb.height.reshape(37, -1)[:5]
            f = np.array([.25] * 4)
            al = Alignment(states=simulate_alignment(par[0], hgt[0], 400, 2.0, f, 0.1, seed=1))
            sh = al.shard(rank, world)
            got = parallel.pattern_sharded_logl(lambda: orc.tree_likelihood_batch(
                sh.patterns, sh.patternWeight, par, hgt, [0.1], [1.0], f, 2.0))
            want = orc.tree_likelihood_batch(al.patterns, al.patternWeight, par, hgt, [0.1], [1.0],
                                             f, 2.0)
        if what == "parts":
            # pattern shards under a proportion of invariant sites: every rank hands (unscaled
            # sum, rescaled sum, "the unscaled evaluation underflowed") to the all-reduce and
            # BEAST's choice between the two is made over ALL patterns.  The
            # per-rank evaluation is the oracle's two passes; -inf marks the underflow.
            def shard_parts():
                kw = dict(pInv=0.2)
                u = orc.tree_likelihood_batch(sh.patterns, sh.patternWeight, par, hgt, [0.1 / 0.8],
                                              [0.8], f, 2.0, useScaling=0, **kw)
                r = orc.tree_likelihood_batch(sh.patterns, sh.patternWeight, par, hgt, [0.1 / 0.8],
                                              [0.8], f, 2.0, useScaling=1, **kw)
                flag = np.isinf(u).astype(np.float64)
                return np.stack([np.where(np.isinf(u), 0.0, u), r, flag])

            class FakeChains:                       # the shape of PeelChains.update(parts=True)
                def update(self, parent, height, parts=False, **_):
                    assert parts
                    return shard_parts()
            got = parallel.pattern_sharded_logl_parts(shard_parts)
            got2 = parallel.pattern_sharded_chains_update(FakeChains(), par, hgt)
            assert np.array_equal(got, got2)
            want = orc.tree_likelihood_batch(al.patterns, al.patternWeight, par, hgt, [0.1 / 0.8],
                                             [0.8], f, 2.0, pInv=0.2, useScaling=2)
        q.put((rank, float(np.max(np.abs(got - want) / np.abs(want)))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("what", ["trees", "patterns", "parts"])
def test_world_size_2(what):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, what, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r for r, _ in res) == [0, 1]
    assert all(err < 1e-12 for _, err in res)


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 4096, 1000003):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
