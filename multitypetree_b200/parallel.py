"""Multi-GPU decomposition of the hot path, one process per GPU.

* trees / chains are independent -> contiguous shards of the batch per rank, NO collective on
  the data path (the host gathers B doubles);
* very large alignments (config 4) -> site PATTERNS are sharded contiguously, the tree is
  replicated, and the per-rank partial sums  sum_p w_p logL_p  of the B trees are combined with
  one fp64 all-reduce (NCCL over NVLink on GPUs; gloo in the CPU tests).

`torch.distributed` is plumbing only; the arithmetic is in the CUDA library.
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np

from .tree import TreeBatch


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) of n items for `rank` of `world`."""
    return (n * rank) // world, (n * (rank + 1)) // world


def shard_batch(batch: TreeBatch, rank: int, world: int) -> TreeBatch:
    lo, hi = shard_range(batch.B, rank, world)
    return batch.slice(lo, hi)


def _dist():
    import torch.distributed as dist
    return dist


def world_info() -> Tuple[int, int]:
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def allreduce_sum(x, group=None):
    """In-place sum over ranks of a torch tensor (CUDA -> NCCL, CPU -> gloo) or numpy array."""
    import torch
    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x
    if isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x))
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        x[...] = t.cpu().numpy()
        return x
    dist.all_reduce(x, op=dist.ReduceOp.SUM, group=group)
    return x


def gather_concat(local: np.ndarray, total: int, group=None) -> np.ndarray:
    """Concatenate per-rank result slices (shard_range order) on every rank."""
    import torch
    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    out = np.zeros(total, local.dtype)
    lo, hi = shard_range(total, rank, world)
    out[lo:hi] = local
    return allreduce_sum(out, group)          # disjoint slices: a sum is a concatenation


def sharded_density(batch: TreeBatch, compute: Callable[[TreeBatch], np.ndarray],
                    group=None) -> np.ndarray:
    """logP of every tree of `batch`, trees sharded across ranks; `compute` evaluates a
    sub-batch on this rank's GPU (e.g. lambda b: sc_logp_batch(b, m, N, ctx=ctx))."""
    rank, world = world_info()
    local = compute(shard_batch(batch, rank, world))
    return gather_concat(np.asarray(local, np.float64), batch.B, group)


def pattern_sharded_logl(compute_partial: Callable[[], np.ndarray], group=None) -> np.ndarray:
    """logL[B] over ALL patterns from this rank's partial sums over ITS pattern shard
    (a TreeLikelihood built on ``alignment.shard(rank, world)``): one all-reduce."""
    return allreduce_sum(np.asarray(compute_partial(), np.float64).copy(), group)


def pattern_sharded_logl_parts(compute_parts: Callable[[], np.ndarray], group=None) -> np.ndarray:
    """The same for alignments with a proportion of invariant sites: `compute_parts` returns this
    rank's parts[3][B] of mtt_peel_logl_parts_batch_dev (unscaled sum, rescaled sum, number of
    tiles that saw an underflow); BEAST's choice between the two evaluations is made after the
    all-reduce, over ALL patterns."""
    parts = allreduce_sum(np.asarray(compute_parts(), np.float64).copy(), group)
    return np.where(parts[2] > 0, parts[1], parts[0])


def pattern_sharded_chains_update(chains, parent, height, group=None, **kw) -> np.ndarray:
    """One proposal per resident chain when every rank holds a pattern shard (`chains` is a
    ``PeelChains`` of a TreeLikelihood built on ``alignment.shard(rank, world)``; trees are
    replicated): this rank's parts of ``mtt_peel_chains_update_parts``, one all-reduce of
    3 x nChains doubles, BEAST's choice between its two evaluations over ALL patterns.  Every rank
    gets the same logL[nChains] and must answer with the same ``chains.accept(mask)``."""
    return pattern_sharded_logl_parts(lambda: chains.update(parent, height, parts=True, **kw),
                                      group)
