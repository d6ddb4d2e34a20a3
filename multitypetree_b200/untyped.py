"""Host-side mirror of distributions/StructuredCoalescentUntypedTreeDensity.java (SURVEY 8f
rank 1): the structured-coalescent density of an UNTYPED tree, estimated by importance sampling
over ``nParticles`` typings drawn with the uniformization path sampler.  All arithmetic runs in
``mtt_untyped_logp_batch`` (csrc/k4_untyped.cu, one warp per particle); this module draws the
random inputs from the chain's ``Randomizer`` and moves data.

Inputs keep the reference's names: ``migrationModel``, ``nParticles``, ``typeLabel``, ``tree``.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from . import _lib
from .tree import MultiTypeTree
from .uniformization import Randomizer, ST_NEED_MORE_UNIFORMS, ST_OK


def shuffled(rng: Randomizer, n: int) -> np.ndarray:
    """BEAST ``Randomizer.shuffled(n)`` ([EXT]): identity permutation, then for i = 0..n-1
    swap element i with element i + nextInt(n - i)."""
    a = np.arange(n, dtype=np.int32)
    for i in range(n):
        j = rng.nextInt(n - i) + i
        a[i], a[j] = a[j], a[i]
    return a


def untyped_logp_batch(um, N, trees: Sequence[MultiTypeTree], nParticles: int, leafOrder, uniforms,
                       uOff, model=None, ctx=None):
    """mtt_untyped_logp_batch for a list of trees (topology, heights and LEAF types are read).
    leafOrder: int32, tree-major / particle-major permutations; uniforms + uOff[nTrees*P+1]: one
    window per particle.  Returns dict(logP[nTrees], logW, status, uUsed [nTrees, P])."""
    ctx = ctx or um.ctx
    um._activate()
    B = len(trees)
    nodeOff = np.zeros(B + 1, np.int64)
    for b, t in enumerate(trees):
        nodeOff[b + 1] = nodeOff[b] + t.getNodeCount()
    height = np.concatenate([np.asarray(t.height, np.float64) for t in trees])
    parent = np.concatenate([np.asarray(t.parent, np.int32) for t in trees])
    leafType = np.concatenate([np.asarray(t.nodeType, np.uint8) for t in trees])
    N = _lib.as_c(N, np.float64)
    order = _lib.as_c(leafOrder, np.int32).ravel()
    u = _lib.as_c(uniforms, np.float64).ravel()
    uOff = _lib.as_c(uOff, np.int64)
    mdl = None if model is None else _lib.as_c(model, np.int32)
    nPart = B * nParticles
    logP = np.empty(B)
    logW = np.empty(nPart)
    status = np.empty(nPart, np.int32)
    used = np.empty(nPart, np.int64)
    ctx.check(_lib.lib().mtt_untyped_logp_batch(
        ctx.handle, B, nParticles, _lib.ptr(mdl), _lib.ptr(N), _lib.ptr(nodeOff),
        _lib.ptr(height), _lib.ptr(parent), _lib.ptr(leafType), _lib.ptr(order), _lib.ptr(u),
        _lib.ptr(uOff), _lib.ptr(logP), _lib.ptr(logW), _lib.ptr(status), _lib.ptr(used)))
    return dict(logP=logP, logW=logW.reshape(B, nParticles),
                status=status.reshape(B, nParticles), uUsed=used.reshape(B, nParticles))


class StructuredCoalescentUntypedTreeDensity:
    """distributions/StructuredCoalescentUntypedTreeDensity.java.  ``tree`` supplies topology,
    node heights and the leaf types (the reference reads them from the trait set named
    ``typeLabel``).  ``isStochastic()`` is true: every evaluation draws new particles."""

    def __init__(self, migrationModel, tree: Optional[MultiTypeTree] = None, nParticles: int = 100,
                 typeLabel: str = "type", randomizer: Optional[Randomizer] = None, ctx=None):
        self.migrationModel = migrationModel
        self.tree = tree
        self.nParticles = int(nParticles)
        self.typeLabel = typeLabel
        self.rng = randomizer or Randomizer(127)
        self.ctx = ctx
        self.logP = float("nan")

    def _window(self, tree: MultiTypeTree, mu: float) -> int:
        # a lineage uses 1 + 2 nVirt uniforms, nVirt ~ Poisson(mu L) conditioned on the end type
        L = float(tree.height.max() - tree.height.min())
        per = 2.0 * (mu * L + 6.0 * np.sqrt(mu * L + 1.0)) + 8.0
        return int(64 + tree.getLeafNodeCount() * per / 2.0)

    def calculateLogPBatch(self, trees: Sequence[MultiTypeTree]) -> np.ndarray:
        """One estimate per tree; all trees x particles in one device call.  Every particle
        gets its leaf order and then its own window of uniforms from the Randomizer (windows
        are consumed whole, so the stream differs from a particle-by-particle Java run)."""
        um = self.migrationModel._ensure_device(False)
        _, N = self.migrationModel.dense()
        mu = um.matrices(0)[2]
        P = self.nParticles
        win = max(self._window(t, mu) for t in trees)
        while True:
            orders, us = [], []
            for t in trees:
                n = t.getLeafNodeCount()
                for _ in range(P):
                    orders.append(shuffled(self.rng, n))
                    us.append(self.rng.nextDoubles(win))
            uOff = np.arange(len(trees) * P + 1, dtype=np.int64) * win
            r = untyped_logp_batch(um, N, trees, P, np.concatenate(orders), np.concatenate(us),
                                   uOff, ctx=self.ctx)
            if np.all(r["status"] == ST_OK):
                return r["logP"]
            if np.any((r["status"] != ST_OK) & (r["status"] != ST_NEED_MORE_UNIFORMS)):
                raise RuntimeError("untyped density: unexpected particle status %r"
                                   % np.unique(r["status"]))
            win *= 4

    def calculateLogP(self) -> float:
        """StructuredCoalescentUntypedTreeDensity.java:116-207."""
        self.logP = float(self.calculateLogPBatch([self.tree])[0])
        return self.logP

    def requiresRecalculation(self) -> bool:
        return True

    def isStochastic(self) -> bool:
        return True
