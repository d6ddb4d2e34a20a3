"""Host-side mirror of distributions/StructuredCoalescentTreeDensity.java over the C-ABI.

Same Inputs as the reference class (``migrationModel``, ``multiTypeTree``, ``checkValidity``;
StructuredCoalescentTreeDensity.java:43-50, MultiTypeTreeDistribution.java:35-36) plus a batch
entry point.  All arithmetic runs in csrc/k1_density.cu.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .tree import MultiTypeTree, TreeBatch


def sc_logp_batch(batch: TreeBatch, m: np.ndarray, N: np.ndarray, checkValidity: bool = False,
                  ctx=None, with_parent: bool = None, want_valid: bool = False,
                  layout: str = "owner"):
    """logP of every tree of ``batch`` (host buffers).  m: [d,d] (shared model) or [B,d,d];
    N: [d] or [B,d].  layout "owner": mtt_sc_logp_owner_batch with the per-change owner list
    (what the kernels read); "counts": mtt_sc_logp_batch with per-node change counts."""
    ctx = ctx or _lib.default_context()
    L = _lib.lib()
    B = batch.B
    m = _lib.as_c(m, np.float64)
    N = _lib.as_c(N, np.float64)
    d = m.shape[-1]
    nModels = 1 if m.ndim == 2 else m.shape[0]
    if with_parent is None:
        with_parent = checkValidity or want_valid
    logP = np.empty(B, np.float64)
    valid = np.empty(B, np.uint8) if want_valid else None
    if layout not in ("owner", "counts"):
        raise ValueError("layout must be 'owner' or 'counts'")
    owner = layout == "owner"
    args = [_lib.as_c(batch.nodeOff, np.int64), _lib.as_c(batch.height, np.float64),
            _lib.as_c(batch.parent, np.int32) if with_parent else None,
            _lib.as_c(batch.nodeType, np.uint8),
            _lib.as_c(batch.chgNode if owner else batch.chgCount, np.uint16),
            _lib.as_c(batch.chgBase, np.int64), _lib.as_c(batch.chgTime, np.float64),
            _lib.as_c(batch.chgType, np.uint8)]
    fn = L.mtt_sc_logp_owner_batch if owner else L.mtt_sc_logp_batch
    ctx.check(fn(ctx.handle, B, d, nModels, _lib.ptr(m), _lib.ptr(N),
                 *[_lib.ptr(a) for a in args], int(bool(checkValidity)),
                 _lib.ptr(logP), _lib.ptr(valid)))
    return (logP, valid.astype(bool)) if want_valid else logP


def set_tips(leafHeight, leafType, ctx=None) -> None:
    """Install the tips every later ``sc_logp_tips_batch`` call shares (mtt_sc_set_tips): tip
    dates and tip types are data, fixed for a whole analysis."""
    ctx = ctx or _lib.default_context()
    h = _lib.as_c(leafHeight, np.float64)
    t = _lib.as_c(leafType, np.uint8)
    if h.shape != t.shape or h.ndim != 1:
        raise ValueError("leafHeight and leafType must be 1-d arrays of equal length")
    ctx.check(_lib.lib().mtt_sc_set_tips(ctx.handle, len(h), _lib.ptr(h), _lib.ptr(t)))
    ctx._tips = (h.copy(), t.copy())


def sc_logp_tips_batch(batch: TreeBatch, m: np.ndarray, N: np.ndarray,
                       checkValidity: bool = False, ctx=None, want_valid: bool = False):
    """logP of every tree of ``batch`` through mtt_sc_logp_tips_batch: all trees are over the
    taxon set installed by ``set_tips``; only internal-node heights / types and the change lists
    cross the host link.  Bit-identical to ``sc_logp_batch``."""
    ctx = ctx or _lib.default_context()
    tips = getattr(ctx, "_tips", None)
    if tips is None:
        raise ValueError("set_tips has not been called on this context")
    nL = len(tips[0])
    nN, B = 2 * nL - 1, batch.B
    if not np.array_equal(batch.nodeOff, np.arange(B + 1, dtype=np.int64) * nN):
        raise ValueError("every tree must have %d nodes" % nN)
    H = np.asarray(batch.height).reshape(B, nN)
    T = np.asarray(batch.nodeType).reshape(B, nN)
    if B and not (np.array_equal(H[:, :nL], np.broadcast_to(tips[0], (B, nL))) and
                  np.array_equal(T[:, :nL], np.broadcast_to(tips[1], (B, nL)))):
        raise ValueError("a tree's tips differ from the installed ones")
    m = _lib.as_c(m, np.float64)
    N = _lib.as_c(N, np.float64)
    d = m.shape[-1]
    nModels = 1 if m.ndim == 2 else m.shape[0]
    with_parent = checkValidity or want_valid
    intH = _lib.as_c(H[:, nL:], np.float64)
    intT = _lib.as_c(T[:, nL:], np.uint8)
    par = _lib.as_c(batch.parent, np.int32) if with_parent else None
    logP = np.empty(B, np.float64)
    valid = np.empty(B, np.uint8) if want_valid else None
    ctx.check(_lib.lib().mtt_sc_logp_tips_batch(
        ctx.handle, B, d, nModels, _lib.ptr(m), _lib.ptr(N), _lib.ptr(intH), _lib.ptr(par),
        _lib.ptr(intT), _lib.ptr(_lib.as_c(batch.chgNode, np.uint16)),
        _lib.ptr(_lib.as_c(batch.chgBase, np.int64)), _lib.ptr(_lib.as_c(batch.chgTime, np.float64)),
        _lib.ptr(_lib.as_c(batch.chgType, np.uint8)), int(bool(checkValidity)), _lib.ptr(logP),
        _lib.ptr(valid)))
    return (logP, valid.astype(bool)) if want_valid else logP


def tree_validity(batch: TreeBatch, ctx=None) -> np.ndarray:
    """MultiTypeTree.isValid for every tree (MultiTypeTree.java:396-429)."""
    d = int(max(batch.nodeType.max(initial=0), batch.chgType.max(initial=0))) + 1
    d = max(d, 2)
    m = np.ones((d, d)) - np.eye(d)
    _, valid = sc_logp_batch(batch, m, np.ones(d), False, ctx, True, True)
    return valid


class StructuredCoalescentTreeDensity:
    """Likelihood of a MultiTypeTree under the structured coalescent."""

    def __init__(self, migrationModel, multiTypeTree: MultiTypeTree = None,
                 checkValidity: bool = False, ctx=None):
        self.migrationModel = migrationModel
        self.mtTree = multiTypeTree
        self.checkValidity = bool(checkValidity)
        self.ctx = ctx
        self.logP = float("nan")
        self.initAndValidate()

    def initAndValidate(self):
        # type-set compatibility (StructuredCoalescentTreeDensity.java:81-83)
        if self.mtTree is not None:
            top = int(self.mtTree.nodeType.max(initial=0))
            for c in self.mtTree.changeTypes:
                if c:
                    top = max(top, max(c))
            if top >= self.migrationModel.getNTypes():
                raise ValueError("Tree and migration model have incompatible type sets.")

    def calculateLogP(self) -> float:
        """StructuredCoalescentTreeDensity.java:86-147."""
        m, N = self.migrationModel.dense()
        lp = sc_logp_batch(TreeBatch.from_trees([self.mtTree]), m, N, self.checkValidity,
                           self.ctx)
        self.logP = float(lp[0])
        return self.logP

    def calculateLogPBatch(self, batch: TreeBatch) -> np.ndarray:
        m, N = self.migrationModel.dense()
        return sc_logp_batch(batch, m, N, self.checkValidity, self.ctx)

    def requiresRecalculation(self) -> bool:
        return True                                                          # :250-253
