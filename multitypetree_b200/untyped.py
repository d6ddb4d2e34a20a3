"""Host-side mirror of distributions/StructuredCoalescentUntypedTreeDensity.java (SURVEY 8f
rank 1): the structured-coalescent density of an UNTYPED tree, estimated by importance sampling
over ``nParticles`` typings drawn with the uniformization path sampler.  All arithmetic runs in
``mtt_untyped_logp_batch`` (csrc/k4_untyped.cu, one warp per particle); this module draws the
random inputs from the chain's ``Randomizer`` and moves data.

Inputs keep the reference's names: ``migrationModel``, ``nParticles``, ``typeLabel``, ``tree``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

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


# ---------------------------------------------------------------------------------------------
# The typings themselves: untyped tree -> typed trees (tree initialisers)
# ---------------------------------------------------------------------------------------------
def untyped_sample_batch(um, N, trees: Sequence[MultiTypeTree], nParticles: int, leafOrder,
                         uniforms, uOff, chgCap: Optional[int] = None, model=None, ctx=None):
    """mtt_untyped_sample_batch: every particle returned as a typing of its tree.  Inputs as
    ``untyped_logp_batch``.  Returns dict(nodeType: list per tree of u8 [P, n], nChg [B, P],
    chgTime / chgType / chgNode [B, P, chgCap], logW, status, uUsed [B, P]); room for the
    changes is doubled until every particle fits."""
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
    cap = int(chgCap) if chgCap is not None else max(16, int(np.diff(nodeOff).max()))
    while True:
        nodeType = np.empty(int(nodeOff[B]) * nParticles, np.uint8)
        nChg = np.empty(nPart, np.int32)
        chgTime = np.empty(nPart * cap)
        chgType = np.empty(nPart * cap, np.uint8)
        chgNode = np.empty(nPart * cap, np.int32)
        logW = np.empty(nPart)
        status = np.empty(nPart, np.int32)
        used = np.empty(nPart, np.int64)
        ctx.check(_lib.lib().mtt_untyped_sample_batch(
            ctx.handle, B, nParticles, _lib.ptr(mdl), _lib.ptr(N), _lib.ptr(nodeOff),
            _lib.ptr(height), _lib.ptr(parent), _lib.ptr(leafType), _lib.ptr(order), _lib.ptr(u),
            _lib.ptr(uOff), cap, _lib.ptr(nodeType), _lib.ptr(nChg), _lib.ptr(chgTime),
            _lib.ptr(chgType), _lib.ptr(chgNode), _lib.ptr(logW), _lib.ptr(status),
            _lib.ptr(used)))
        if chgCap is not None or nChg.max(initial=0) <= cap:
            break
        cap = 2 * int(nChg.max())
    types = [nodeType[nodeOff[b] * nParticles:nodeOff[b + 1] * nParticles]
             .reshape(nParticles, int(nodeOff[b + 1] - nodeOff[b])) for b in range(B)]
    return dict(nodeType=types, nChg=nChg.reshape(B, nParticles),
                chgTime=chgTime.reshape(B, nParticles, cap),
                chgType=chgType.reshape(B, nParticles, cap),
                chgNode=chgNode.reshape(B, nParticles, cap), logW=logW.reshape(B, nParticles),
                status=status.reshape(B, nParticles), uUsed=used.reshape(B, nParticles))


def typed_tree_from_sample(tree: MultiTypeTree, nodeType, chgNode, chgType, chgTime) -> MultiTypeTree:
    """One particle of ``untyped_sample_batch`` as a MultiTypeTree (changes appended to their
    nodes in the order the sampler added them: increasing time per node)."""
    n = tree.getNodeCount()
    out = MultiTypeTree(np.array(tree.height, np.float64), np.array(tree.parent, np.int32),
                        np.array(nodeType, np.uint8), [[] for _ in range(n)],
                        [[] for _ in range(n)], None if tree.taxa is None else list(tree.taxa))
    for v, ty, t in zip(chgNode, chgType, chgTime):
        out.addChange(int(v), int(ty), float(t))
    return out


def colour_untyped_trees(migrationModel, trees: Sequence[MultiTypeTree], rng: Randomizer,
                         ctx=None):
    """The colouring loop of MultiTypeTreeFromUntypedNewick.java:104-134 for many untyped trees
    at once: one particle per tree and round; trees whose colouring has no valid path
    (NoValidPathException, "Colour simulation failed.  Retrying.") draw a new leaf order and go
    again, all in one device call per round."""
    um = migrationModel._ensure_device(False)
    _, N = migrationModel.dense()
    mu = um.matrices(0)[2]
    out: List[Optional[MultiTypeTree]] = [None] * len(trees)
    todo = list(range(len(trees)))
    win = 0
    for t in trees:
        L = float(t.height.max() - t.height.min())
        per = 2.0 * (mu * L + 6.0 * np.sqrt(mu * L + 1.0)) + 8.0
        win = max(win, int(64 + t.getLeafNodeCount() * per / 2.0))
    for _ in range(1000):
        if not todo:
            return out
        sub = [trees[i] for i in todo]
        orders = [shuffled(rng, t.getLeafNodeCount()) for t in sub]
        us = [rng.nextDoubles(win) for _ in sub]
        uOff = np.arange(len(sub) + 1, dtype=np.int64) * win
        r = untyped_sample_batch(um, N, sub, 1, np.concatenate(orders), np.concatenate(us), uOff,
                                 ctx=ctx)
        again = []
        for k, i in enumerate(todo):
            st = int(r["status"][k, 0])
            if st == ST_OK:
                c = int(r["nChg"][k, 0])
                out[i] = typed_tree_from_sample(sub[k], r["nodeType"][k][0],
                                                r["chgNode"][k, 0, :c], r["chgType"][k, 0, :c],
                                                r["chgTime"][k, 0, :c])
            else:
                again.append(i)
                if st == ST_NEED_MORE_UNIFORMS:
                    win *= 2
        todo = again
    raise RuntimeError("colour simulation failed 1000 times")


def _untyped_from_flat(flat, leafTypeOf) -> MultiTypeTree:
    """Array tree of an untyped Newick: leaves in order of appearance, internal nodes in
    post-order (root last), single-child nodes skipped (TreeParser 'singlechild')."""
    leaves, internals = [], []
    heights, parents, names = {}, {}, {}
    counter = [0]

    def build(fnode, parent_id):
        while len(fnode.children) == 1:
            fnode = fnode.children[0]
        my = counter[0]
        counter[0] += 1
        heights[my], parents[my], names[my] = fnode.height, parent_id, fnode.label
        if not fnode.children:
            leaves.append(my)
        elif len(fnode.children) == 2:
            for c in fnode.children:
                build(c, my)
            internals.append(my)
        else:
            raise ValueError("multifurcations are not supported")
        return my

    build(flat, -1)
    order = leaves + internals
    newnr = {tmp: i for i, tmp in enumerate(order)}
    n = len(order)
    nodeType = np.zeros(n, np.uint8)
    for i, tmp in enumerate(leaves):
        nodeType[i] = leafTypeOf(names[tmp])
    return MultiTypeTree(np.array([heights[t] for t in order], np.float64),
                         np.array([newnr[parents[t]] if parents[t] >= 0 else -1 for t in order],
                                  np.int32),
                         nodeType, [[] for _ in range(n)], [[] for _ in range(n)],
                         [names[t] for t in leaves])


def MultiTypeTreeFromUntypedNewick(value: str, migrationModel, typeTrait, typeLabel: str = "type",
                                   randomizer: Optional[Randomizer] = None, ctx=None,
                                   adjustTipHeights: bool = True) -> MultiTypeTree:
    """evolution/tree/MultiTypeTreeFromUntypedNewick.java: a typed tree from an untyped Newick
    string, the leaf types from ``typeTrait`` (taxon -> type name or index; the reference's
    TraitSet named ``typeLabel``), the rest coloured lineage by lineage with the uniformization
    sampler on the device.  Inputs keep the reference's names (``value``, ``migrationModel``,
    ``typeLabel``, ``adjustTipHeights`` -- default true here, :37-38: BEAST's TreeParser then
    puts every tip at height 0 when there is no date trait [EXT])."""
    from .tree import _type_index, parse_flat_newick
    if typeTrait is None:
        raise ValueError("Trait set (with name '%s') must be provided." % typeLabel)   # :98-101
    typeSet = getattr(migrationModel, "typeSet", None)
    tree = _untyped_from_flat(parse_flat_newick(value),
                              lambda name: _type_index(typeTrait[name], typeSet))
    if adjustTipHeights:
        tree.height[:tree.getLeafNodeCount()] = 0.0
    return colour_untyped_trees(migrationModel, [tree], randomizer or Randomizer(127), ctx=ctx)[0]


def StructuredCoalescentUntypedTree(migrationModel, leafTypes, leafTimes=None, seed: int = 1,
                                    ctx=None) -> MultiTypeTree:
    """evolution/tree/StructuredCoalescentUntypedTree.java:103-294: a tree simulated under the
    structured coalescent from leaf types and times, returned WITHOUT its type changes (a
    regular BEAST tree: topology, heights, leaf types).  The simulation is the device
    generator behind ``DeviceTreeBatch`` (the same event loop as
    StructuredCoalescentMultiTypeTree); internal node types are zeroed."""
    from .tree import DeviceTreeBatch
    m, N = migrationModel.dense()
    leafTypes = np.asarray(leafTypes, np.int32)
    if leafTimes is None:                     # no date trait: every leaf at height 0 (:136-139)
        leafTimes = np.zeros(len(leafTypes))
    dv = DeviceTreeBatch(1, len(leafTypes), len(N), m, N, seed=seed, ctx=ctx,
                         leafTypes=leafTypes, leafTimes=leafTimes)
    t = dv.to_host().tree(0)
    dv.close()
    n = t.getNodeCount()
    nodeType = np.array(t.nodeType, np.uint8)
    nodeType[t.getLeafNodeCount():] = 0
    return MultiTypeTree(t.height, t.parent, nodeType, [[] for _ in range(n)],
                         [[] for _ in range(n)], t.taxa)
