"""Host-side mirror of the reference's typed-tree state, flattened to the tree-SoA layout.

Mirrors (names and meaning, not code) evolution/tree/MultiTypeNode.java:33-102,
evolution/tree/MultiTypeTree.java:396-429 (isValid), :538-665 (initFromFlatTree: single-child
Newick nodes are type changes) and evolution/tree/MultiTypeTreeFromNewick.java:30-43.
All arithmetic on these trees happens in the CUDA library; this module only moves data.
"""
from __future__ import annotations

import ctypes as C
import re
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib


@dataclass
class MultiTypeTree:
    """One flattened multi-type tree.  Node Nr in [0, nLeaves) are leaves (BEAST numbering)."""

    height: np.ndarray            # f64 [nNodes]
    parent: np.ndarray            # i32 [nNodes], -1 = root
    nodeType: np.ndarray          # u8  [nNodes]
    changeTimes: List[List[float]] = field(default_factory=list)   # per node, youngest first
    changeTypes: List[List[int]] = field(default_factory=list)
    taxa: Optional[List[str]] = None
    # explicit child order (BEAST's getLeft / getRight); built from `parent` on first use and
    # kept up to date by the topology-changing operators (operators.py)
    left: Optional[np.ndarray] = None     # i32 [nNodes], -1 = none
    right: Optional[np.ndarray] = None

    # ---- MultiTypeNode / Tree accessors (same names as the reference) ----
    def _ensure_children(self) -> None:
        if self.left is not None:
            return
        n = len(self.height)
        self.left = np.full(n, -1, np.int32)
        self.right = np.full(n, -1, np.int32)
        for v in range(n):                       # lower node number = left child
            p = int(self.parent[v])
            if p >= 0:
                if self.left[p] < 0:
                    self.left[p] = v
                else:
                    self.right[p] = v

    def getLeft(self, node: int) -> int:
        self._ensure_children()
        return int(self.left[node])

    def getRight(self, node: int) -> int:
        self._ensure_children()
        return int(self.right[node])

    def getParent(self, node: int) -> int:
        return int(self.parent[node])

    def isRoot(self, node: int) -> bool:
        return self.parent[node] < 0

    def isLeaf(self, node: int) -> bool:
        return node < self.getLeafNodeCount()

    def getHeight(self, node: int) -> float:
        return float(self.height[node])

    def setHeight(self, node: int, h: float) -> None:
        self.height[node] = h

    def removeChild(self, node: int, child: int) -> None:
        """beast Node.removeChild: drop `child` from the ordered child list of `node`."""
        self._ensure_children()
        if self.left[node] == child:
            self.left[node], self.right[node] = self.right[node], -1
        elif self.right[node] == child:
            self.right[node] = -1

    def addChild(self, node: int, child: int) -> None:
        """beast Node.addChild: append to the child list and set the child's parent."""
        self._ensure_children()
        if self.left[node] < 0:
            self.left[node] = child
        else:
            self.right[node] = child
        self.parent[child] = node

    def truncateChanges(self, node: int, newNChanges: int) -> None:
        """MultiTypeNode.truncateChanges (MultiTypeNode.java:164-173)."""
        del self.changeTimes[node][newNChanges:]
        del self.changeTypes[node][newNChanges:]

    def setChangeTime(self, node: int, idx: int, time: float) -> None:
        self.changeTimes[node][idx] = float(time)

    def getNodeCount(self) -> int:
        return len(self.height)

    def getLeafNodeCount(self) -> int:
        return (len(self.height) + 1) // 2

    def getInternalNodeCount(self) -> int:
        return len(self.height) - self.getLeafNodeCount()

    def getRoot(self) -> int:
        return int(np.nonzero(self.parent < 0)[0][0])

    def getChangeCount(self, node: int) -> int:
        return len(self.changeTimes[node])

    def getChangeTime(self, node: int, idx: int) -> float:
        return self.changeTimes[node][idx]

    def getChangeType(self, node: int, idx: int) -> int:
        return self.changeTypes[node][idx]

    def getNodeType(self, node: int) -> int:
        return int(self.nodeType[node])

    def getFinalType(self, node: int) -> int:
        return self.changeTypes[node][-1] if self.changeTypes[node] else int(self.nodeType[node])

    def setNodeType(self, node: int, t: int) -> None:
        self.nodeType[node] = t

    def clearChanges(self, node: int) -> None:
        self.changeTimes[node] = []
        self.changeTypes[node] = []

    def addChange(self, node: int, newType: int, time: float) -> None:
        self.changeTimes[node].append(float(time))
        self.changeTypes[node].append(int(newType))

    def getChildren(self, node: int) -> List[int]:
        self._ensure_children()
        return [int(c) for c in (self.left[node], self.right[node]) if c >= 0]

    def getTotalNumberOfChanges(self) -> int:
        """MultiTypeTree.java:694-705: changes on the root branch are not counted."""
        root = self.getRoot()
        return sum(len(c) for v, c in enumerate(self.changeTimes) if v != root)

    def copy(self) -> "MultiTypeTree":
        return MultiTypeTree(self.height.copy(), self.parent.copy(), self.nodeType.copy(),
                             [list(c) for c in self.changeTimes],
                             [list(c) for c in self.changeTypes],
                             list(self.taxa) if self.taxa else None,
                             None if self.left is None else self.left.copy(),
                             None if self.right is None else self.right.copy())

    def assignFrom(self, other: "MultiTypeTree") -> None:
        """Restore this tree in place from a copy (BEAST ``StateNode.restore``); operators keep
        their reference to the same object."""
        self.height[:] = other.height
        self.parent[:] = other.parent
        self.nodeType[:] = other.nodeType
        self.changeTimes = [list(c) for c in other.changeTimes]
        self.changeTypes = [list(c) for c in other.changeTypes]
        self.left = None if other.left is None else other.left.copy()
        self.right = None if other.right is None else other.right.copy()

    def getFlattenedNewick(self, typeLabel: str = "type") -> str:
        """The 'flat' Newick of evolution/tree/MultiTypeTree.java:444-526 (getFlattenedTree):
        every type change becomes a single-child node whose ``[&typeLabel=k]`` is the type
        above it and whose height is the change time -- the format
        ``MultiTypeTreeFromNewick`` / ``initFromFlatTree`` reads back, and the one BEAST uses to
        log and checkpoint multi-type trees."""
        self._ensure_children()
        kids = {v: self.getChildren(v) for v in range(len(self.parent)) if self.left[v] >= 0}
        root = self.getRoot()
        pieces = {}
        # iterative post-order so that 5000-tip caterpillars do not hit the recursion limit
        stack, order = [root], []
        while stack:
            v = stack.pop()
            order.append(v)
            stack.extend(kids.get(v, []))
        for v in reversed(order):
            ch = kids.get(v, [])
            name = (self.taxa[v] if self.taxa and v < len(self.taxa) and self.taxa[v] else str(v)) \
                if not ch else ""
            body = ("(" + ",".join(pieces.pop(c) for c in ch) + ")" if ch else "") + name
            s = "%s[&%s=%d]" % (body, typeLabel, int(self.nodeType[v]))
            if v == root:
                pieces[v] = s + ":0.0;"
                continue
            t_prev = float(self.height[v])
            for t, c in zip(self.changeTimes[v], self.changeTypes[v]):
                s = "(%s:%r)[&%s=%d]" % (s, float(t) - t_prev, typeLabel, int(c))
                t_prev = float(t)
            pieces[v] = "%s:%r" % (s, float(self.height[int(self.parent[v])]) - t_prev)
        return pieces[root]

    # ---- Loggable (MultiTypeTree.java:795-831): the NEXUS tree log BEAST writes and resumes from
    def _taxon(self, i: int) -> str:
        return self.taxa[i] if self.taxa and i < len(self.taxa) and self.taxa[i] else str(i)

    def init(self, stream, typeLabel: str = "type") -> None:
        """Header of the tree log: taxa block and the Translate table (leaf Nr + 1 -> taxon)."""
        n = self.getLeafNodeCount()
        stream.write("#NEXUS\n\nBegin taxa;\n\tDimensions ntax=%d;\n\t\tTaxlabels\n" % n)
        for i in range(n):
            stream.write("\t\t\t%s\n" % self._taxon(i))
        stream.write("\t\t\t;\nEnd;\nBegin trees;\n\tTranslate\n")
        for i in range(n):
            stream.write("\t\t\t%d %s%s\n" % (i + 1, self._taxon(i), "," if i < n - 1 else ""))
        stream.write("\t\t\t;")
        self._logLabel = typeLabel

    def log(self, sample: int, stream) -> None:
        """One ``tree STATE_i = <flat Newick with translated leaf numbers>;`` statement."""
        saved = self.taxa
        self.taxa = [str(i + 1) for i in range(self.getLeafNodeCount())]
        try:
            nw = self.getFlattenedNewick(getattr(self, "_logLabel", "type"))
        finally:
            self.taxa = saved
        stream.write("\ntree STATE_%d = %s" % (sample, nw))

    def close(self, stream) -> None:
        stream.write("\nEnd;\n")

    def isValid(self, ctx=None) -> bool:
        """MultiTypeTree.isValid (MultiTypeTree.java:396-429), evaluated on the device."""
        from .density import tree_validity
        return bool(tree_validity(TreeBatch.from_trees([self]), ctx)[0])


@dataclass
class TreeBatch:
    """B trees back to back in the C-ABI layout (include/mtt_b200.h)."""

    nodeOff: np.ndarray    # i64 [B+1]
    height: np.ndarray     # f64 [nodes]
    parent: np.ndarray     # i32 [nodes]
    nodeType: np.ndarray   # u8  [nodes]
    chgCount: np.ndarray   # u16 [nodes]
    chgBase: np.ndarray    # i64 [B+1]
    chgTime: np.ndarray    # f64 [changes]
    chgType: np.ndarray    # u8  [changes]
    # owner layout of the change lists (see chgNode); filled by the simulator / set_owner_list,
    # otherwise derived from chgCount on first use
    owners: Optional[np.ndarray] = field(default=None, repr=False, compare=False)

    @property
    def B(self) -> int:
        return len(self.nodeOff) - 1

    @property
    def chgNode(self) -> np.ndarray:
        """u16 [changes]: tree-local number of the node that owns each change -- the OWNER
        layout of ``mtt_sc_logp_owner_batch`` (derived once from ``chgCount`` and cached)."""
        cached = self.owners
        if cached is None or len(cached) != len(self.chgTime):
            csum = np.concatenate([[0], np.cumsum(self.chgCount, dtype=np.int64)])
            if not np.array_equal(csum[self.nodeOff], self.chgBase):
                raise ValueError("chgCount does not add up to chgBase: no owner list exists for "
                                 "this batch (the counts layout reports such trees as invalid)")
            local = np.arange(len(self.chgCount), dtype=np.int64) - np.repeat(
                self.nodeOff[:-1], np.diff(self.nodeOff))
            cached = np.repeat(local, self.chgCount).astype(np.uint16)
            self.owners = cached
        return cached

    def set_owner_list(self, chgNode: Optional[np.ndarray]) -> None:
        """Install an owner list (u16 [changes]) as is, or drop the cached one (None)."""
        self.owners = None if chgNode is None else np.ascontiguousarray(chgNode, np.uint16)

    def n_events(self) -> np.ndarray:
        """E = nodes + changes per tree (the unit of algorithmic traffic)."""
        return np.diff(self.nodeOff) + np.diff(self.chgBase)

    @staticmethod
    def from_trees(trees: Sequence[MultiTypeTree]) -> "TreeBatch":
        nodeOff = np.zeros(len(trees) + 1, np.int64)
        chgBase = np.zeros(len(trees) + 1, np.int64)
        hs, ps, ts, ks, cts, ctys = [], [], [], [], [], []
        for b, t in enumerate(trees):
            n = t.getNodeCount()
            nodeOff[b + 1] = nodeOff[b] + n
            hs.append(np.asarray(t.height, np.float64))
            ps.append(np.asarray(t.parent, np.int32))
            ts.append(np.asarray(t.nodeType, np.uint8))
            root = t.getRoot()
            cnt = np.zeros(n, np.uint16)
            m = 0
            for i in range(n):
                if i == root:
                    continue      # root-branch changes are never exported
                c = len(t.changeTimes[i])
                cnt[i] = c
                m += c
                cts.extend(t.changeTimes[i])
                ctys.extend(t.changeTypes[i])
            ks.append(cnt)
            chgBase[b + 1] = chgBase[b] + m
        cat = lambda xs, dt: (np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt))
        return TreeBatch(nodeOff, cat(hs, np.float64), cat(ps, np.int32), cat(ts, np.uint8),
                         cat(ks, np.uint16), chgBase, np.asarray(cts, np.float64),
                         np.asarray(ctys, np.uint8))

    def tree(self, b: int) -> MultiTypeTree:
        n0, n1 = int(self.nodeOff[b]), int(self.nodeOff[b + 1])
        c = int(self.chgBase[b])
        times, types = [], []
        for i in range(n0, n1):
            k = int(self.chgCount[i])
            times.append([float(x) for x in self.chgTime[c:c + k]])
            types.append([int(x) for x in self.chgType[c:c + k]])
            c += k
        return MultiTypeTree(self.height[n0:n1].copy(), self.parent[n0:n1].copy(),
                             self.nodeType[n0:n1].copy(), times, types)

    def slice(self, b0: int, b1: int) -> "TreeBatch":
        n0, n1 = int(self.nodeOff[b0]), int(self.nodeOff[b1])
        c0, c1 = int(self.chgBase[b0]), int(self.chgBase[b1])
        out = TreeBatch(self.nodeOff[b0:b1 + 1] - n0, self.height[n0:n1], self.parent[n0:n1],
                        self.nodeType[n0:n1], self.chgCount[n0:n1],
                        self.chgBase[b0:b1 + 1] - c0, self.chgTime[c0:c1], self.chgType[c0:c1])
        if self.owners is not None and len(self.owners) == len(self.chgTime):
            out.owners = self.owners[c0:c1]
        return out

    def nbytes(self) -> int:
        """bytes the density path moves for this batch in the owner layout (parent[] excluded)"""
        return sum(a.nbytes for a in (self.nodeOff, self.height, self.nodeType, self.chgBase,
                                      self.chgTime, self.chgType)) + 2 * len(self.chgTime)


# ---------------------------------------------------------------------------- Newick

class _FlatNode:
    __slots__ = ("children", "label", "meta", "length", "height")

    def __init__(self):
        self.children = []
        self.label = None
        self.meta = {}
        self.length = 0.0
        self.height = 0.0


_META_ITEM = re.compile(r'\s*([^=,\s]+)\s*=\s*("([^"]*)"|\'([^\']*)\'|\{[^}]*\}|[^,\]]*)')


def _parse_meta(s: str) -> dict:
    out = {}
    for mm in _META_ITEM.finditer(s):
        key = mm.group(1)
        val = mm.group(3) if mm.group(3) is not None else (
            mm.group(4) if mm.group(4) is not None else mm.group(2).strip())
        out[key] = val
    return out


def parse_flat_newick(s: str) -> _FlatNode:
    """Newick with BEAST metadata comments ``[&k=v,...]`` and single-child nodes allowed
    (TreeParser with singlechild=true, MultiTypeTreeFromNewick.java:35-40)."""
    s = s.strip()
    pos = 0

    def parse_node() -> _FlatNode:
        nonlocal pos
        node = _FlatNode()
        if s[pos] == "(":
            pos += 1
            while True:
                node.children.append(parse_node())
                if s[pos] == ",":
                    pos += 1
                    continue
                if s[pos] == ")":
                    pos += 1
                    break
                raise ValueError("bad Newick at %d: %r" % (pos, s[pos:pos + 20]))
        # label
        start = pos
        if pos < len(s) and s[pos] in "'\"":
            q = s[pos]
            pos = s.index(q, pos + 1) + 1
            node.label = s[start + 1:pos - 1]
        else:
            while pos < len(s) and s[pos] not in ":,()[;":
                pos += 1
            node.label = s[start:pos].strip() or None
        # metadata and branch length (metadata may come before or after the colon)
        while pos < len(s) and s[pos] in "[:":
            if s[pos] == "[":
                end = s.index("]", pos)
                body = s[pos + 1:end]
                if body.startswith("&"):
                    node.meta.update(_parse_meta(body[1:]))
                pos = end + 1
            else:
                pos += 1
                start = pos
                while pos < len(s) and s[pos] not in ",()[;":
                    pos += 1
                txt = s[start:pos].strip()
                if txt:
                    node.length = float(txt)
        return node

    root = parse_node()
    # heights: distance from the youngest tip (TreeParser without adjustTipHeights keeps the
    # lengths; heights are measured back from the most recent tip)
    def depth(node, d0):
        node.height = d0
        for c in node.children:
            depth(c, d0 + c.length)
    depth(root, 0.0)
    stack, maxd = [root], 0.0
    while stack:
        nd = stack.pop()
        maxd = max(maxd, nd.height)
        stack.extend(nd.children)
    stack = [root]
    while stack:
        nd = stack.pop()
        nd.height = maxd - nd.height
        stack.extend(nd.children)
    return root


def _type_index(val, typeSet) -> int:
    """MultiTypeTree.java:568-582: Integer, rounded Double, or type-set name."""
    try:
        return int(val)
    except (TypeError, ValueError):
        pass
    try:
        return int(round(float(val)))
    except (TypeError, ValueError):
        pass
    if typeSet is None:
        raise ValueError("non-numeric type %r needs a typeSet" % (val,))
    return typeSet.getTypeIndex(str(val))


def MultiTypeTreeFromNewick(value: str, typeLabel: str = "type", typeSet=None) -> MultiTypeTree:
    """Build a MultiTypeTree from a 'flat' Newick string whose single-child nodes mark type
    changes: a single-child node's ``[&typeLabel=k]`` is the type ABOVE it and its height the
    change time; changes are ordered youngest -> oldest (MultiTypeTree.java:538-665).
    Leaves are numbered in order of appearance, internal nodes in post-order (root last)."""
    flat = parse_flat_newick(value)
    # strip trailing tiny negative zeros
    leaves, internals = [], []
    heights, parents, types, ctimes, ctypes, names = {}, {}, {}, {}, {}, {}
    counter = [0]

    def build(fnode, parent_id):
        col, tim = [], []
        while len(fnode.children) == 1:
            col.append(_type_index(fnode.meta.get(typeLabel), typeSet))
            tim.append(fnode.height)
            fnode = fnode.children[0]
        col.reverse()
        tim.reverse()
        my = counter[0]
        counter[0] += 1
        heights[my] = fnode.height
        parents[my] = parent_id
        types[my] = _type_index(fnode.meta.get(typeLabel), typeSet)
        ctimes[my], ctypes[my] = tim, col
        names[my] = fnode.label
        if len(fnode.children) == 0:
            leaves.append(my)
        elif len(fnode.children) == 2:
            kids = [build(c, my) for c in fnode.children]
            internals.append(my)       # post-order: children first
        else:
            raise ValueError("multifurcations are not supported")
        return my

    root = build(flat, -1)
    order = leaves + internals          # new Nr -> temp id
    newnr = {tmp: i for i, tmp in enumerate(order)}
    n = len(order)
    height = np.array([heights[t] for t in order], np.float64)
    parent = np.array([newnr[parents[t]] if parents[t] >= 0 else -1 for t in order], np.int32)
    nodeType = np.array([types[t] for t in order], np.uint8)
    assert newnr[root] == n - 1
    return MultiTypeTree(height, parent, nodeType, [list(ctimes[t]) for t in order],
                         [list(ctypes[t]) for t in order],
                         [names[t] for t in order[:len(leaves)]])


def read_nexus_tree(path: str) -> str:
    """First ``tree NAME = ...;`` Newick string of a NEXUS file."""
    return read_nexus_trees(path)[0]


def read_nexus_trees(path: str) -> List[str]:
    """Every ``tree NAME = ...;`` Newick string of a NEXUS tree log, in file order."""
    txt = open(path).read()
    out = [m.group(1).strip() for m in
           re.finditer(r"tree\s+\S+\s*=\s*(?:\[&R\]\s*)?(.*?;)", txt, re.S | re.I)]
    if not out:
        raise ValueError("no tree statement in %s" % path)
    return out


# ---------------------------------------------------------------------------- simulation

def simulate_trees(B: int, nLeaves: int, d: int, m: np.ndarray, N: np.ndarray,
                   leafTypes=None, leafTimes=None, seed: int = 1, nThreads: int = 8) -> TreeBatch:
    """Synthetic structured-coalescent trees (host-side input generator; the semantics of
    evolution/tree/StructuredCoalescentMultiTypeTree.java:191-442, see csrc/host/simulator.cpp).
    Defaults follow SURVEY 8d: leaf type i mod d, leaf height 0.5 * (i mod 4)."""
    L = _lib.lib()
    m = _lib.as_c(m, np.float64).reshape(d, d)
    N = _lib.as_c(N, np.float64)
    if leafTypes is None:
        leafTypes = np.arange(nLeaves) % d
    if leafTimes is None:
        leafTimes = 0.5 * (np.arange(nLeaves) % 4)
    lt = _lib.as_c(leafTypes, np.uint8)
    ltm = _lib.as_c(leafTimes, np.float64)
    h = L.mtt_sim_run(B, nLeaves, d, _lib.ptr(m), _lib.ptr(N), _lib.ptr(lt), _lib.ptr(ltm),
                      seed, nThreads)
    if not h:
        raise ValueError("mtt_sim_run: bad arguments")
    try:
        nNodes = 2 * nLeaves - 1
        tot = L.mtt_sim_total_changes(h)
        tb = TreeBatch(np.zeros(B + 1, np.int64), np.zeros(B * nNodes, np.float64),
                       np.zeros(B * nNodes, np.int32), np.zeros(B * nNodes, np.uint8),
                       np.zeros(B * nNodes, np.uint16), np.zeros(B + 1, np.int64),
                       np.zeros(tot, np.float64), np.zeros(tot, np.uint8))
        owners = np.zeros(tot, np.uint16)
        L.mtt_sim_export(h, _lib.ptr(tb.nodeOff), _lib.ptr(tb.height), _lib.ptr(tb.parent),
                         _lib.ptr(tb.nodeType), _lib.ptr(tb.chgCount), _lib.ptr(tb.chgBase),
                         _lib.ptr(tb.chgTime), _lib.ptr(tb.chgType), _lib.ptr(owners))
        tb.owners = owners
    finally:
        L.mtt_sim_free(h)
    return tb


class DeviceTreeBatch:
    """B trees simulated ON the device (mtt_sim_trees_create, csrc/k5_simulate.cu; SURVEY 8f
    rank 4): the flattened tree SoA in the owner layout, resident in HBM.  ``pointers()`` gives
    the device addresses for the ``*_dev`` entries, ``to_host()`` a TreeBatch."""

    def __init__(self, B: int, nLeaves: int, d: int, m: np.ndarray, N: np.ndarray,
                 leafTypes=None, leafTimes=None, seed: int = 1, ctx=None):
        self.ctx = ctx or _lib.default_context()
        self.B, self.nLeaves, self.d = int(B), int(nLeaves), int(d)
        m = _lib.as_c(m, np.float64).reshape(d, d)
        N = _lib.as_c(N, np.float64)
        if leafTypes is None:
            leafTypes = np.arange(nLeaves) % d
        if leafTimes is None:
            leafTimes = 0.5 * (np.arange(nLeaves) % 4)
        lt = _lib.as_c(leafTypes, np.uint8)
        ltm = _lib.as_c(leafTimes, np.float64)
        self._h = _lib.VP()
        self.ctx.check(_lib.lib().mtt_sim_trees_create(
            self.ctx.handle, self.B, self.nLeaves, self.d, _lib.ptr(m), _lib.ptr(N), _lib.ptr(lt),
            _lib.ptr(ltm), seed, C.byref(self._h)))
        self.total_changes = int(_lib.lib().mtt_sim_trees_total_changes(self._h))

    def pointers(self) -> dict:
        names = ("nodeOff", "height", "parent", "nodeType", "chgNode", "chgBase", "chgTime",
                 "chgType")
        out = [_lib.VP() for _ in names]
        self.ctx.check(_lib.lib().mtt_sim_trees_device(self._h, *[C.byref(x) for x in out]))
        return {k: v.value for k, v in zip(names, out)}

    def to_host(self) -> "TreeBatch":
        B, nN, tot = self.B, 2 * self.nLeaves - 1, self.total_changes
        tb = TreeBatch(np.zeros(B + 1, np.int64), np.zeros(B * nN, np.float64),
                       np.zeros(B * nN, np.int32), np.zeros(B * nN, np.uint8),
                       np.zeros(B * nN, np.uint16), np.zeros(B + 1, np.int64),
                       np.zeros(tot, np.float64), np.zeros(tot, np.uint8))
        owners = np.zeros(tot, np.uint16)
        self.ctx.check(_lib.lib().mtt_sim_trees_export(
            self._h, _lib.ptr(tb.nodeOff), _lib.ptr(tb.height), _lib.ptr(tb.parent),
            _lib.ptr(tb.nodeType), _lib.ptr(owners), _lib.ptr(tb.chgBase), _lib.ptr(tb.chgTime),
            _lib.ptr(tb.chgType)))
        tb.set_owner_list(owners)
        if tot:                          # per-node change counts from the owner list
            tree = np.searchsorted(tb.chgBase, np.arange(tot), side="right") - 1
            np.add.at(tb.chgCount, tree * nN + owners.astype(np.int64), 1)
        return tb

    def close(self):
        if self._h:
            _lib.lib().mtt_sim_trees_destroy(self._h)
            self._h = _lib.VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def synthetic_model(d: int):
    """Migration model of the synthetic workloads: N_c = 1 + 0.25 c,
    m(c,c') = (1 + 0.1 ((c d + c') mod 3)) / (d - 1)."""
    N = 1.0 + 0.25 * np.arange(d)
    m = np.zeros((d, d))
    for c in range(d):
        for cp in range(d):
            if c != cp:
                m[c, cp] = (1.0 + 0.1 * ((c * d + cp) % 3)) / (d - 1)
    return m, N
