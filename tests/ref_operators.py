"""TEST INFRASTRUCTURE (not shipped): the reference's tree operators restated on a BEAST-style
object graph (Node objects with ordered child lists) on top of the CPU oracle's retypeBranch /
getBranchTypeProb.  Used to check multitypetree_b200/operators.py (array-based, device-backed)
proposal by proposal under the same Randomizer seed.

Follows operators/MultiTypeTreeOperator.java:66-248, UniformizationRetypeOperator.java:116-275,
NodeRetype.java:31-63, NodeShiftRetype.java:43-143, TypedSubtreeExchange.java:36-89,
TypedWilsonBalding.java:40-293, MultiTypeTreeScale.java:146-217, MultiTypeUniform.java:46-121.
"""
import math

import numpy as np

from multitypetree_b200.tree import MultiTypeTree

NEG_INF = float("-inf")


class NoValidPath(Exception):
    pass


class Node:
    def __init__(self, nr):
        self.nr = nr
        self.height = 0.0
        self.type = 0
        self.parent = None
        self.children = []
        self.ctimes = []
        self.ctypes = []

    # beast.base.evolution.tree.Node ([EXT])
    def left(self):
        return self.children[0] if len(self.children) > 0 else None

    def right(self):
        return self.children[1] if len(self.children) > 1 else None

    def is_root(self):
        return self.parent is None

    def is_leaf(self):
        return len(self.children) == 0

    def remove_child(self, c):
        self.children.remove(c)

    def add_child(self, c):
        c.parent = self
        self.children.append(c)

    def set_left(self, c):
        if len(self.children) == 0:
            self.children.append(c)
        else:
            self.children[0] = c

    def set_right(self, c):
        if len(self.children) == 0:
            self.children.append(None)
        if len(self.children) == 1:
            self.children.append(c)
        else:
            self.children[1] = c

    # MultiTypeNode
    def final_type(self):
        return self.ctypes[-1] if self.ctypes else self.type

    def final_change_time(self):
        return self.ctimes[-1] if self.ctimes else self.height

    def add_change(self, ty, tm):
        self.ctypes.append(int(ty))
        self.ctimes.append(float(tm))

    def clear_changes(self):
        self.ctypes, self.ctimes = [], []


class RefTree:
    def __init__(self, t: MultiTypeTree):
        n = t.getNodeCount()
        self.nodes = [Node(i) for i in range(n)]
        self.nLeaves = t.getLeafNodeCount()
        for i, nd in enumerate(self.nodes):
            nd.height = float(t.height[i])
            nd.type = int(t.nodeType[i])
            nd.ctimes = list(t.changeTimes[i])
            nd.ctypes = list(t.changeTypes[i])
        for i in range(n):
            for c in t.getChildren(i):
                self.nodes[i].children.append(self.nodes[c])
                self.nodes[c].parent = self.nodes[i]

    def export(self) -> MultiTypeTree:
        n = len(self.nodes)
        t = MultiTypeTree(np.array([nd.height for nd in self.nodes]),
                          np.array([nd.parent.nr if nd.parent else -1 for nd in self.nodes],
                                   np.int32),
                          np.array([nd.type for nd in self.nodes], np.uint8),
                          [list(nd.ctimes) for nd in self.nodes],
                          [list(nd.ctypes) for nd in self.nodes])
        t.left = np.array([nd.left().nr if nd.left() else -1 for nd in self.nodes], np.int32)
        t.right = np.array([nd.right().nr if nd.right() else -1 for nd in self.nodes], np.int32)
        return t

    def root(self):
        return [nd for nd in self.nodes if nd.parent is None][0]


class RefOperators:
    """All operators over one RefTree, one oracle UnifModel and one Randomizer."""

    def __init__(self, tree: RefTree, um, d, rng):
        self.tree, self.um, self.d, self.rng = tree, um, d, rng

    # ---- UniformizationRetypeOperator
    def branch_prob(self, node):
        p = node.parent
        _, lp = self.um.branch_type_prob(node.height, p.height, node.type, node.ctimes,
                                         node.ctypes, p.type)
        return lp

    def retype(self, node):
        p = node.parent
        window = 256
        forced = -1
        while True:
            u = self.rng.peekDoubles(window)
            r = self.um.retype_branch(node.height, p.height, node.type, p.type, u,
                                      nVirtForced=forced, cap=window)
            if r["status"] == 5:        # need more uniforms
                window *= 4
                continue
            if r["status"] == 4:        # steady-state Poisson tail (:96-102)
                self.rng.skipDoubles(r["uUsed"])
                muL = self.um.mu * (p.height - node.height)
                while True:
                    forced = self.rng.nextPoisson(muL)
                    if forced >= r["nVirt"]:
                        break
                continue
            break
        self.rng.skipDoubles(r["uUsed"])
        if r["status"] == 1:
            raise NoValidPath()
        if r["status"] != 0:
            return NEG_INF
        node.clear_changes()
        for k in range(r["nChg"]):
            node.add_change(r["chgType"][k], r["chgTime"][k])
        return r["logProb"]

    # ---- MultiTypeTreeOperator
    @staticmethod
    def other_child(parent, child):
        return parent.right() if parent.left().nr == child.nr else parent.left()

    @staticmethod
    def replace(node, child, replacement):
        node.remove_child(child)
        node.add_child(replacement)

    def disconnect_branch(self, node):
        parent = node.parent
        sister = self.other_child(parent, node)
        for ty, tm in zip(parent.ctypes, parent.ctimes):
            sister.add_change(ty, tm)
        self.replace(parent.parent, parent, sister)
        parent.clear_changes()

    def disconnect_branch_from_root(self, node):
        parent = node.parent
        sister = self.other_child(parent, node)
        sister.parent = None
        parent.remove_child(sister)
        sister.clear_changes()

    def connect_branch(self, node, dest, destTime):
        parent = node.parent
        parent.height = destTime
        split = 0
        while split < len(dest.ctimes):
            if dest.ctimes[split] > destTime:
                break
            split += 1
        parent.clear_changes()
        for idx in range(split, len(dest.ctimes)):
            parent.add_change(dest.ctypes[idx], dest.ctimes[idx])
        del dest.ctimes[split:]
        del dest.ctypes[split:]
        parent.type = dest.final_type()
        self.replace(dest.parent, dest, parent)
        dest.parent = parent
        if parent.left() is node:
            parent.set_right(dest)
        elif parent.right() is node:
            parent.set_left(dest)

    def connect_branch_to_root(self, node, oldRoot, destTime):
        newRoot = node.parent
        newRoot.height = destTime
        newRoot.parent = None
        if newRoot.left() is node:
            newRoot.set_right(oldRoot)
        elif newRoot.right() is node:
            newRoot.set_left(oldRoot)
        oldRoot.parent = newRoot

    # ---- operators
    def node_retype(self):
        T = self.tree
        logHR = 0.0
        node = T.nodes[T.nLeaves + self.rng.nextInt(len(T.nodes) - T.nLeaves)]
        if not node.is_root():
            logHR += self.branch_prob(node)
        logHR += self.branch_prob(node.left()) + self.branch_prob(node.right())
        node.type = self.rng.nextInt(self.d)
        try:
            if not node.is_root():
                logHR -= self.retype(node)
            logHR -= self.retype(node.left()) + self.retype(node.right())
        except NoValidPath:
            return NEG_INF
        return logHR

    def node_shift_retype(self, rootOnly=False, noRoot=False, rootScaleFactor=0.8):
        T = self.tree
        if rootOnly:
            node = T.root()
        else:
            while True:
                node = T.nodes[T.nLeaves + self.rng.nextInt(len(T.nodes) - T.nLeaves)]
                if not (noRoot and node.is_root()):
                    break
        logHR = 0.0
        if node.is_root():
            logHR += self.branch_prob(node.left()) + self.branch_prob(node.right())
            u = self.rng.nextDouble()
            f = u * rootScaleFactor + (1 - u) / rootScaleFactor
            oldest = max(node.left().height, node.right().height)
            node.height = oldest + f * (node.height - oldest)
            logHR -= math.log(f)
            node.type = self.rng.nextInt(self.d)
            try:
                logHR -= self.retype(node.left()) + self.retype(node.right())
            except NoValidPath:
                return NEG_INF
            return logHR
        logHR += (self.branch_prob(node) + self.branch_prob(node.left())
                  + self.branch_prob(node.right()))
        upper = node.parent.height
        lower = max(node.left().height, node.right().height)
        node.height = lower + (upper - lower) * self.rng.nextDouble()
        node.type = self.rng.nextInt(self.d)
        try:
            logHR -= self.retype(node) + self.retype(node.left()) + self.retype(node.right())
        except NoValidPath:
            return NEG_INF
        return logHR

    def typed_subtree_exchange(self, isNarrow=True):
        T = self.tree
        n = len(T.nodes)
        logHR = 0.0
        if isNarrow:
            while True:
                src = T.nodes[self.rng.nextInt(n)]
                if not (src.is_root() or src.parent.is_root()):
                    break
            srcP = src.parent
            dest = self.other_child(srcP.parent, srcP)
            destP = dest.parent
        else:
            while True:
                src = T.nodes[self.rng.nextInt(n)]
                if not src.is_root():
                    break
            srcP = src.parent
            while True:
                dest = T.nodes[self.rng.nextInt(n)]
                if not (dest is src or dest.is_root() or dest.parent is src.parent):
                    break
            destP = dest.parent
        if dest.height > srcP.height or src.height > destP.height:
            return NEG_INF
        logHR += self.branch_prob(src) + self.branch_prob(dest)
        self.replace(srcP, src, dest)
        self.replace(destP, dest, src)
        try:
            logHR -= self.retype(src) + self.retype(dest)
        except NoValidPath:
            return NEG_INF
        return logHR

    def _invalid_src(self, src):
        if src.is_root():
            return True
        parent = src.parent
        if parent.is_root():
            sister = self.other_child(parent, src)
            if sister.is_leaf():
                return True
            if src.height >= sister.height:
                return True
        return False

    @staticmethod
    def _invalid_dest(src, dest):
        if dest is src or dest is src.parent or dest.parent is src.parent:
            return True
        destP = dest.parent
        if destP is not None and destP.height <= src.height:
            return True
        return False

    def typed_wilson_balding(self, alpha):
        T = self.tree
        n = len(T.nodes)
        while True:
            src = T.nodes[self.rng.nextInt(n)]
            if not self._invalid_src(src):
                break
        srcP = src.parent
        srcS = self.other_child(srcP, src)
        t_src, t_srcP, t_srcS = src.height, srcP.height, srcS.height
        while True:
            dest = T.nodes[self.rng.nextInt(n)]
            if not self._invalid_dest(src, dest):
                break
        destP = dest.parent
        t_dest = dest.height

        if dest.is_root():
            logHR = 0.0
            logHR += self.branch_prob(src)
            t_srcG = srcP.parent.height
            newTime = t_dest + self.rng.nextExponential(1.0 / (alpha * t_dest))
            self.disconnect_branch(src)
            self.connect_branch_to_root(src, dest, newTime)
            try:
                # retypeRootBranches
                logProb = 0.0
                sP = src.parent
                sS = self.other_child(sP, src)
                sP.type = self.rng.nextInt(self.d)
                logProb += math.log(1.0 / self.d)
                logProb += self.retype(src)
                logProb += self.retype(sS)
                logHR -= logProb
            except NoValidPath:
                return NEG_INF
            logHR += (math.log(alpha * t_dest) + (1.0 / alpha) * (newTime / t_dest - 1.0)
                      - math.log(t_srcG - max(t_src, t_srcS)))
            return logHR

        if srcP.is_root():
            logHR = 0.0
            logProb = 0.0
            logProb += math.log(1.0 / self.d)
            logProb += self.branch_prob(src)
            logProb += self.branch_prob(srcS)
            logHR += logProb
            oldTime = t_srcP
            min_newTime = max(t_src, t_dest)
            t_destP = destP.height
            span = t_destP - min_newTime
            newTime = min_newTime + span * self.rng.nextDouble()
            self.disconnect_branch_from_root(src)
            self.connect_branch(src, dest, newTime)
            srcS.parent = None
            try:
                logHR -= self.retype(src)
            except NoValidPath:
                return NEG_INF
            logHR += (math.log(t_destP - max(t_src, t_dest)) - math.log(alpha * t_srcS)
                      - (1.0 / alpha) * (oldTime / t_srcS - 1.0))
            return logHR

        logHR = 0.0
        logHR += self.branch_prob(src)
        t_srcG = srcP.parent.height
        min_newTime = max(t_dest, t_src)
        t_destP = destP.height
        span = t_destP - min_newTime
        newTime = min_newTime + span * self.rng.nextDouble()
        self.disconnect_branch(src)
        self.connect_branch(src, dest, newTime)
        try:
            logHR -= self.retype(src)
        except NoValidPath:
            return NEG_INF
        logHR += (math.log(t_destP - max(t_src, t_dest))
                  - math.log(t_srcG - max(t_src, t_srcS)))
        return logHR

    def tree_scale(self, scaleFactor, useOldTreeScaler=False):
        T = self.tree
        u = self.rng.nextDouble()
        f = u * scaleFactor + (1.0 - u) / scaleFactor
        logf = math.log(f)
        logHR = -2 * logf
        leaves, internals = T.nodes[:T.nLeaves], T.nodes[T.nLeaves:]
        for leaf in leaves:
            if not useOldTreeScaler:
                lold = leaf.parent.height - leaf.height
                lnew = f * leaf.parent.height - leaf.height
                for c in range(len(leaf.ctimes)):
                    leaf.ctimes[c] = leaf.height + (leaf.ctimes[c] - leaf.height) * lnew / lold
                ratio = lnew / lold
                logHR += len(leaf.ctimes) * (math.log(ratio) if ratio > 0 else math.nan)
            else:
                for c in range(len(leaf.ctimes)):
                    leaf.ctimes[c] = f * leaf.ctimes[c]
                    logHR += logf
        for node in internals:
            node.height = node.height * f
            logHR += logf
            for c in range(len(node.ctimes)):
                node.ctimes[c] = f * node.ctimes[c]
                logHR += logf
        if f < 1.0:
            for leaf in leaves:
                if leaf.parent.height < leaf.height:
                    return NEG_INF
                if leaf.ctimes and leaf.ctimes[0] < leaf.height:
                    if not useOldTreeScaler:
                        raise RuntimeError("change time below leaf")
                    return NEG_INF
        return logHR

    def mt_uniform(self, includeRoot=False, rootScaleFactor=0.9):
        T = self.tree
        nInt = len(T.nodes) - T.nLeaves
        total = sum(len(nd.ctimes) for nd in T.nodes if not nd.is_root())
        event = self.rng.nextInt(nInt + total)
        node, changeIdx = None, -1
        if event < nInt:
            node = T.nodes[T.nLeaves + event]
            if not includeRoot and node.is_root():
                return NEG_INF
        else:
            event -= nInt
            for nd in T.nodes:
                if nd.is_root():
                    continue
                if event < len(nd.ctimes):
                    node, changeIdx = nd, event
                    break
                event -= len(nd.ctimes)
        if changeIdx == -1:
            tmin = max(node.left().final_change_time(), node.right().final_change_time())
            if node.is_root():
                u = self.rng.nextDouble()
                f = u * rootScaleFactor + (1 - u) / rootScaleFactor
                node.height = tmin + f * (node.height - tmin)
                return -math.log(f)
            tmax = node.ctimes[0] if node.ctimes else node.parent.height
            u = self.rng.nextDouble()
            node.height = u * tmin + (1.0 - u) * tmax
            return 0.0
        tmax = node.ctimes[changeIdx + 1] if changeIdx + 1 < len(node.ctimes) \
            else node.parent.height
        tmin = node.height if changeIdx - 1 < 0 else node.ctimes[changeIdx - 1]
        u = self.rng.nextDouble()
        node.ctimes[changeIdx] = u * tmin + (1 - u) * tmax
        return 0.0
