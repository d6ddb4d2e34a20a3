"""Host-side mirrors of the reference's uniformization tree operators.

    MultiTypeTreeOperator          operators/MultiTypeTreeOperator.java:35-248
    UniformizationRetypeOperator   operators/UniformizationRetypeOperator.java:36-275
    NodeRetype                     operators/NodeRetype.java:31-63
    NodeShiftRetype                operators/NodeShiftRetype.java:43-143
    TypedSubtreeExchange           operators/TypedSubtreeExchange.java:36-89
    TypedWilsonBalding             operators/TypedWilsonBalding.java:40-293
    MultiTypeTreeScale             operators/MultiTypeTreeScale.java:146-217 (tree part)
    MultiTypeUniform               operators/MultiTypeUniform.java:46-121

What stays on the host is what stays in Java in the real plugin: picking nodes / heights /
types with the ``Randomizer`` and editing the topology.  Every ``retypeBranch`` and
``getBranchTypeProb`` is answered by the uniformization kernels (csrc/k3_unif.cu) through
``mtt_unif_retype_batch`` / ``mtt_unif_branch_logp_batch``.

Each operator's proposal is written once, as a coroutine (``steps()``) that *yields* its
device requests.  ``proposal()`` drives one coroutine; ``run_proposals(ops)`` drives many
chains in lock step and merges their requests into ONE batched C-ABI call per round -- the
shape the B200 path is built for (64 chains).  The Randomizer stream of
every chain is consumed exactly in the reference's order either way.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np

from .tree import MultiTypeTree
from .uniformization import (Randomizer, ST_NEED_MORE_UNIFORMS, ST_NEEDS_POISSON_TAIL,
                             ST_NO_VALID_PATH, ST_OK)

NEG_INF = float("-inf")


class _Prob:
    """request: getBranchTypeProb of the branches above `nodes` -> list of log-probabilities"""
    __slots__ = ("nodes",)

    def __init__(self, nodes):
        self.nodes = [int(n) for n in nodes]


class _Retype:
    """request: retypeBranch of the branches above `nodes`, in order, from the chain's uniform
    stream -> list of log-probabilities, or None after a NoValidPathException"""
    __slots__ = ("nodes",)

    def __init__(self, nodes):
        self.nodes = [int(n) for n in nodes]


# ---------------------------------------------------------------------------------------
class MultiTypeTreeOperator:
    """operators/MultiTypeTreeOperator.java: Inputs ``multiTypeTree``, ``migrationModel``
    (+ BEAST's ``weight``) and the branch surgery shared by the topology operators."""

    def __init__(self, multiTypeTree: MultiTypeTree, migrationModel, weight: float = 1.0,
                 randomizer: Optional[Randomizer] = None, ctx=None):
        self.mtTree = multiTypeTree
        self.migModel = migrationModel
        self.weight = float(weight)
        self.rng = randomizer or Randomizer(127)
        self.ctx = ctx

    # -- MultiTypeTreeOperator.java:66-89
    def getOtherChild(self, parent: int, child: int) -> int:
        t = self.mtTree
        return t.getRight(parent) if t.getLeft(parent) == child else t.getLeft(parent)

    def replace(self, node: int, child: int, replacement: int) -> None:
        self.mtTree.removeChild(node, child)
        self.mtTree.addChild(node, replacement)

    # -- :100-129  the sister inherits the parent's changes and hangs off the grandmother
    def disconnectBranch(self, node: int) -> None:
        t = self.mtTree
        parent = t.getParent(node)
        if t.isRoot(node) or t.isRoot(parent):
            raise ValueError("Illegal argument to disconnectBranch().")
        sister = self.getOtherChild(parent, node)
        for idx in range(t.getChangeCount(parent)):
            t.addChange(sister, t.getChangeType(parent, idx), t.getChangeTime(parent, idx))
        self.replace(t.getParent(parent), parent, sister)
        t.clearChanges(parent)

    # -- :137-158
    def disconnectBranchFromRoot(self, node: int) -> None:
        t = self.mtTree
        if t.isRoot(node) or not t.isRoot(t.getParent(node)):
            raise ValueError("Illegal argument to disconnectBranchFromRoot().")
        parent = t.getParent(node)
        sister = self.getOtherChild(parent, node)
        t.parent[sister] = -1
        t.removeChild(parent, sister)
        t.clearChanges(sister)

    # -- :169-216  split the destination branch at destTime
    def connectBranch(self, node: int, destBranchBase: int, destTime: float) -> None:
        t = self.mtTree
        if t.isRoot(node) or t.isRoot(destBranchBase):
            raise ValueError("Illegal argument to connectBranch().")
        parent = t.getParent(node)
        t.setHeight(parent, destTime)
        n = t.getChangeCount(destBranchBase)
        split = 0
        while split < n and not t.getChangeTime(destBranchBase, split) > destTime:
            split += 1
        t.clearChanges(parent)
        for idx in range(split, n):
            t.addChange(parent, t.getChangeType(destBranchBase, idx),
                        t.getChangeTime(destBranchBase, idx))
        t.truncateChanges(destBranchBase, split)
        t.setNodeType(parent, t.getFinalType(destBranchBase))
        self.replace(t.getParent(destBranchBase), destBranchBase, parent)
        t.parent[destBranchBase] = parent
        if t.getLeft(parent) == node:
            t.right[parent] = destBranchBase
        elif t.getRight(parent) == node:
            t.left[parent] = destBranchBase

    # -- :226-248
    def connectBranchToRoot(self, node: int, oldRoot: int, destTime: float) -> None:
        t = self.mtTree
        if t.isRoot(node) or not t.isRoot(oldRoot):
            raise ValueError("Illegal argument to connectBranchToRoot().")
        newRoot = t.getParent(node)
        t.setHeight(newRoot, destTime)
        t.parent[newRoot] = -1
        if t.getLeft(newRoot) == node:
            t.right[newRoot] = oldRoot
        elif t.getRight(newRoot) == node:
            t.left[newRoot] = oldRoot
        t.parent[oldRoot] = newRoot

    # -- proposal protocol
    def steps(self):
        """Coroutine form of ``proposal()``: yields _Prob / _Retype requests, returns logHR."""
        raise NotImplementedError
        yield  # pragma: no cover

    def proposal(self) -> float:
        """Returns the log Hastings ratio or -inf, as ``Operator.proposal()`` does."""
        return run_proposals([self])[0]


class UniformizationRetypeOperator(MultiTypeTreeOperator):
    """operators/UniformizationRetypeOperator.java: adds Input ``useSymmetrizedRates``.
    ``retypeBranch`` (:116-225) and ``getBranchTypeProb`` (:233-275) run on the device."""

    def __init__(self, multiTypeTree, migrationModel, useSymmetrizedRates: bool = False,
                 weight: float = 1.0, randomizer=None, ctx=None, uniform_window: int = 256,
                 model_index: int = 0):
        super().__init__(multiTypeTree, migrationModel, weight, randomizer, ctx)
        self.sym = bool(useSymmetrizedRates)
        self.window = int(uniform_window)
        self.model_index = int(model_index)     # row of a multi-model UniformizationModels

    def _models(self):
        return self.migModel._ensure_device(self.sym)


# ---------------------------------------------------------------------------------------
class NodeRetype(UniformizationRetypeOperator):
    """operators/NodeRetype.java:31-63."""

    def steps(self):
        t = self.mtTree
        logHR = 0.0
        node = t.getLeafNodeCount() + self.rng.nextInt(t.getInternalNodeCount())   # :36-37
        left, right = t.getLeft(node), t.getRight(node)
        root = t.isRoot(node)
        # probability of the current types along the attached branches (:40-44)
        if not root:
            p = yield _Prob([node, left, right])
            logHR += p[0]
            logHR += p[1] + p[2]
        else:
            p = yield _Prob([left, right])
            logHR += p[0] + p[1]
        t.setNodeType(node, self.rng.nextInt(self.migModel.getNTypes()))           # :47-48
        # retype the attached branches, node -> left -> right (:52-56)
        if not root:
            q = yield _Retype([node, left, right])
            if q is None:
                return NEG_INF                                                     # :57-58
            logHR -= q[0]
            logHR -= q[1] + q[2]
        else:
            q = yield _Retype([left, right])
            if q is None:
                return NEG_INF
            logHR -= q[0] + q[1]
        return logHR


class NodeShiftRetype(UniformizationRetypeOperator):
    """operators/NodeShiftRetype.java: Inputs ``rootOnly``, ``noRoot``, ``rootScaleFactor``."""

    def __init__(self, multiTypeTree, migrationModel, rootOnly: bool = False,
                 noRoot: bool = False, rootScaleFactor: float = 0.8, **kw):
        super().__init__(multiTypeTree, migrationModel, **kw)
        if rootOnly and noRoot:                                                    # :58-60
            raise ValueError("rootOnly and noRoot inputs cannot both be set to true "
                             "simultaneously.")
        self.rootOnly, self.noRoot = bool(rootOnly), bool(noRoot)
        self.rootScaleFactor = float(rootScaleFactor)

    def steps(self):
        t = self.mtTree
        if self.rootOnly:                                                          # :66-73
            node = t.getRoot()
        else:
            while True:
                node = t.getLeafNodeCount() + self.rng.nextInt(t.getInternalNodeCount())
                if not (self.noRoot and t.isRoot(node)):
                    break
        left, right = t.getLeft(node), t.getRight(node)
        logHR = 0.0
        if t.isRoot(node):                                                         # :88-117
            p = yield _Prob([left, right])
            logHR += p[0] + p[1]
            u = self.rng.nextDouble()
            f = u * self.rootScaleFactor + (1 - u) / self.rootScaleFactor
            oldest = max(t.getHeight(left), t.getHeight(right))
            t.setHeight(node, oldest + f * (t.getHeight(node) - oldest))
            logHR -= math.log(f)
            t.setNodeType(node, self.rng.nextInt(self.migModel.getNTypes()))
            q = yield _Retype([left, right])
            if q is None:
                return NEG_INF
            logHR -= q[0] + q[1]
            return logHR
        p = yield _Prob([node, left, right])                                       # :124-143
        logHR += p[0] + p[1] + p[2]
        upper = t.getHeight(t.getParent(node))
        lower = max(t.getHeight(left), t.getHeight(right))
        t.setHeight(node, lower + (upper - lower) * self.rng.nextDouble())
        t.setNodeType(node, self.rng.nextInt(self.migModel.getNTypes()))
        q = yield _Retype([node, left, right])
        if q is None:
            return NEG_INF
        logHR -= q[0] + q[1] + q[2]
        return logHR


class TypedSubtreeExchange(UniformizationRetypeOperator):
    """operators/TypedSubtreeExchange.java: Input ``isNarrow``.

    Behaviour kept as in the reference, including that the WIDE variant does not exclude
    dest == parent(src) / src == parent(dest) (:64-69; the height test at :74-76 lets the
    equal-height case through), which detaches a subtree.  The narrow variant (the default,
    and the one the shipped XMLs use) cannot do this."""

    def __init__(self, multiTypeTree, migrationModel, isNarrow: bool = True, **kw):
        super().__init__(multiTypeTree, migrationModel, **kw)
        self.isNarrow = bool(isNarrow)

    def steps(self):
        t = self.mtTree
        n = t.getNodeCount()
        logHR = 0.0
        if self.isNarrow:                                                          # :47-55
            while True:
                src = self.rng.nextInt(n)
                if not (t.isRoot(src) or t.isRoot(t.getParent(src))):
                    break
            srcP = t.getParent(src)
            dest = self.getOtherChild(t.getParent(srcP), srcP)
            destP = t.getParent(dest)
        else:                                                                      # :59-70
            while True:
                src = self.rng.nextInt(n)
                if not t.isRoot(src):
                    break
            srcP = t.getParent(src)
            while True:
                dest = self.rng.nextInt(n)
                if not (dest == src or t.isRoot(dest) or t.getParent(dest) == srcP):
                    break
            destP = t.getParent(dest)
        # negative branch lengths (:74-76)
        if t.getHeight(dest) > t.getHeight(srcP) or t.getHeight(src) > t.getHeight(destP):
            return NEG_INF
        p = yield _Prob([src, dest])                                               # :79
        logHR += p[0] + p[1]
        self.replace(srcP, src, dest)                                              # :82-83
        self.replace(destP, dest, src)
        q = yield _Retype([src, dest])                                             # :86-90
        if q is None:
            return NEG_INF
        logHR -= q[0] + q[1]
        return logHR


class TypedWilsonBalding(UniformizationRetypeOperator):
    """operators/TypedWilsonBalding.java: Input ``alpha`` (root height proposal parameter)."""

    def __init__(self, multiTypeTree, migrationModel, alpha: float, **kw):
        super().__init__(multiTypeTree, migrationModel, **kw)
        self.alpha = float(alpha)

    def _invalidSrcNode(self, src: int) -> bool:                                   # :197-216
        t = self.mtTree
        if t.isRoot(src):
            return True
        parent = t.getParent(src)
        if t.isRoot(parent):
            sister = self.getOtherChild(parent, src)
            if t.isLeaf(sister):
                return True
            if t.getHeight(src) >= t.getHeight(sister):
                return True
        return False

    def _invalidDestNode(self, src: int, dest: int) -> bool:                       # :226-239
        t = self.mtTree
        if dest == src or dest == t.getParent(src) or t.getParent(dest) == t.getParent(src):
            return True
        destP = t.getParent(dest)
        if destP >= 0 and t.getHeight(destP) <= t.getHeight(src):
            return True
        return False

    def steps(self):
        t = self.mtTree
        alpha = self.alpha
        if t.getLeafNodeCount() < 3:                                               # :57-59
            raise RuntimeError("Tree too small for TypedWilsonBalding operator.")
        n = t.getNodeCount()
        while True:                                                                # :62-65
            src = self.rng.nextInt(n)
            if not self._invalidSrcNode(src):
                break
        srcP = t.getParent(src)
        srcS = self.getOtherChild(srcP, src)
        t_src, t_srcP, t_srcS = t.getHeight(src), t.getHeight(srcP), t.getHeight(srcS)
        while True:                                                                # :73-76
            dest = self.rng.nextInt(n)
            if not self._invalidDestNode(src, dest):
                break
        destP = t.getParent(dest)
        t_dest = t.getHeight(dest)
        dTypes = self.migModel.getNTypes()

        if t.isRoot(dest):                                                         # :82-113
            logHR = 0.0
            p = yield _Prob([src])
            logHR += p[0]
            t_srcG = t.getHeight(t.getParent(srcP))
            newTime = t_dest + self.rng.nextExponential(1.0 / (alpha * t_dest))
            self.disconnectBranch(src)
            self.connectBranchToRoot(src, dest, newTime)
            # retypeRootBranches (:249-269): new root type, then both root branches
            t.setNodeType(srcP, self.rng.nextInt(dTypes))
            logProb = 0.0
            logProb += math.log(1.0 / dTypes)
            q = yield _Retype([src, self.getOtherChild(srcP, src)])   # sister is now the old root
            if q is None:
                return NEG_INF
            logProb += q[0]
            logProb += q[1]
            logHR -= logProb
            logHR += (math.log(alpha * t_dest) + (1.0 / alpha) * (newTime / t_dest - 1.0)
                      - math.log(t_srcG - max(t_src, t_srcS)))
            return logHR

        if t.isRoot(srcP):                                                         # :115-152
            logHR = 0.0
            # getRootBranchTypeProb (:281-292)
            p = yield _Prob([src, srcS])
            logProb = 0.0
            logProb += math.log(1.0 / dTypes)
            logProb += p[0]
            logProb += p[1]
            logHR += logProb
            oldTime = t_srcP
            min_newTime = max(t_src, t_dest)
            t_destP = t.getHeight(destP)
            span = t_destP - min_newTime
            newTime = min_newTime + span * self.rng.nextDouble()
            self.disconnectBranchFromRoot(src)
            self.connectBranch(src, dest, newTime)
            t.parent[srcS] = -1
            q = yield _Retype([src])
            if q is None:
                return NEG_INF
            logHR -= q[0]
            logHR += (math.log(t_destP - max(t_src, t_dest)) - math.log(alpha * t_srcS)
                      - (1.0 / alpha) * (oldTime / t_srcS - 1.0))
            return logHR

        logHR = 0.0                                                                # :154-187
        p = yield _Prob([src])
        logHR += p[0]
        t_srcG = t.getHeight(t.getParent(srcP))
        min_newTime = max(t_dest, t_src)
        t_destP = t.getHeight(destP)
        span = t_destP - min_newTime
        newTime = min_newTime + span * self.rng.nextDouble()
        self.disconnectBranch(src)
        self.connectBranch(src, dest, newTime)
        q = yield _Retype([src])
        if q is None:
            return NEG_INF
        logHR -= q[0]
        logHR += (math.log(t_destP - max(t_src, t_dest))
                  - math.log(t_srcG - max(t_src, t_srcS)))
        return logHR


# --------------------------------------------------------------------------------------
# operators without device work (needed to run the reference's operator tests end to end)
class MultiTypeTreeScale(MultiTypeTreeOperator):
    """operators/MultiTypeTreeScale.java:146-217, tree part: Inputs ``scaleFactor``,
    ``useOldTreeScaler``.  (The optional ``parameter`` / ``parameterInverse`` lists scale BEAST
    RealParameters and are outside this path.)"""

    def __init__(self, multiTypeTree, migrationModel, scaleFactor: float,
                 useOldTreeScaler: bool = False, **kw):
        super().__init__(multiTypeTree, migrationModel, **kw)
        self.scaleFactor = float(scaleFactor)
        self.useOldTreeScaler = bool(useOldTreeScaler)

    def steps(self):
        t = self.mtTree
        u = self.rng.nextDouble()
        f = u * self.scaleFactor + (1.0 - u) / self.scaleFactor
        logf = math.log(f)
        logHR = -2 * logf
        nLeaves = t.getLeafNodeCount()
        for leaf in range(nLeaves):                                                # :159-181
            hp, hl = t.getHeight(t.getParent(leaf)), t.getHeight(leaf)
            if not self.useOldTreeScaler:
                lold = hp - hl
                lnew = f * hp - hl
                for c in range(t.getChangeCount(leaf)):
                    t.setChangeTime(leaf, c, hl + (t.getChangeTime(leaf, c) - hl) * lnew / lold)
                ratio = lnew / lold
                logHR += t.getChangeCount(leaf) * (math.log(ratio) if ratio > 0 else math.nan)
            else:
                for c in range(t.getChangeCount(leaf)):
                    t.setChangeTime(leaf, c, f * t.getChangeTime(leaf, c))
                    logHR += logf
        for node in range(nLeaves, t.getNodeCount()):                              # :184-194
            t.setHeight(node, t.getHeight(node) * f)
            logHR += logf
            for c in range(t.getChangeCount(node)):
                t.setChangeTime(node, c, f * t.getChangeTime(node, c))
                logHR += logf
        if f < 1.0:                                                                # :197-211
            for leaf in range(nLeaves):
                if t.getHeight(t.getParent(leaf)) < t.getHeight(leaf):
                    return NEG_INF
                if t.getChangeCount(leaf) > 0 and t.getChangeTime(leaf, 0) < t.getHeight(leaf):
                    if not self.useOldTreeScaler:
                        raise RuntimeError("Scaled colour change time has dipped below age "
                                           "of leaf - this should never happen!")
                    return NEG_INF
        return logHR
        yield  # pragma: no cover  (a coroutine without device requests)


class MultiTypeUniform(MultiTypeTreeOperator):
    """operators/MultiTypeUniform.java:46-121: Inputs ``includeRoot``, ``rootScaleFactor``."""

    def __init__(self, multiTypeTree, migrationModel, includeRoot: bool = False,
                 rootScaleFactor: float = 0.9, **kw):
        super().__init__(multiTypeTree, migrationModel, **kw)
        self.includeRoot = bool(includeRoot)
        self.rootScaleFactor = float(rootScaleFactor)

    def _finalChangeTime(self, node: int) -> float:
        t = self.mtTree
        k = t.getChangeCount(node)
        return t.getChangeTime(node, k - 1) if k else t.getHeight(node)

    def steps(self):
        t = self.mtTree
        nInt = t.getInternalNodeCount()
        root = t.getRoot()
        nChanges = t.getTotalNumberOfChanges()
        event = self.rng.nextInt(nInt + nChanges)                                  # :48
        node, changeIdx = None, -1
        if event < nInt:
            node = t.getLeafNodeCount() + event
            if not self.includeRoot and node == root:
                return NEG_INF
        else:
            event -= nInt
            for v in range(t.getNodeCount()):                                      # :59-69
                if v == root:
                    continue
                if event < t.getChangeCount(v):
                    node, changeIdx = v, event
                    break
                event -= t.getChangeCount(v)
        if node is None:
            raise RuntimeError("Event selection loop fell through!")
        if changeIdx == -1:
            tmin = max(self._finalChangeTime(t.getLeft(node)),
                       self._finalChangeTime(t.getRight(node)))
            if node == root:                                                       # :76-88
                u = self.rng.nextDouble()
                f = u * self.rootScaleFactor + (1 - u) / self.rootScaleFactor
                t.setHeight(node, tmin + f * (t.getHeight(node) - tmin))
                return -math.log(f)
            tmax = t.getChangeTime(node, 0) if t.getChangeCount(node) > 0 \
                else t.getHeight(t.getParent(node))                                # :91-101
            u = self.rng.nextDouble()
            t.setHeight(node, u * tmin + (1.0 - u) * tmax)
            return 0.0
        if changeIdx + 1 < t.getChangeCount(node):                                 # :103-118
            tmax = t.getChangeTime(node, changeIdx + 1)
        else:
            tmax = t.getHeight(t.getParent(node))
        tmin = t.getHeight(node) if changeIdx - 1 < 0 else t.getChangeTime(node, changeIdx - 1)
        u = self.rng.nextDouble()
        t.setChangeTime(node, changeIdx, u * tmin + (1 - u) * tmax)
        return 0.0
        yield  # pragma: no cover


# --------------------------------------------------------------------------------------
# the driver: serve the coroutines' requests with batched device calls
class _RetypeJob:
    __slots__ = ("op", "nodes", "start", "forced", "lp", "window", "answer", "done")

    def __init__(self, op, nodes):
        self.op, self.nodes = op, nodes
        self.start = 0
        self.forced = [-1] * len(nodes)
        self.lp = [0.0] * len(nodes)
        self.window = op.window
        self.answer = None
        self.done = False


def _branch(t: MultiTypeTree, node: int):
    p = int(t.parent[node])
    return float(t.height[node]), float(t.height[p]), int(t.nodeType[node]), int(t.nodeType[p])


def _serve_probs(um, reqs) -> Dict[int, List[float]]:
    """one mtt_unif_branch_logp_batch call for every pending getBranchTypeProb"""
    t0, t1, a, b, mdl, off, ct, cty = [], [], [], [], [], [0], [], []
    for _, op, req in reqs:
        t = op.mtTree
        for nd in req.nodes:
            x = _branch(t, nd)
            t0.append(x[0]); t1.append(x[1]); a.append(x[2]); b.append(x[3])
            mdl.append(op.model_index)
            ct.extend(t.changeTimes[nd])
            cty.extend(t.changeTypes[nd])
            off.append(len(ct))
    lp, _ = um.branch_logp(t0, t1, a, b, off, ct, cty, model=mdl)
    out, k = {}, 0
    for i, _, req in reqs:
        out[i] = [float(x) for x in lp[k:k + len(req.nodes)]]
        k += len(req.nodes)
    return out


def _serve_retypes(um, reqs) -> Dict[int, Optional[List[float]]]:
    """mtt_unif_retype_batch for every pending retypeBranch sequence (one job per chain).
    Jobs whose status asks for host action (larger uniform window, the Poisson tail of
    UniformizationRetypeOperator.java:96-102) are resubmitted until every job has finished."""
    jobs = {i: _RetypeJob(op, req.nodes) for i, op, req in reqs}
    pending = list(jobs.values())
    while pending:
        brOff, t0, t1, a, b, mdl, forced, us, uOff = [0], [], [], [], [], [], [], [], [0]
        for j in pending:
            t = j.op.mtTree
            for nd in j.nodes[j.start:]:
                x = _branch(t, nd)
                t0.append(x[0]); t1.append(x[1]); a.append(x[2]); b.append(x[3])
                mdl.append(j.op.model_index)
            forced.extend(j.forced[j.start:])
            brOff.append(len(t0))
            us.append(j.op.rng.peekDoubles(j.window))
            uOff.append(uOff[-1] + j.window)
        outCap = max(j.window for j in pending)
        r = um.retype(brOff, t0, t1, a, b, np.concatenate(us), uOff, outCap=outCap, model=mdl,
                      nVirtForced=forced)
        nxt = []
        for ji, j in enumerate(pending):
            lo, hi = brOff[ji], brOff[ji + 1]
            st = r["status"][lo:hi]
            if np.any(st == ST_NEED_MORE_UNIFORMS):
                j.window *= 4
                nxt.append(j)
                continue
            t = j.op.mtTree
            bad = np.nonzero(st != ST_OK)[0]
            k = int(bad[0]) if len(bad) else hi - lo
            for i in range(k):                       # branches retyped successfully
                nd = j.nodes[j.start + i]
                t.clearChanges(nd)
                for c in range(int(r["nChg"][lo + i])):
                    t.addChange(nd, int(r["chgType"][lo + i, c]), float(r["chgTime"][lo + i, c]))
                j.lp[j.start + i] = float(r["logProb"][lo + i])
            j.op.rng.skipDoubles(int(r["uUsed"][ji]))
            if k == hi - lo:
                j.answer, j.done = j.lp, True
                continue
            if st[k] == ST_NO_VALID_PATH:            # NoValidPathException (:137-140)
                j.answer, j.done = None, True
                continue
            if st[k] == ST_NEEDS_POISSON_TAIL:
                thresh = int(r["nVirt"][lo + k])
                muL = um.matrices(j.op.model_index)[2] * (t1[lo + k] - t0[lo + k])
                while True:
                    nv = j.op.rng.nextPoisson(muL)
                    if nv >= thresh:
                        break
                j.forced[j.start + k] = nv
                j.start += k
                nxt.append(j)
                continue
            # Pba rubbish / forward-backward fell through: retypeBranch returned -Infinity and
            # the proposal carries on with the remaining branches (:141-144, :179-190)
            j.lp[j.start + k] = NEG_INF
            j.start += k + 1
            if j.start >= len(j.nodes):
                j.answer, j.done = j.lp, True
            else:
                nxt.append(j)
        pending = nxt
    return {i: j.answer for i, j in jobs.items()}


def run_proposals(ops: Sequence[MultiTypeTreeOperator]) -> List[float]:
    """Run one proposal of every operator in `ops` (one per chain; each with its own tree and
    Randomizer) in lock step.  Requests of the same kind are merged into one device call per
    round and per device-resident model set."""
    gens = [op.steps() for op in ops]
    results: List[Optional[float]] = [None] * len(ops)
    pending: Dict[int, object] = {}

    def advance(i, value, first=False):
        try:
            pending[i] = next(gens[i]) if first else gens[i].send(value)
        except StopIteration as e:
            results[i] = e.value

    for i in range(len(ops)):
        advance(i, None, first=True)
    while pending:
        groups: Dict[tuple, list] = {}
        for i, req in pending.items():
            um = ops[i]._models()
            groups.setdefault((id(um), type(req)), [um, []])[1].append((i, ops[i], req))
        pending = {}
        answers = {}
        for (_, kind), (um, reqs) in groups.items():
            answers.update(_serve_probs(um, reqs) if kind is _Prob else _serve_retypes(um, reqs))
        for i, ans in answers.items():
            advance(i, ans)
    return results
