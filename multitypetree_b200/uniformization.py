"""Host-side wrappers of the uniformization entry points (csrc/k3_unif.cu) and the mirror of
operators/NodeRetype.java + operators/UniformizationRetypeOperator.java.

The numerical work -- Q/R/mu (SCMigrationModel.java:150-201), the R^k chain (:404-473), expm,
the virtual-event count, forward-backward path sampling and path probabilities
(UniformizationRetypeOperator.java:74-275) -- runs on the device.  What stays on the host is
what stays in Java in the real plugin: choosing the node and its new type with the Randomizer
(NodeRetype.java:36-37, 47-48) and writing the returned changes back into the tree.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from .tree import MultiTypeTree

ST_OK, ST_NO_VALID_PATH, ST_PBA_RUBBISH, ST_FB_FELL_THROUGH = 0, 1, 2, 3
ST_NEEDS_POISSON_TAIL, ST_NEED_MORE_UNIFORMS, ST_NEED_MORE_POWERS, ST_SKIPPED = 4, 5, 6, 7


class Randomizer:
    """BEAST 2 ``Randomizer`` stream (MersenneTwisterFast) restated on the host
    (csrc/host/randomizer.cpp; [EXT], parity with a JVM unpinned)."""

    def __init__(self, seed: int = 127):
        self._h = _lib.lib().mtt_rng_new(seed)

    def __del__(self):
        try:
            _lib.lib().mtt_rng_free(self._h)
        except Exception:
            pass

    def setSeed(self, seed: int):
        _lib.lib().mtt_rng_set_seed(self._h, seed)

    def nextDouble(self) -> float:
        return _lib.lib().mtt_rng_next_double(self._h)

    def nextInt(self, n: int) -> int:
        return _lib.lib().mtt_rng_next_int(self._h, n)

    def nextDoubles(self, n: int) -> np.ndarray:
        out = np.empty(n, np.float64)
        _lib.lib().mtt_rng_next_doubles(self._h, n, _lib.ptr(out))
        return out

    def peekDoubles(self, n: int) -> np.ndarray:
        out = np.empty(n, np.float64)
        _lib.lib().mtt_rng_peek_doubles(self._h, n, _lib.ptr(out))
        return out

    def skipDoubles(self, n: int):
        _lib.lib().mtt_rng_skip_doubles(self._h, n)

    def nextPoisson(self, lam: float) -> int:
        """Poisson draw for the steady-state tail (UniformizationRetypeOperator.java:100-102).
        BEAST's own algorithm for this call is [EXT] and not reproduced: for lam < 40 the
        multiplicative method, otherwise a normal-approximation-free inversion by summation.
        Distributionally exact; NOT stream-compatible with a JVM run."""
        if lam < 40:
            p, n, r = np.exp(-lam), 0, 1.0
            while True:
                r *= self.nextDouble()
                if r < p:
                    return n
                n += 1
        u, n = self.nextDouble(), 0
        logp = -lam
        cdf = np.exp(logp)
        while cdf < u and n < 10 * lam + 100:
            n += 1
            logp += np.log(lam) - np.log(n)
            cdf += np.exp(logp)
        return n


class UniformizationModels:
    """nModels uniformised migration models resident on the device (one per chain)."""

    def __init__(self, m: np.ndarray, sym: bool = False, ctx=None):
        self.ctx = ctx or _lib.default_context()
        m = _lib.as_c(m, np.float64)
        if m.ndim == 2:
            m = m[None]
        self.nModels, self.d = m.shape[0], m.shape[1]
        self.sym = bool(sym)
        self.ctx.check(_lib.lib().mtt_unif_set_models(self.ctx.handle, self.nModels, self.d,
                                                      _lib.ptr(m), int(self.sym)))

    def matrices(self, model: int = 0):
        """(Q, R, mu) as getQ/getR/getMu(sym) return them."""
        Q = np.empty((self.d, self.d))
        R = np.empty((self.d, self.d))
        mu = C.c_double()
        self.ctx.check(_lib.lib().mtt_unif_get_matrices(self.ctx.handle, model, _lib.ptr(Q),
                                                        _lib.ptr(R), C.byref(mu)))
        return Q, R, mu.value

    def rpow(self, model: int, kmax: int):
        """(R^0..R^kmax [kmax+1,d,d], RpowSteadyN) -- SCMigrationModel.getRpowN / RpowSteadyN."""
        out = np.empty((kmax + 1, self.d, self.d))
        steady = C.c_int()
        self.ctx.check(_lib.lib().mtt_unif_rpow(self.ctx.handle, model, kmax, _lib.ptr(out),
                                                C.byref(steady)))
        return out, steady.value

    def expm(self, L: Sequence[float], model: Optional[Sequence[int]] = None) -> np.ndarray:
        """expm(Q * L) for every L (the jblas call of UniformizationRetypeOperator.java:132)."""
        L = _lib.as_c(L, np.float64).ravel()
        mdl = np.zeros(len(L), np.int32) if model is None else _lib.as_c(model, np.int32)
        out = np.empty((len(L), self.d, self.d))
        self.ctx.check(_lib.lib().mtt_unif_expm_batch(self.ctx.handle, len(L), _lib.ptr(mdl),
                                                      _lib.ptr(L), _lib.ptr(out)))
        return out

    def branch_logp(self, tBottom, tTop, typeBottom, typeTop, chgOff, chgTime, chgType,
                    model=None):
        """getBranchTypeProb for a batch of branches -> (logProb, status)."""
        t0 = _lib.as_c(tBottom, np.float64)
        n = len(t0)
        t1 = _lib.as_c(tTop, np.float64)
        a = _lib.as_c(typeBottom, np.uint8)
        b = _lib.as_c(typeTop, np.uint8)
        off = _lib.as_c(chgOff, np.int64)
        ct = _lib.as_c(chgTime, np.float64)
        cty = _lib.as_c(chgType, np.uint8)
        mdl = np.zeros(n, np.int32) if model is None else _lib.as_c(model, np.int32)
        lp = np.empty(n)
        st = np.empty(n, np.int32)
        self.ctx.check(_lib.lib().mtt_unif_branch_logp_batch(
            self.ctx.handle, n, _lib.ptr(mdl), _lib.ptr(t0), _lib.ptr(t1), _lib.ptr(a),
            _lib.ptr(b), _lib.ptr(off), _lib.ptr(ct), _lib.ptr(cty), _lib.ptr(lp), _lib.ptr(st)))
        return lp, st

    def retype(self, brOff, tBottom, tTop, typeBottom, typeTop, uniforms, uOff, outCap=64,
               model=None, nVirtForced=None):
        """retypeBranch for a batch of jobs (mtt_unif_retype_batch).  Returns a dict of
        per-branch arrays nChg, chgTime [nBr,outCap], chgType, logProb, status, nVirt and
        per-job uUsed."""
        brOff = _lib.as_c(brOff, np.int64)
        nJobs = len(brOff) - 1
        t0 = _lib.as_c(tBottom, np.float64)
        nBr = len(t0)
        t1 = _lib.as_c(tTop, np.float64)
        a = _lib.as_c(typeBottom, np.uint8)
        b = _lib.as_c(typeTop, np.uint8)
        u = _lib.as_c(uniforms, np.float64)
        uOff = _lib.as_c(uOff, np.int64)
        mdl = np.zeros(nBr, np.int32) if model is None else _lib.as_c(model, np.int32)
        forced = None if nVirtForced is None else _lib.as_c(nVirtForced, np.int32)
        nChg = np.empty(nBr, np.int32)
        oct_ = np.zeros((nBr, outCap))
        octy = np.zeros((nBr, outCap), np.uint8)
        lp = np.empty(nBr)
        st = np.empty(nBr, np.int32)
        nv = np.empty(nBr, np.int32)
        used = np.empty(nJobs, np.int64)
        self.ctx.check(_lib.lib().mtt_unif_retype_batch(
            self.ctx.handle, nJobs, _lib.ptr(brOff), _lib.ptr(mdl), _lib.ptr(t0), _lib.ptr(t1),
            _lib.ptr(a), _lib.ptr(b), _lib.ptr(forced), _lib.ptr(u), _lib.ptr(uOff), outCap,
            _lib.ptr(nChg), _lib.ptr(oct_), _lib.ptr(octy), _lib.ptr(lp), _lib.ptr(st),
            _lib.ptr(nv), _lib.ptr(used)))
        return dict(nChg=nChg, chgTime=oct_, chgType=octy, logProb=lp, status=st, nVirt=nv,
                    uUsed=used)


class NodeRetype:
    """operators/NodeRetype.java: retype a random internal node and its <= 3 attached
    branches with the uniformization sampler.  Same Inputs: ``multiTypeTree``,
    ``migrationModel``, ``useSymmetrizedRates``."""

    def __init__(self, multiTypeTree: MultiTypeTree, migrationModel,
                 useSymmetrizedRates: bool = False, randomizer: Randomizer = None, ctx=None,
                 uniform_window: int = 256):
        self.mtTree = multiTypeTree
        self.migModel = migrationModel
        self.sym = bool(useSymmetrizedRates)
        self.rng = randomizer or Randomizer(127)
        self.ctx = ctx
        self.window = uniform_window

    def _models(self) -> UniformizationModels:
        return self.migModel._ensure_device(self.sym)

    def _branches(self, node: int) -> List[int]:
        """node (if not root), left, right -- the order of NodeRetype.java:40-44, 52-56."""
        t = self.mtTree
        kids = t.getChildren(node)
        return ([node] if t.parent[node] >= 0 else []) + kids

    def proposal(self) -> float:
        """Returns the log Hastings ratio, or -inf (NodeRetype.java:31-63)."""
        t = self.mtTree
        um = self._models()
        node = t.getLeafNodeCount() + self.rng.nextInt(t.getInternalNodeCount())   # :36-37
        branches = self._branches(node)

        def desc(bs):
            t0 = [t.height[b] for b in bs]
            t1 = [t.height[t.parent[b]] for b in bs]
            a = [t.getNodeType(b) for b in bs]
            bb = [t.getNodeType(int(t.parent[b])) for b in bs]
            return t0, t1, a, bb

        # probability of the current types along the attached branches (:40-44)
        t0, t1, a, bb = desc(branches)
        off = np.cumsum([0] + [t.getChangeCount(b) for b in branches])
        ct = [x for b in branches for x in t.changeTimes[b]]
        cty = [x for b in branches for x in t.changeTypes[b]]
        lp_old, st_old = um.branch_logp(t0, t1, a, bb, off, ct, cty)
        logHR = float(np.sum(lp_old))
        # new node type (:47-48)
        t.setNodeType(node, self.rng.nextInt(self.migModel.getNTypes()))
        # retype the attached branches from one uniform stream (:52-56)
        t0, t1, a, bb = desc(branches)
        nb = len(branches)
        new_lp = [0.0] * nb
        forced = [-1] * nb
        start, window = 0, self.window
        while start < nb:
            u = self.rng.peekDoubles(window)
            sl = slice(start, nb)
            r = um.retype([0, nb - start], t0[sl], t1[sl], a[sl], bb[sl], u, [0, window],
                          outCap=window, nVirtForced=forced[sl])
            st = r["status"]
            if np.any(st == ST_NEED_MORE_UNIFORMS):
                window *= 4
                continue
            bad = np.nonzero(st != ST_OK)[0]
            k = int(bad[0]) if len(bad) else nb - start
            for i in range(k):                      # branches retyped successfully
                b = branches[start + i]
                t.clearChanges(b)
                for j in range(int(r["nChg"][i])):
                    t.addChange(b, int(r["chgType"][i, j]), float(r["chgTime"][i, j]))
                new_lp[start + i] = float(r["logProb"][i])
            self.rng.skipDoubles(int(r["uUsed"][0]))
            if k == nb - start:
                break
            if st[k] == ST_NO_VALID_PATH:
                return float("-inf")                # NoValidPathException (:57-58)
            if st[k] == ST_NEEDS_POISSON_TAIL:
                # steady-state tail (UniformizationRetypeOperator.java:96-102): the host
                # Randomizer draws nVirt >= thresh, the branch is resubmitted with it forced
                thresh = int(r["nVirt"][k])
                muL = um.matrices(0)[2] * (t1[start + k] - t0[start + k])
                while True:
                    nv = self.rng.nextPoisson(muL)
                    if nv >= thresh:
                        break
                forced[start + k] = nv
                start += k
                continue
            # Pba rubbish / forward-backward failure: retypeBranch returned -Infinity and the
            # proposal carries on with the remaining branches (:141-144, :179-190)
            new_lp[start + k] = float("-inf")
            start += k + 1
        first = new_lp[0] if t.parent[node] >= 0 else 0.0
        rest = new_lp[1:] if t.parent[node] >= 0 else new_lp
        logHR -= first                               # logHR -= retypeBranch(node)
        logHR -= sum(rest)                           # logHR -= retype(left) + retype(right)
        return logHR
