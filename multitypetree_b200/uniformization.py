"""Host-side wrappers of the uniformization entry points (csrc/k3_unif.cu); the operators that
call them (NodeRetype, NodeShiftRetype, ...) live in operators.py.

The numerical work -- Q/R/mu (SCMigrationModel.java:150-201), the R^k chain (:404-473), expm,
the virtual-event count, forward-backward path sampling and path probabilities
(UniformizationRetypeOperator.java:74-275) -- runs on the device.  What stays on the host is
what stays in Java in the real plugin: choosing the node and its new type with the Randomizer
(NodeRetype.java:36-37, 47-48) and writing the returned changes back into the tree.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib

ST_OK, ST_NO_VALID_PATH, ST_PBA_RUBBISH, ST_FB_FELL_THROUGH = 0, 1, 2, 3
ST_NEEDS_POISSON_TAIL, ST_NEED_MORE_UNIFORMS, ST_NEED_MORE_POWERS, ST_SKIPPED = 4, 5, 6, 7


class Randomizer:
    """BEAST 2 ``Randomizer`` stream (MersenneTwisterFast) restated on the host
    (csrc/host/randomizer.cpp; [EXT], parity with a JVM unpinned)."""

    def __init__(self, seed: int = 127):
        self._h = _lib.lib().mtt_rng_new(seed)
        if not self._h:
            raise ValueError("Non zero random seed required (low 32 bits of %d are zero)" % seed)

    def __del__(self):
        try:
            if self._h:
                _lib.lib().mtt_rng_free(self._h)
        except Exception:
            pass

    def setSeed(self, seed: int):
        if _lib.lib().mtt_rng_set_seed(self._h, seed) != 0:
            raise ValueError("Non zero random seed required (low 32 bits of %d are zero)" % seed)

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

    def nextExponential(self, lam: float) -> float:
        """BEAST ``Randomizer.nextExponential``: -log(1 - nextDouble()) / lambda  ([EXT])."""
        return -1.0 * np.log(1 - self.nextDouble()) / lam

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
        self._m = m
        self._upload()

    def _upload(self):
        self.ctx.check(_lib.lib().mtt_unif_set_models(self.ctx.handle, self.nModels, self.d,
                                                      _lib.ptr(self._m), int(self.sym)))
        self.ctx._unif_owner = self

    def _activate(self):
        """A ctx holds ONE model set (mtt_unif_set_models); if another UniformizationModels has
        been installed on this ctx since, install this one again (its R^k cache restarts)."""
        if getattr(self.ctx, "_unif_owner", None) is not self:
            self._upload()

    def matrices(self, model: int = 0):
        """(Q, R, mu) as getQ/getR/getMu(sym) return them."""
        self._activate()
        Q = np.empty((self.d, self.d))
        R = np.empty((self.d, self.d))
        mu = C.c_double()
        self.ctx.check(_lib.lib().mtt_unif_get_matrices(self.ctx.handle, model, _lib.ptr(Q),
                                                        _lib.ptr(R), C.byref(mu)))
        return Q, R, mu.value

    def rpow(self, model: int, kmax: int):
        """(R^0..R^kmax [kmax+1,d,d], RpowSteadyN) -- SCMigrationModel.getRpowN / RpowSteadyN."""
        self._activate()
        out = np.empty((kmax + 1, self.d, self.d))
        steady = C.c_int()
        self.ctx.check(_lib.lib().mtt_unif_rpow(self.ctx.handle, model, kmax, _lib.ptr(out),
                                                C.byref(steady)))
        return out, steady.value

    def expm(self, L: Sequence[float], model: Optional[Sequence[int]] = None) -> np.ndarray:
        """expm(Q * L) for every L (the jblas call of UniformizationRetypeOperator.java:132)."""
        self._activate()
        L = _lib.as_c(L, np.float64).ravel()
        mdl = np.zeros(len(L), np.int32) if model is None else _lib.as_c(model, np.int32)
        out = np.empty((len(L), self.d, self.d))
        self.ctx.check(_lib.lib().mtt_unif_expm_batch(self.ctx.handle, len(L), _lib.ptr(mdl),
                                                      _lib.ptr(L), _lib.ptr(out)))
        return out

    def branch_logp(self, tBottom, tTop, typeBottom, typeTop, chgOff, chgTime, chgType,
                    model=None):
        """getBranchTypeProb for a batch of branches -> (logProb, status)."""
        self._activate()
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

    @staticmethod
    def retype_buffers(nJobs: int, nBr: int, outCap: int, pin: bool = False) -> dict:
        """Output buffers of ``retype`` for reuse across calls (``out=``); with ``pin`` they are
        page-locked (mtt_host_register) so the read-back runs at the link rate."""
        out = dict(nChg=np.empty(nBr, np.int32), chgTime=np.zeros((nBr, outCap)),
                   chgType=np.zeros((nBr, outCap), np.uint8), logProb=np.empty(nBr),
                   status=np.empty(nBr, np.int32), nVirt=np.empty(nBr, np.int32),
                   uUsed=np.empty(nJobs, np.int64))
        if pin:
            for v in out.values():
                _lib.lib().mtt_host_register(v.ctypes.data, v.nbytes)
        return out

    def retype(self, brOff, tBottom, tTop, typeBottom, typeTop, uniforms, uOff, outCap=64,
               model=None, nVirtForced=None, out=None):
        """retypeBranch for a batch of jobs (mtt_unif_retype_batch).  Returns a dict of
        per-branch arrays nChg, chgTime [nBr,outCap], chgType, logProb, status, nVirt and
        per-job uUsed (``out``: buffers from ``retype_buffers`` to write into; entries of
        chgTime / chgType beyond a branch's nChg are then left as they were)."""
        self._activate()
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
        if out is None:
            out = self.retype_buffers(nJobs, nBr, outCap)
        elif out["chgTime"].shape != (nBr, outCap) or len(out["uUsed"]) != nJobs:
            raise ValueError("out buffers do not match this call")
        nChg, oct_, octy, lp, st, nv, used = (out[k] for k in (
            "nChg", "chgTime", "chgType", "logProb", "status", "nVirt", "uUsed"))
        self.ctx.check(_lib.lib().mtt_unif_retype_batch(
            self.ctx.handle, nJobs, _lib.ptr(brOff), _lib.ptr(mdl), _lib.ptr(t0), _lib.ptr(t1),
            _lib.ptr(a), _lib.ptr(b), _lib.ptr(forced), _lib.ptr(u), _lib.ptr(uOff), outCap,
            _lib.ptr(nChg), _lib.ptr(oct_), _lib.ptr(octy), _lib.ptr(lp), _lib.ptr(st),
            _lib.ptr(nv), _lib.ptr(used)))
        return out
