"""ctypes harness of the CPU oracle (oracle/libmtt_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py -- never by the product package multitypetree_b200.
PARITY: piece 1 is pinned by the reference's KAT (SCLikelihoodTest.java:43-73); pieces 2 and 3
are "parity unpinned" (no reference test pins them; see mtt_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_DIR, "libmtt_oracle.so")
_lib = None
VP = C.c_void_p


def build(force: bool = False) -> str:
    srcs = [os.path.join(_DIR, f) for f in ("orc_density.c", "orc_unif.c", "orc_peel.c",
                                            "mtt_oracle.h", "Makefile")]
    stale = force or not os.path.exists(_LIB) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs)
    if stale:
        r = subprocess.run(["make", "-C", _DIR], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.orc_sc_logp_batch.restype = None
        L.orc_sc_logp_batch.argtypes = [C.c_int, C.c_int, C.c_int] + [VP] * 10 + [
            C.c_int, C.c_int, C.c_int, VP]
        L.orc_update_matrices.restype = None
        L.orc_update_matrices.argtypes = [C.c_int] + [VP] * 7
        L.orc_rpow_new.restype = VP
        L.orc_rpow_new.argtypes = [C.c_int, VP]
        L.orc_rpow_free.restype = None
        L.orc_rpow_free.argtypes = [VP]
        L.orc_rpow_get.restype = C.POINTER(C.c_double)
        L.orc_rpow_get.argtypes = [VP, C.c_int]
        L.orc_rpow_steady_n.restype = C.c_int
        L.orc_rpow_steady_n.argtypes = [VP]
        L.orc_rpow_size.restype = C.c_int
        L.orc_rpow_size.argtypes = [VP]
        L.orc_expm.restype = None
        L.orc_expm.argtypes = [C.c_int, VP, VP]
        L.orc_retype_branch.restype = C.c_int
        L.orc_retype_branch.argtypes = [VP, C.c_double, C.c_double, C.c_int, C.c_int, VP,
                                        C.c_int64, C.c_int, C.c_int, VP, VP, VP, VP, VP, VP]
        L.orc_branch_type_prob.restype = C.c_int
        L.orc_branch_type_prob.argtypes = [VP, C.c_double, C.c_double, C.c_int, C.c_int, VP, VP,
                                           C.c_int, VP]
        L.orc_hky_matrix.restype = None
        L.orc_hky_matrix.argtypes = [C.c_double, VP, C.c_double, VP]
        L.orc_tree_likelihood.restype = C.c_double
        L.orc_tree_likelihood.argtypes = [C.c_int, C.c_int, C.c_int, VP, VP, VP, VP, VP, VP, VP,
                                          C.c_double, C.c_double, C.c_double, C.c_int, VP]
        L.orc_tree_likelihood_mats.restype = C.c_double
        L.orc_tree_likelihood_mats.argtypes = [C.c_int, C.c_int, C.c_int, VP, VP, VP, VP, VP, VP,
                                               C.c_double, C.c_int, VP]
        L.orc_tree_likelihood_batch.restype = None
        L.orc_tree_likelihood_batch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, VP, VP, VP,
                                                VP, VP, VP, VP, C.c_double, C.c_double,
                                                C.c_double, C.c_int, C.c_int, VP]
        L.orc_backward_rate.restype = C.c_double
        L.orc_backward_rate.argtypes = [VP, C.c_int, C.c_int]
        L.orc_resolve_dense.restype = None
        L.orc_resolve_dense.argtypes = [VP, VP, VP]
        L.orc_untyped_particle.restype = C.c_int
        L.orc_untyped_particle.argtypes = [VP, VP, C.c_int, VP, VP, VP, VP, VP, C.c_int64, VP,
                                           VP, VP, VP]
        L.orc_untyped_particle_tree.restype = C.c_int
        L.orc_untyped_particle_tree.argtypes = [VP, VP, C.c_int, VP, VP, VP, VP, VP, C.c_int64,
                                                VP, VP, VP, C.c_int, VP, VP, VP, VP]
        L.orc_untyped_combine.restype = C.c_double
        L.orc_untyped_combine.argtypes = [VP, C.c_int]
        L.orc_poisson_window.restype = C.c_int
        L.orc_poisson_window.argtypes = [C.c_double, VP, C.c_int64, VP, VP]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(VP)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


# ------------------------------------------------------------------ piece 1

FAITHFUL, SORTED, SUFFSTAT = 0, 1, 2


def sc_logp_batch(batch, m, N, checkValidity=False, variant=FAITHFUL, nThreads=1):
    """batch: any object with the TreeBatch fields (nodeOff, height, parent, nodeType,
    chgCount, chgBase, chgTime, chgType)."""
    L = lib()
    m = _c(m, np.float64)
    N = _c(N, np.float64)
    d = m.shape[-1]
    nModels = 1 if m.ndim == 2 else m.shape[0]
    B = len(batch.nodeOff) - 1
    out = np.empty(B, np.float64)
    arrs = [_c(batch.nodeOff, np.int64), _c(batch.height, np.float64),
            _c(batch.parent, np.int32), _c(batch.nodeType, np.uint8),
            _c(batch.chgCount, np.uint16), _c(batch.chgBase, np.int64),
            _c(batch.chgTime, np.float64), _c(batch.chgType, np.uint8)]
    L.orc_sc_logp_batch(B, d, nModels, _p(m), _p(N), *[_p(a) for a in arrs],
                        int(bool(checkValidity)), variant, nThreads, _p(out))
    return out


class _OrcModel(C.Structure):
    _fields_ = [("nTypes", C.c_int), ("rateMatrix", VP), ("rateDim", C.c_int),
                ("popSizes", VP), ("rateScale", C.c_double), ("popScale", C.c_double),
                ("rateFlags", VP), ("useForward", C.c_int)]


def resolve_dense(nTypes, rateMatrix, popSizes, rateScale=1.0, popScale=1.0, rateFlags=None,
                  useForward=False):
    """SCMigrationModel.getBackwardRate / getPopSize for every (i,j) from the raw Inputs."""
    L = lib()
    rm = _c(rateMatrix, np.float64)
    ps = _c(popSizes, np.float64)
    fl = None if rateFlags is None else _c(rateFlags, np.uint8)
    mod = _OrcModel(nTypes, _p(rm), len(rm), _p(ps), rateScale, popScale, _p(fl),
                    int(bool(useForward)))
    m = np.zeros((nTypes, nTypes))
    N = np.zeros(nTypes)
    L.orc_resolve_dense(C.byref(mod), _p(m), _p(N))
    return m, N


# ------------------------------------------------------------------ piece 3

class _OrcUnifModel(C.Structure):
    _fields_ = [("d", C.c_int), ("Q", VP), ("R", VP), ("mu", C.c_double), ("rpow", VP)]


class UnifModel:
    """Q, R, mu and the lazy R^k cache of one migration model (sym or not)."""

    def __init__(self, m, sym=False):
        L = lib()
        m = _c(m, np.float64)
        d = m.shape[0]
        self.d = d
        Q, Qs, R, Rs = (np.zeros((d, d)) for _ in range(4))
        mu, mus = C.c_double(), C.c_double()
        L.orc_update_matrices(d, _p(m), _p(Q), _p(Qs), C.byref(mu), C.byref(mus), _p(R), _p(Rs))
        self.Q, self.R, self.mu = (Qs, Rs, mus.value) if sym else (Q, R, mu.value)
        self._rp = L.orc_rpow_new(d, _p(self.R))
        self._s = _OrcUnifModel(d, _p(self.Q), _p(self.R), self.mu, self._rp)

    def __del__(self):
        try:
            lib().orc_rpow_free(self._rp)
        except Exception:
            pass

    def rpow(self, n):
        p = lib().orc_rpow_get(self._rp, n)
        return np.ctypeslib.as_array(p, shape=(self.d, self.d)).copy()

    def steady_n(self):
        return lib().orc_rpow_steady_n(self._rp)

    def rpow_size(self):
        return lib().orc_rpow_size(self._rp)

    def retype_branch(self, tBottom, tTop, typeBottom, typeTop, u, nVirtForced=-1, cap=4096):
        L = lib()
        u = _c(u, np.float64)
        nChg, nVirt = C.c_int(), C.c_int()
        uUsed = C.c_int64()
        lp = C.c_double()
        ct = np.zeros(cap, np.float64)
        cty = np.zeros(cap, np.int32)
        st = L.orc_retype_branch(C.byref(self._s), tBottom, tTop, typeBottom, typeTop, _p(u),
                                 len(u), nVirtForced, cap, C.byref(nChg), _p(ct), _p(cty),
                                 C.byref(lp), C.byref(uUsed), C.byref(nVirt))
        n = min(nChg.value, cap)
        return dict(status=st, nChg=nChg.value, chgTime=ct[:n].copy(), chgType=cty[:n].copy(),
                    logProb=lp.value, uUsed=uUsed.value, nVirt=nVirt.value)

    def branch_type_prob(self, tBottom, tTop, typeBottom, chgTime, chgType, typeTop):
        L = lib()
        ct = _c(chgTime, np.float64)
        cty = _c(chgType, np.int32)
        lp = C.c_double()
        st = L.orc_branch_type_prob(C.byref(self._s), tBottom, tTop, typeBottom, len(ct), _p(ct),
                                    _p(cty), typeTop, C.byref(lp))
        return st, lp.value


def untyped_particle(um: "UnifModel", N, parent, height, leafType, leafOrder, u):
    """One particle of StructuredCoalescentUntypedTreeDensity.calculateLogP (:121-189):
    returns dict(status, logW, uUsed, nodeType, nEvents)."""
    L = lib()
    N = _c(N, np.float64)
    parent = _c(parent, np.int32)
    height = _c(height, np.float64)
    leafType = _c(leafType, np.uint8)
    leafOrder = _c(leafOrder, np.int32)
    u = _c(u, np.float64)
    n = len(parent)
    logW = C.c_double()
    used = C.c_int64()
    nodeType = np.zeros(n, np.int32)
    nEv = C.c_int32()
    st = L.orc_untyped_particle(C.byref(um._s), _p(N), n, _p(parent), _p(height), _p(leafType),
                                _p(leafOrder), _p(u), len(u), C.byref(logW), C.byref(used),
                                _p(nodeType), C.byref(nEv))
    return dict(status=st, logW=logW.value, uUsed=used.value, nodeType=nodeType, nEvents=nEv.value)


def untyped_particle_tree(um: "UnifModel", N, parent, height, leafType, leafOrder, u, chgCap=4096):
    """The same particle returned as a typed tree (MultiTypeTreeFromUntypedNewick.java:104-134):
    dict(status, logW, uUsed, nodeType, nChg, chgNode, chgType, chgTime) -- changes in the order
    they are added."""
    L = lib()
    N = _c(N, np.float64)
    parent = _c(parent, np.int32)
    height = _c(height, np.float64)
    leafType = _c(leafType, np.uint8)
    leafOrder = _c(leafOrder, np.int32)
    u = _c(u, np.float64)
    n = len(parent)
    logW = C.c_double()
    used = C.c_int64()
    nChg = C.c_int32()
    nodeType = np.zeros(n, np.int32)
    chgNode = np.zeros(chgCap, np.int32)
    chgType = np.zeros(chgCap, np.int32)
    chgTime = np.zeros(chgCap, np.float64)
    st = L.orc_untyped_particle_tree(C.byref(um._s), _p(N), n, _p(parent), _p(height),
                                     _p(leafType), _p(leafOrder), _p(u), len(u), C.byref(logW),
                                     C.byref(used), _p(nodeType), chgCap, C.byref(nChg),
                                     _p(chgNode), _p(chgType), _p(chgTime))
    k = min(nChg.value, chgCap)
    return dict(status=st, logW=logW.value, uUsed=used.value, nodeType=nodeType, nChg=nChg.value,
                chgNode=chgNode[:k], chgType=chgType[:k], chgTime=chgTime[:k])


def untyped_combine(logW):
    """log(mean(exp(logW))) as the reference computes it (:193-201)."""
    logW = _c(logW, np.float64)
    return lib().orc_untyped_combine(_p(logW), len(logW))


def expm(A):
    A = _c(A, np.float64)
    out = np.zeros_like(A)
    lib().orc_expm(A.shape[0], _p(A), _p(out))
    return out


# ------------------------------------------------------------------ piece 2

def hky_matrix(kappa, freqs, distance):
    f = _c(freqs, np.float64)
    P = np.zeros((4, 4))
    lib().orc_hky_matrix(kappa, _p(f), distance, _p(P))
    return P


def tree_likelihood(tipStates, weights, parent, height, catRates, catProps, freqs, kappa,
                    branchRate=1.0, pInv=0.0, useScaling=2, want_patterns=False):
    L = lib()
    ts = _c(tipStates, np.int32)
    nLeaves, P = ts.shape
    w = _c(weights, np.int32)
    par = _c(parent, np.int32)
    h = _c(height, np.float64)
    cr = _c(catRates, np.float64)
    cp = _c(catProps, np.float64)
    f = _c(freqs, np.float64)
    pl = np.zeros(P) if want_patterns else None
    lp = L.orc_tree_likelihood(nLeaves, P, len(cr), _p(ts), _p(w), _p(par), _p(h), _p(cr),
                               _p(cp), _p(f), kappa, branchRate, pInv, useScaling, _p(pl))
    return (lp, pl) if want_patterns else lp


def tree_likelihood_mats(tipStates, weights, parent, mats, catProps, freqs, pInv=0.0,
                         useScaling=2):
    L = lib()
    ts = _c(tipStates, np.int32)
    nLeaves, P = ts.shape
    w = _c(weights, np.int32)
    par = _c(parent, np.int32)
    mt = _c(mats, np.float64)
    cp = _c(catProps, np.float64)
    f = _c(freqs, np.float64)
    return L.orc_tree_likelihood_mats(nLeaves, P, len(cp), _p(ts), _p(w), _p(par), _p(mt),
                                      _p(cp), _p(f), pInv, useScaling, None)


def tree_likelihood_batch(tipStates, weights, parent, height, catRates, catProps, freqs, kappa,
                          branchRate=1.0, pInv=0.0, useScaling=2, nThreads=1):
    L = lib()
    ts = _c(tipStates, np.int32)
    nLeaves, P = ts.shape
    par = _c(parent, np.int32)
    h = _c(height, np.float64)
    B = par.shape[0]
    w = _c(weights, np.int32)
    cr = _c(catRates, np.float64)
    cp = _c(catProps, np.float64)
    f = _c(freqs, np.float64)
    out = np.zeros(B)
    L.orc_tree_likelihood_batch(B, nLeaves, P, len(cr), _p(ts), _p(w), _p(par), _p(h), _p(cr),
                                _p(cp), _p(f), kappa, branchRate, pInv, useScaling, nThreads,
                                _p(out))
    return out
