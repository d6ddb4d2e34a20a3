"""ctypes binding of libmtt_b200.so (the C-ABI declared in include/mtt_b200.h).

There is no CPU fallback: if the shared library is missing this module raises, and a Context
cannot be created without a B200.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MTT_B200_LIB") or os.path.join(_PKG, "libmtt_b200.so")

_lib = None

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)
c_u8p = C.POINTER(C.c_uint8)
c_u16p = C.POINTER(C.c_uint16)
VP = C.c_void_p

# name -> (restype, argtypes); every symbol include/mtt_b200.h declares
ABI = {
    "mtt_abi_version": (C.c_int, []),
    "mtt_ctx_create": (C.c_int, [C.c_int, C.POINTER(VP)]),
    "mtt_ctx_destroy": (C.c_int, [VP]),
    "mtt_last_error": (C.c_char_p, [VP]),
    "mtt_ctx_set_stream": (C.c_int, [VP, VP]),
    "mtt_ctx_synchronize": (C.c_int, [VP]),
    "mtt_ctx_launch_count": (C.c_int64, [VP]),
    "mtt_host_register": (C.c_int, [VP, C.c_int64]),
    "mtt_host_unregister": (C.c_int, [VP]),
    "mtt_sc_logp_batch": (C.c_int, [VP, C.c_int64, C.c_int, C.c_int64, VP, VP, VP, VP, VP, VP,
                                    VP, VP, VP, VP, C.c_int, VP, VP]),
    "mtt_sc_logp_batch_dev": (C.c_int, [VP, C.c_int64, C.c_int, C.c_int64, VP, VP, VP, VP, VP,
                                        VP, VP, VP, VP, VP, C.c_int, VP, VP]),
    "mtt_sc_logp_owner_batch": (C.c_int, [VP, C.c_int64, C.c_int, C.c_int64, VP, VP, VP, VP, VP,
                                          VP, VP, VP, VP, VP, C.c_int, VP, VP]),
    "mtt_sc_logp_owner_batch_dev": (C.c_int, [VP, C.c_int64, C.c_int, C.c_int64, VP, VP, VP, VP,
                                              VP, VP, VP, VP, VP, VP, C.c_int, VP, VP]),
    "mtt_sc_set_max_tree_events": (C.c_int, [VP, C.c_int64]),
    "mtt_sc_set_tips": (C.c_int, [VP, C.c_int, VP, VP]),
    "mtt_sc_logp_tips_batch": (C.c_int, [VP, C.c_int64, C.c_int, C.c_int64, VP, VP, VP, VP, VP,
                                         VP, VP, VP, VP, C.c_int, VP, VP]),
    "mtt_unif_set_models": (C.c_int, [VP, C.c_int, C.c_int, VP, C.c_int]),
    "mtt_unif_get_matrices": (C.c_int, [VP, C.c_int, VP, VP, VP]),
    "mtt_unif_rpow": (C.c_int, [VP, C.c_int, C.c_int, VP, VP]),
    "mtt_unif_expm_batch": (C.c_int, [VP, C.c_int64, VP, VP, VP]),
    "mtt_unif_branch_logp_batch": (C.c_int, [VP, C.c_int64, VP, VP, VP, VP, VP, VP, VP, VP, VP,
                                             VP]),
    "mtt_untyped_logp_batch": (C.c_int, [VP, C.c_int64, C.c_int, VP, VP, VP, VP, VP, VP, VP, VP,
                                        VP, VP, VP, VP, VP]),
    "mtt_untyped_sample_batch": (C.c_int, [VP, C.c_int64, C.c_int, VP, VP, VP, VP, VP, VP, VP, VP,
                                          VP, C.c_int, VP, VP, VP, VP, VP, VP, VP, VP]),
    "mtt_unif_retype_batch": (C.c_int, [VP, C.c_int64, VP, VP, VP, VP, VP, VP, VP, VP, VP,
                                        C.c_int, VP, VP, VP, VP, VP, VP, VP]),
    "mtt_unif_node_retype_batch": (C.c_int, [VP, C.c_int64, VP, VP, VP, VP, VP, VP, VP, VP, VP, VP,
                                             C.c_int64, VP, VP, VP, VP, VP, VP, VP, VP, VP,
                                             C.c_int, VP, VP, VP, VP, VP, VP, VP]),
    "mtt_peel_create": (C.c_int, [VP, C.c_int, C.c_int64, C.c_int, C.POINTER(VP)]),
    "mtt_peel_destroy": (C.c_int, [VP]),
    "mtt_peel_set_tip_states": (C.c_int, [VP, VP]),
    "mtt_peel_set_pattern_weights": (C.c_int, [VP, VP]),
    "mtt_peel_set_site_model": (C.c_int, [VP, VP, VP, VP, C.c_double, C.c_double]),
    "mtt_peel_set_hky": (C.c_int, [VP, C.c_double]),
    "mtt_peel_logl_batch": (C.c_int, [VP, C.c_int64, VP, VP, VP]),
    "mtt_peel_logl_batch_params": (C.c_int, [VP, C.c_int64, VP, VP, VP, VP, VP]),
    "mtt_peel_logl_batch_rates": (C.c_int, [VP, C.c_int64, VP, VP, VP, VP]),
    "mtt_peel_logl_batch_mats": (C.c_int, [VP, C.c_int64, VP, VP, VP]),
    "mtt_peel_logl_batch_dev": (C.c_int, [VP, C.c_int64, VP, VP, VP]),
    "mtt_peel_logl_parts_batch": (C.c_int, [VP, C.c_int64, VP, VP, VP]),
    "mtt_peel_logl_parts_batch_dev": (C.c_int, [VP, C.c_int64, VP, VP, VP]),
    "mtt_peel_pattern_logl": (C.c_int, [VP, VP, VP, VP]),
    "mtt_peel_chains_create": (C.c_int, [VP, C.c_int, C.POINTER(VP)]),
    "mtt_peel_chains_destroy": (C.c_int, [VP]),
    "mtt_peel_chains_update": (C.c_int, [VP, VP, VP, VP, VP, VP, VP, VP]),
    "mtt_peel_chains_update_sparse": (C.c_int, [VP, VP, VP, VP, VP, VP, VP, VP, VP]),
    "mtt_peel_chains_update_parts": (C.c_int, [VP, VP, VP, VP, VP, VP, VP, VP]),
    "mtt_peel_chains_accept": (C.c_int, [VP, VP]),
    "mtt_peel_chains_last_entries": (C.c_int64, [VP]),
    "mtt_peel_chains_set_timing": (C.c_int, [VP, C.c_int]),
    "mtt_peel_chains_set_verify": (C.c_int, [VP, C.c_int]),
    "mtt_peel_chains_last_device_ms": (C.c_double, [VP]),
    "mtt_sim_trees_create": (C.c_int, [VP, C.c_int64, C.c_int, C.c_int, VP, VP, VP, VP,
                                       C.c_uint64, C.POINTER(VP)]),
    "mtt_sim_trees_destroy": (C.c_int, [VP]),
    "mtt_sim_trees_total_changes": (C.c_int64, [VP]),
    "mtt_sim_trees_device": (C.c_int, [VP] + [C.POINTER(VP)] * 8),
    "mtt_sim_trees_export": (C.c_int, [VP] * 9),
}

# host-side helpers that live in the same shared object (not part of the drop-in boundary)
HOST = {
    "mtt_sim_run": (VP, [C.c_int64, C.c_int, C.c_int, VP, VP, VP, VP, C.c_uint64, C.c_int]),
    "mtt_sim_total_changes": (C.c_int64, [VP]),
    "mtt_sim_export": (None, [VP, VP, VP, VP, VP, VP, VP, VP, VP, VP]),
    "mtt_sim_free": (None, [VP]),
    "mtt_rng_new": (VP, [C.c_int64]),
    "mtt_rng_free": (None, [VP]),
    "mtt_rng_set_seed": (C.c_int, [VP, C.c_int64]),
    "mtt_rng_next_double": (C.c_double, [VP]),
    "mtt_rng_next_int": (C.c_int32, [VP, C.c_int32]),
    "mtt_rng_next_doubles": (None, [VP, C.c_int64, VP]),
    "mtt_rng_peek_doubles": (None, [VP, C.c_int64, VP]),
    "mtt_rng_skip_doubles": (None, [VP, C.c_int64]),
}


class MttError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mtt_b200 error %d: %s" % (code, msg))
        self.code = code


def lib():
    """Load libmtt_b200.so; fail loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: run `python -m multitypetree_b200.build` (or "
                "__graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for table in (ABI, HOST):
            for name, (res, args) in table.items():
                fn = getattr(L, name)   # AttributeError if a declared symbol is not exported
                fn.restype = res
                fn.argtypes = args
        _lib = L
    return _lib


def ptr(a):
    """Pointer to a numpy array's data (None -> NULL)."""
    if a is None:
        return None
    return a.ctypes.data_as(VP)


def as_c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Context:
    """One mtt_ctx: a device, a stream and the library-owned device buffers."""

    def __init__(self, device: int = -1):
        self._h = VP()
        rc = lib().mtt_ctx_create(device, C.byref(self._h))
        if rc != 0:
            raise MttError(rc, lib().mtt_last_error(None).decode())

    @property
    def handle(self):
        return self._h

    def check(self, rc):
        if rc != 0:
            raise MttError(rc, lib().mtt_last_error(self._h).decode())

    def set_stream(self, cuda_stream_ptr):
        self.check(lib().mtt_ctx_set_stream(self._h, cuda_stream_ptr))

    def synchronize(self):
        self.check(lib().mtt_ctx_synchronize(self._h))

    def launch_count(self) -> int:
        return int(lib().mtt_ctx_launch_count(self._h))

    def close(self):
        if self._h:
            lib().mtt_ctx_destroy(self._h)
            self._h = VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(-1)
    return _default_ctx
