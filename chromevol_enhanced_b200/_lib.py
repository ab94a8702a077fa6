"""ctypes binding of the C ABI declared in include/chromevol_b200.h.

There is no CPU fallback: if the shared library is missing or a call fails, an exception is
raised.  Pointers are raw device addresses taken from torch tensors (``Tensor.data_ptr()``).
"""
from __future__ import annotations

import ctypes
import os

from . import _build

_c = ctypes
_lib = None

c_dp = _c.c_void_p  # device pointer (any element type)


class CevbError(RuntimeError):
    pass


MAX_STATES = 288     # CEVB_MAX_N in include/chromevol_b200.h: largest matrix order of the kernels


_PROTOTYPES = {
    "cevb_abi_version": (_c.c_int, []),
    "cevb_device_info": (_c.c_int, [_c.POINTER(_c.c_int), _c.POINTER(_c.c_int), _c.POINTER(_c.c_int),
                                    _c.POINTER(_c.c_size_t)]),
    "cevb_last_error": (_c.c_char_p, []),
    "cevb_ldq": (_c.c_int, [_c.c_int]),
    "cevb_ldp": (_c.c_int, [_c.c_int]),
    "cevb_expm_scratch_bytes": (_c.c_size_t, [_c.c_int]),
    "cevb_expm_grid": (_c.c_int, [_c.c_int, _c.c_longlong]),
    "cevb_build_q": (_c.c_int, [c_dp, _c.c_int, _c.c_int, c_dp, _c.c_void_p]),
    "cevb_expm_prepare": (_c.c_int, [c_dp, _c.c_int, _c.c_int, c_dp, c_dp, c_dp, _c.c_void_p]),
    "cevb_expm_branches": (_c.c_int, [c_dp, c_dp, c_dp, c_dp, c_dp, _c.c_int, _c.c_int, _c.c_int,
                                      c_dp, c_dp, c_dp, c_dp, _c.c_int, _c.c_void_p]),
    "cevb_expm_branches_list": (_c.c_int, [c_dp, c_dp, c_dp, c_dp, c_dp, _c.c_int, _c.c_int,
                                           _c.c_int, c_dp, c_dp, c_dp, c_dp, _c.c_int, c_dp, c_dp,
                                           _c.c_longlong, _c.c_void_p]),
    "cevb_taylor_doubles": (_c.c_size_t, [_c.c_int]),
    "cevb_taylor_matrices_supported": (_c.c_int, [_c.c_int]),
    "cevb_taylor_scratch_doubles": (_c.c_size_t, [_c.c_int]),
    "cevb_taylor_prepare": (_c.c_int, [c_dp, _c.c_int, _c.c_int, _c.c_double, c_dp, c_dp, c_dp, c_dp,
                                       _c.c_void_p]),
    "cevb_expm_plan": (_c.c_int, [c_dp, c_dp, _c.c_int, _c.c_int, _c.c_longlong, c_dp, c_dp, c_dp,
                                  c_dp, c_dp, c_dp, c_dp, _c.c_void_p]),
    "cevb_taylor_matrices": (_c.c_int, [c_dp, c_dp, c_dp, _c.c_int, _c.c_int, _c.c_int, c_dp,
                                        c_dp, c_dp, _c.c_longlong, c_dp, c_dp, c_dp, c_dp, c_dp,
                                        c_dp, c_dp, _c.c_void_p]),
    "cevb_taylor_apply": (_c.c_int, [c_dp, c_dp, _c.c_int, _c.c_int, _c.c_int, c_dp, _c.c_int,
                                     _c.c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, _c.c_void_p]),
    "cevb_prune_fused": (_c.c_int, [c_dp, c_dp, c_dp, c_dp, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                    _c.c_int, c_dp, _c.POINTER(_c.c_int32), _c.c_int, c_dp, c_dp,
                                    c_dp, c_dp, c_dp, _c.POINTER(_c.c_int32),
                                    c_dp, _c.c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp,
                                    c_dp, c_dp, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p]),
    "cevb_prune": (_c.c_int, [c_dp, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, c_dp,
                              _c.POINTER(_c.c_int32), _c.c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp,
                              c_dp, c_dp, _c.c_void_p]),
    "cevb_argmax_states": (_c.c_int, [c_dp, _c.c_int, _c.c_int, _c.c_int, c_dp, c_dp, c_dp,
                                      _c.c_void_p]),
    "cevb_marginal_down": (_c.c_int, [c_dp, _c.c_int, _c.c_int, _c.c_int, _c.c_int, c_dp, c_dp, c_dp,
                                      c_dp, c_dp, _c.POINTER(_c.c_int32), _c.c_int, c_dp, c_dp, c_dp,
                                      c_dp, _c.c_void_p]),
    "cevb_joint_reconstruct": (_c.c_int, [c_dp, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                          c_dp, c_dp, _c.POINTER(_c.c_int32), _c.c_int, c_dp, c_dp,
                                          c_dp, c_dp, c_dp, _c.c_int, c_dp, c_dp, c_dp, c_dp, c_dp,
                                          c_dp, _c.c_void_p]),
    "cevb_stoch_map": (_c.c_int, [c_dp, _c.c_int, c_dp, _c.c_int, c_dp, c_dp, c_dp, _c.c_int,
                                  _c.c_ulonglong, _c.c_longlong, _c.c_longlong, _c.c_int, c_dp, c_dp,
                                  c_dp, _c.c_longlong, c_dp, _c.c_void_p]),
    "cevb_stoch_map_sets": (_c.c_int, [c_dp, _c.c_int, _c.c_longlong, c_dp, c_dp, _c.c_int, c_dp,
                                       _c.c_int, c_dp, c_dp, c_dp, _c.c_ulonglong, _c.c_longlong,
                                       _c.c_longlong, _c.c_int, c_dp, c_dp, c_dp, _c.c_longlong,
                                       c_dp, _c.c_void_p]),
    "cevb_stoch_branch_trace": (_c.c_int, [c_dp, _c.c_int, _c.c_int, _c.c_double, _c.c_int,
                                           _c.c_ulonglong, _c.c_ulonglong, _c.c_int, c_dp, c_dp, c_dp,
                                           c_dp, c_dp, _c.c_void_p]),
    "cevb_diag_fp64_peak": (_c.c_int, [_c.c_int, _c.c_int, _c.c_int, c_dp, _c.c_void_p]),
    "cevb_diag_set_phase_buffer": (_c.c_int, [c_dp]),
    "cevb_diag_dmma_probe": (_c.c_int, [_c.c_int, _c.c_int, _c.c_int, _c.c_int, c_dp, _c.c_void_p]),
}

# symbols include/chromevol_b200.h declares (the CPU test-suite checks every one is exported)
HEADER_SYMBOLS = [name for name in _PROTOTYPES if not name.startswith("cevb_diag_")]


def library_path():
    return _build.LIB_PATH


def load(build_if_missing=True):
    """Load (building first if the .so is absent) and type the shared library."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("CEVB_LIB")                        # CEVB_LIB: development builds only
    if not path:
        path = _build.LIB_PATH
        if _build.needs_build():
            # missing, or built from other sources than the tree holds (an edited kernel, a
            # pulled commit): a stale binary is never used silently
            if not build_if_missing and not os.path.exists(path):
                raise CevbError(f"CUDA extension not built: {path}")
            try:
                _build.build()
            except RuntimeError as e:
                raise CevbError(f"CUDA extension is missing or stale and cannot be rebuilt: {e}") from e
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in _PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.cevb_abi_version() != 2:
        raise CevbError("libcevb200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        msg = load().cevb_last_error()
        raise CevbError(f"{what} failed (status {status}): {msg.decode() if msg else ''}")


def require_device():
    """Fail loudly unless a compute-capability-10.x GPU is current (no CPU fallback exists)."""
    import torch
    if not torch.cuda.is_available():
        raise CevbError("chromevol_enhanced_b200 needs a CUDA device (sm_100a); none is visible "
                        "and there is no CPU fallback")
    lib = load()
    sm, maj, mnr, optin = _c.c_int(), _c.c_int(), _c.c_int(), _c.c_size_t()
    check(lib.cevb_device_info(_c.byref(sm), _c.byref(maj), _c.byref(mnr), _c.byref(optin)),
          "cevb_device_info")
    return {"sm_count": sm.value, "cc": (maj.value, mnr.value), "smem_optin": optin.value}


def ptr(tensor):
    return _c.c_void_p(tensor.data_ptr()) if tensor is not None else _c.c_void_p(0)
