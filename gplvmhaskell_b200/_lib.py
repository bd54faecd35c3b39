"""ctypes binding of libgplvm_b200.so -- the C ABI declared in include/gplvm_b200.h.

There is no CPU fallback: importing works anywhere (so that `-m "not gpu"` tests can check the exported
symbols), but every compute entry point returns GPLVM_ERR_CUDA without a usable B200 and the wrappers
raise ``GplvmError``.  If the shared object is missing it is built with nvcc (gplvmhaskell_b200.build);
if that is impossible the import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_c_double_p = C.POINTER(C.c_double)
_c_i64_p = C.POINTER(C.c_int64)
_c_int_p = C.POINTER(C.c_int)

OK, ERR_ARG, ERR_CUDA, ERR_NCCL, ERR_NUMERIC, ERR_UNSUPPORTED, ERR_STATE = range(7)
STOP_ITERATIONS, STOP_TOLERANCE = 0, 1
MAX_M = 32


class GplvmError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__("libgplvm_b200 error %d: %s" % (code, message))
        self.code = code
        self.message = message


# every symbol of include/gplvm_b200.h: name -> (restype, argtypes)
SIGNATURES = {
    "gplvm_last_error": (C.c_char_p, []),
    "gplvm_version": (C.c_char_p, []),
    "gplvm_set_devices": (C.c_int, [C.c_int]),
    "gplvm_get_devices": (C.c_int, []),
    "gplvm_ppca_dense": (C.c_int, [_c_double_p, C.c_int64, C.c_int64, C.c_int64, _c_double_p, C.c_double, C.c_int,
                                   C.c_int64, C.c_double, _c_double_p, _c_double_p, _c_double_p, _c_i64_p]),
    "gplvm_ppca_missing": (C.c_int, [_c_double_p, C.c_int64, C.c_int64, C.c_int64, _c_double_p, C.c_double, C.c_int,
                                     C.c_int64, C.c_double, _c_double_p, _c_double_p, _c_double_p, _c_double_p, _c_i64_p]),
    "gplvm_ppca": (C.c_int, [_c_double_p, C.c_int64, C.c_int64, C.c_int64, _c_double_p, C.c_double, C.c_int,
                             C.c_int64, C.c_double, _c_double_p, _c_double_p, _c_double_p, _c_double_p, _c_int_p, _c_i64_p]),
    "gplvm_rbf_kernel": (C.c_int, [_c_double_p, C.c_int64, _c_double_p, C.c_int64, C.c_int64, _c_double_p, C.c_double,
                                   _c_double_p]),
    "gplvm_gp_chol_lml": (C.c_int, [_c_double_p, C.c_int64, C.c_int64, _c_double_p, _c_double_p, C.c_double, C.c_double,
                                    _c_double_p, _c_double_p, _c_double_p]),
    "gplvm_cholesky": (C.c_int, [_c_double_p, C.c_int64, _c_double_p]),
    "gplvm_tri_solve_lower": (C.c_int, [_c_double_p, C.c_int64, _c_double_p, C.c_int64, _c_double_p]),
    "gplvm_gp_posterior": (C.c_int, [_c_double_p, C.c_int64, _c_double_p, _c_double_p, C.c_int64, C.c_int64, _c_double_p, C.c_double,
                                     C.c_double, C.c_double, _c_double_p, C.c_int64, _c_double_p, _c_double_p, _c_double_p]),
    "gplvm_gp_posterior_upper": (C.c_int, [_c_double_p, C.c_int64, _c_double_p, _c_double_p, C.c_int64, C.c_int64, _c_double_p, C.c_double,
                                           C.c_double, C.c_double, _c_double_p, C.c_int64, _c_double_p, _c_double_p, _c_double_p]),
    "gplvm_pca": (C.c_int, [_c_double_p, C.c_int64, C.c_int64, C.c_int64, _c_double_p, _c_double_p, _c_double_p, _c_double_p,
                            _c_double_p, _c_double_p, _c_int_p]),
    "gplvm_dist_unique_id": (C.c_int, [C.c_void_p]),
    "gplvm_dist_init": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "gplvm_dist_finalize": (C.c_int, []),
    "gplvm_dist_rank": (C.c_int, []),
    "gplvm_dist_world": (C.c_int, []),
    "gplvm_session_create": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int64, C.c_int64,
                                       C.POINTER(C.c_void_p)]),
    "gplvm_session_destroy": (C.c_int, [C.c_void_p]),
    "gplvm_session_load_host": (C.c_int, [C.c_void_p, _c_double_p, C.c_int64]),
    "gplvm_session_synth": (C.c_int, [C.c_void_p, C.c_uint64, C.c_double]),
    "gplvm_session_read_data": (C.c_int, [C.c_void_p, _c_double_p, C.c_int64]),
    "gplvm_session_init": (C.c_int, [C.c_void_p, _c_double_p, C.c_double]),
    "gplvm_session_steps": (C.c_int, [C.c_void_p, C.c_int64]),
    "gplvm_session_run": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_double, _c_i64_p]),
    "gplvm_session_sync": (C.c_int, [C.c_void_p]),
    "gplvm_session_params": (C.c_int, [C.c_void_p, _c_double_p, _c_double_p, _c_double_p, _c_double_p, _c_double_p,
                                       _c_i64_p]),
    "gplvm_session_loglik": (C.c_int, [C.c_void_p, _c_double_p]),
    "gplvm_session_restored": (C.c_int, [C.c_void_p, _c_double_p, C.c_int64]),
    "gplvm_session_collective": (C.c_int, [C.c_void_p]),
    "gplvm_session_set_kernel_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "gplvm_session_last_steps_ms": (C.c_int, [C.c_void_p, _c_double_p, _c_double_p]),
    "gplvm_session_last_segments_ms": (C.c_int, [C.c_void_p, _c_double_p, C.c_int, _c_int_p]),
    "gplvm_session_masked_solver_path": (C.c_int, [C.c_void_p, _c_int_p]),
    "gplvm_session_segment_name": (C.c_char_p, [C.c_void_p, C.c_int]),
    "gplvm_kernel_launches": (C.c_int64, []),
    "gplvm_probe_dmma_tflops": (C.c_int, [_c_double_p]),
}

_lib = None


def library_path() -> str:
    return _build.LIB


def load() -> C.CDLL:
    """Load (building first if necessary) the shared library and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build_library()
    if not os.path.exists(path):
        raise ImportError("libgplvm_b200.so is missing and could not be built; there is no CPU fallback")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != OK:
        msg = load().gplvm_last_error()
        raise GplvmError(rc, msg.decode("utf-8", "replace") if msg else "")


def dptr(a):
    """double* of a C-contiguous float64 numpy array (None -> NULL)."""
    if a is None:
        return None
    return a.ctypes.data_as(_c_double_p)
