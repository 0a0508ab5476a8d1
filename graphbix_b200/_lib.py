"""ctypes binding of libgraphbix_cuda.so (the C ABI declared in include/graphbix_cuda.h).

The library is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a).  There is no fallback:
if the shared object is missing, or no CUDA device is visible, the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgraphbix_cuda.so")

GBX_OK = 0
GBX_MAX_Q = 16


class GbxError(RuntimeError):
    pass


class SbmOptions(C.Structure):
    _fields_ = [("degreecorrection", C.c_int), ("priorlearning", C.c_int), ("learningrate", C.c_double),
                ("maxCrs", C.c_double), ("itrmax", C.c_int), ("bp_conv_threshold", C.c_double),
                ("damping", C.c_double), ("check_every", C.c_int), ("mstep_exact", C.c_int)]


class SbmResult(C.Structure):
    _fields_ = [("FE", C.c_double), ("CVBayes", C.c_double), ("CVGP", C.c_double), ("CVGT", C.c_double),
                ("CVMAP", C.c_double), ("varCVBayes", C.c_double), ("varCVGP", C.c_double),
                ("varCVGT", C.c_double), ("varCVMAP", C.c_double), ("fail", C.c_int), ("cnv", C.c_int),
                ("itrnum", C.c_int), ("itrs_done", C.c_int), ("nonfinite", C.c_int), ("BPconv", C.c_double),
                ("max_dpsi", C.c_double)]


class ModOptions(C.Structure):
    _fields_ = [("dc", C.c_int), ("learning", C.c_int), ("priorlearning", C.c_int), ("learningrate", C.c_double),
                ("maxomega", C.c_double), ("itrmax", C.c_int), ("betafrac", C.c_double),
                ("bp_conv_threshold", C.c_double), ("damping", C.c_double), ("check_every", C.c_int)]


class ModResult(C.Structure):
    _fields_ = [("excm", C.c_double), ("alpha", C.c_double), ("beta", C.c_double), ("omegain", C.c_double),
                ("omegaout", C.c_double), ("FE", C.c_double), ("CVBayes", C.c_double), ("CVGP", C.c_double),
                ("CVGT", C.c_double), ("CVMAP", C.c_double), ("varCVBayes", C.c_double), ("varCVGP", C.c_double),
                ("varCVGT", C.c_double), ("varCVMAP", C.c_double), ("cnv", C.c_int), ("itrnum", C.c_int),
                ("itrs_done", C.c_int), ("nonfinite", C.c_int), ("BPconv", C.c_double), ("max_dpsi", C.c_double)]


class LsbmOptions(C.Structure):
    _fields_ = [("dc", C.c_int), ("learning", C.c_int), ("priorlearning", C.c_int), ("learningrate", C.c_double),
                ("itrmax", C.c_int), ("bp_conv_threshold", C.c_double), ("REDfrac", C.c_double), ("damping", C.c_double)]


class LsbmResult(C.Structure):
    _fields_ = [("FE", C.c_double), ("CVBayes", C.c_double), ("CVGP", C.c_double), ("CVGT", C.c_double),
                ("CVMAP", C.c_double), ("varCVBayes", C.c_double), ("varCVGP", C.c_double), ("varCVGT", C.c_double),
                ("varCVMAP", C.c_double), ("cnv", C.c_int), ("itrnum", C.c_int), ("itrs_done", C.c_int),
                ("nonfinite", C.c_int), ("BPconv", C.c_double)]


# name -> (restype, argtypes); every symbol include/graphbix_cuda.h declares
_P = C.c_void_p
_D = C.POINTER(C.c_double)
SIGNATURES = {
    "gbx_version": (C.c_int, []),
    "gbx_last_error": (C.c_char_p, []),
    "gbx_device_count": (C.c_int, []),
    "gbx_sbm_default_options": (None, [C.POINTER(SbmOptions)]),
    "gbx_sbm_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int64, C.c_int64, C.POINTER(C.c_int64), C.c_int]),
    "gbx_sbm_destroy": (C.c_int, [_P]),
    "gbx_sbm_set_stream": (C.c_int, [_P, _P]),
    "gbx_sbm_set_options": (C.c_int, [_P, C.POINTER(SbmOptions)]),
    "gbx_sbm_init": (C.c_int, [_P, _D, _D, _D, C.c_uint64]),
    "gbx_sbm_em": (C.c_int, [_P, C.POINTER(SbmResult)]),
    "gbx_sbm_iterate": (C.c_int, [_P, C.c_int, _D]),
    "gbx_sbm_bp_sweeps": (C.c_int, [_P, C.c_int]),
    "gbx_sbm_finalize": (C.c_int, [_P, C.POINTER(SbmResult)]),
    "gbx_sbm_get_state": (C.c_int, [_P, _D, _D, _D, _D, _D]),
    "gbx_sbm_get_messages": (C.c_int, [_P, _D]),
    "gbx_sbm_get_assignment": (C.c_int, [_P, C.POINTER(C.c_int32)]),
    "gbx_sbm_launch_count": (C.c_int64, [_P]),
    "gbx_sbm_num_links": (C.c_int64, [_P]),
    "gbx_sbm_device_bytes": (C.c_int64, [_P]),
    "gbx_sbm_sweep_bytes": (C.c_double, [_P]),
    "gbx_sbm_sweep_kind": (C.c_int, [_P]),
    "gbx_mod_default_options": (None, [C.POINTER(ModOptions)]),
    "gbx_mod_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int64, C.c_int64, C.POINTER(C.c_int64), C.c_int]),
    "gbx_mod_destroy": (C.c_int, [_P]),
    "gbx_mod_set_stream": (C.c_int, [_P, _P]),
    "gbx_mod_set_options": (C.c_int, [_P, C.POINTER(ModOptions)]),
    "gbx_mod_init": (C.c_int, [_P, _D, _D, C.c_uint64]),
    "gbx_mod_em": (C.c_int, [_P, C.POINTER(ModResult)]),
    "gbx_mod_iterate": (C.c_int, [_P, C.c_int, _D]),
    "gbx_mod_finalize": (C.c_int, [_P, C.POINTER(ModResult)]),
    "gbx_mod_get_state": (C.c_int, [_P, _D, _D, _D, _D]),
    "gbx_mod_get_messages": (C.c_int, [_P, _D]),
    "gbx_mod_get_assignment": (C.c_int, [_P, C.POINTER(C.c_int32)]),
    "gbx_mod_launch_count": (C.c_int64, [_P]),
    "gbx_sbm_create_part": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int64, C.c_int64, C.POINTER(C.c_int64), C.c_int,
                                      C.c_int, C.c_int]),
    "gbx_mod_create_part": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int64, C.c_int64, C.POINTER(C.c_int64), C.c_int,
                                      C.c_int, C.c_int]),
    "gbx_sbm_part_info": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "gbx_sbm_ipc_export": (C.c_int, [_P, C.c_int, C.POINTER(C.c_ubyte)]),
    "gbx_sbm_ipc_import": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_ubyte)]),
    "gbx_sbm_part_set_buffers": (C.c_int, [_P, _P, _P]),
    "gbx_sbm_part_set_exchange": (C.c_int, [_P, C.c_int]),
    "gbx_lsbm_default_options": (None, [C.POINTER(LsbmOptions)]),
    "gbx_lsbm_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int64, C.c_int64, C.POINTER(C.c_int64), C.c_int, C.c_int]),
    "gbx_lsbm_create_batch": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                        C.POINTER(C.POINTER(C.c_int64)), C.c_int, C.c_int]),
    "gbx_lsbm_count": (C.c_int, [_P]),
    "gbx_lsbm_init_batch": (C.c_int, [_P, _D, _D, C.POINTER(_D)]),
    "gbx_lsbm_init_random": (C.c_int, [_P, _D, _D, C.POINTER(C.c_uint64)]),
    "gbx_lsbm_get_state_of": (C.c_int, [_P, C.c_int, _D, _D, _D, _D]),
    "gbx_lsbm_get_messages_of": (C.c_int, [_P, C.c_int, _D]),
    "gbx_lsbm_destroy": (C.c_int, [_P]),
    "gbx_lsbm_set_stream": (C.c_int, [_P, C.c_void_p]),
    "gbx_lsbm_set_options": (C.c_int, [_P, C.POINTER(LsbmOptions)]),
    "gbx_lsbm_init": (C.c_int, [_P, _D, _D, _D]),
    "gbx_lsbm_em": (C.c_int, [_P, C.POINTER(LsbmResult)]),
    "gbx_lsbm_em_batch": (C.c_int, [C.POINTER(_P), C.c_int, C.POINTER(LsbmResult), C.c_int]),
    "gbx_lsbm_iterate": (C.c_int, [_P, C.c_int, _D]),
    "gbx_lsbm_finalize": (C.c_int, [_P, C.POINTER(LsbmResult)]),
    "gbx_lsbm_get_state": (C.c_int, [_P, _D, _D, _D, _D]),
    "gbx_lsbm_get_messages": (C.c_int, [_P, _D]),
    "gbx_lsbm_launch_count": (C.c_int64, [_P]),
    "gbx_trim_cache": (C.c_int, []),
    "gbx_cached_bytes": (C.c_int64, [C.c_int]),
    "gbx_sbm_step": (C.c_int, [_P, C.c_int, C.c_int]),
    "gbx_sbm_read_residual": (C.c_int, [_P, _D]),
}

STEP_PREPARE, STEP_VERTEX_PASS, STEP_LOCAL_TOTALS, STEP_APPLY_TOTALS, STEP_THETA_LOCAL, STEP_THETA_APPLY, \
    STEP_ESCALE, STEP_FE_LOCAL, STEP_FE_APPLY, STEP_FE_FINAL, STEP_COUNT_ITERATION, STEP_UNPACK = range(12)
MODE_SWEEP, MODE_MARGINAL, MODE_LOGZ = 0, 1, 2

_lib = None


def load():
    """Load the shared object (once) and declare the prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GbxError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc != GBX_OK:
        msg = load().gbx_last_error().decode("utf-8", "replace")
        raise GbxError(f"libgraphbix_cuda error {rc}: {msg}")
