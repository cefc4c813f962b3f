"""ctypes binding of libzzb200.so (include/zzb200.h).  This is the exact stub a Julia `ccall` wrapper mirrors
(INTEGRATION.md).  There is no CPU fallback: if the library, the CUDA driver, a B200 or the kernel image is
missing, :func:`init` raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZZB200_LIB") or os.path.join(PKG_DIR, "libzzb200.so")
CUBIN_PATH = os.environ.get("ZZB200_CUBIN") or os.path.join(PKG_DIR, "zzb200_kernels.cubin")

ZZB_OK, ZZB_E_ARG, ZZB_E_CUDA, ZZB_E_BOUND, ZZB_E_GRAPH, ZZB_E_NOMEM, ZZB_E_TRACE, ZZB_E_INTERNAL = 0, 1, 2, 3, 4, 5, 6, 9
ZZB_FLAG_NO_TRACE = 1
ZZB_FLAG_LOCAL_BOUND = 2
ZZB_FLAG_STICKY = 4
ZZB_FLAG_BOOMERANG = 8
ZZB_FLAG_STICKY_REVERSIBLE = 16
ZZB_FLAG_STICKY_STRONG_UB = 32
ZZB_FLAG_STICKY_ZZ = 128
ZZB_FLAG_REFRESH = 64

EVENT_DTYPE = np.dtype([("t", "<f8"), ("i", "<i8"), ("x", "<f8"), ("theta", "<f8")])  # src/trace.jl:38

# every symbol include/zzb200.h declares
SYMBOLS = [
    "zzb_init", "zzb_shutdown", "zzb_last_error", "zzb_device_info", "zzb_event_record", "zzb_event_elapsed_ms", "zzb_problem_create_gaussian", "zzb_problem_create_logistic", "zzb_problem_free",
    "zzb_spdmp_run", "zzb_sspdmp_run", "zzb_sspdmp_adapt_run", "zzb_sspdmp3_run", "zzb_sspdmp4_run", "zzb_spdmp_boomerang_run", "zzb_spdmp_refresh_run", "zzb_run_upload_boomerang", "zzb_run_create", "zzb_run_shard", "zzb_run_ipc_export", "zzb_run_ipc_import", "zzb_run_range", "zzb_run_upload", "zzb_run_upload_kappa", "zzb_run_reset", "zzb_run_execute", "zzb_run_set", "zzb_run_stats",
    "zzb_run_fetch", "zzb_run_counts", "zzb_run_final_state", "zzb_trace_len", "zzb_trace_copy", "zzb_trace_clear", "zzb_trace_moments", "zzb_trace_sums",
    "zzb_run_discretize", "zzb_run_grid", "zzb_run_error_info", "zzb_run_free", "zzb_math_probe", "zzb_run_trace_filter", "zzb_trace_inclusion", "zzb_trace_cummean", "zzb_run_upload_refresh",
]


class ZZBError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class BoundError(ZZBError):
    """`error("Tuning parameter `c` too small.")` of src/sfact.jl:124."""


_lib = None
_inited = False


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ZZBError(ZZB_E_CUDA, f"{LIB_PATH} is missing -- run __graft_entry__.build(); zzb200 has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, f64, u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_uint32
        sig = {
            "zzb_init": [i32, vp, C.c_char_p],
            "zzb_shutdown": [],
            "zzb_last_error": [C.c_char_p, i64],
            "zzb_device_info": [vp, vp, C.c_char_p, i64],
            "zzb_event_record": [i32],
            "zzb_event_elapsed_ms": [vp],
            "zzb_problem_create_gaussian": [vp, i64] + [vp] * 8,
            "zzb_problem_create_logistic": [vp, i64, i64] + [vp] * 9 + [f64, i64] + [vp] * 4,
            "zzb_problem_free": [vp],
            "zzb_spdmp_run": [vp, f64, vp, vp, f64, vp, vp, i32, f64, u32, vp],
            "zzb_sspdmp_run": [vp, f64, vp, vp, f64, vp, vp, vp, u32, vp],
            "zzb_sspdmp_adapt_run": [vp, f64, vp, vp, f64, vp, vp, vp, i32, f64, u32, vp],
            "zzb_sspdmp3_run": [vp, vp, vp, f64, f64, f64, i32, vp, u32, vp],
            "zzb_sspdmp4_run": [vp, f64, vp, vp, f64, vp, vp, vp, u32, vp],
            "zzb_spdmp_refresh_run": [vp, f64, vp, vp, f64, vp, vp, f64, vp, i32, f64, u32, vp],
            "zzb_spdmp_boomerang_run": [vp, f64, vp, vp, f64, vp, vp, f64, f64, vp, i32, f64, u32, vp],
            "zzb_run_upload_boomerang": [vp, vp, f64, f64],
            "zzb_run_create": [vp, u32, i64, vp],
            "zzb_run_upload_kappa": [vp, vp],
            "zzb_run_upload": [vp, f64, vp, vp, vp, vp, i32, f64],
            "zzb_run_shard": [vp, i32, i32],
            "zzb_run_ipc_export": [vp, vp, i64, vp],
            "zzb_run_ipc_import": [vp, i32, vp, i64],
            "zzb_run_range": [vp, vp, vp],
            "zzb_run_reset": [vp],
            "zzb_run_execute": [vp, f64, vp],
            "zzb_run_set": [vp, C.c_char_p, f64],
            "zzb_run_stats": [vp, vp, i32],
            "zzb_run_fetch": [vp] * 10,
            "zzb_run_counts": [vp, vp, vp],
            "zzb_run_final_state": [vp] * 5,
            "zzb_trace_len": [vp, vp],
            "zzb_trace_copy": [vp, vp, i64, i64],
            "zzb_trace_clear": [vp],
            "zzb_trace_moments": [vp] * 3,
            "zzb_trace_sums": [vp] * 3,
            "zzb_run_discretize": [vp, f64, i64],
            "zzb_run_grid": [vp, vp, i64, i64, vp],
            "zzb_run_error_info": [vp] * 5,
            "zzb_math_probe": [i32, i64, vp, vp, vp, vp, vp],
            "zzb_run_trace_filter": [vp, vp, i64],
            "zzb_run_upload_refresh": [vp, vp, f64],
            "zzb_trace_inclusion": [vp, vp],
            "zzb_trace_cummean": [vp, vp, vp, vp],
            "zzb_run_free": [vp],
        }
        for name, args in sig.items():
            f = getattr(L, name)
            f.argtypes = args
            f.restype = i32
        _lib = L
    return _lib


def last_error() -> str:
    buf = C.create_string_buffer(1024)
    lib().zzb_last_error(buf, 1024)
    return buf.value.decode("utf-8", "replace")


def check(st: int):
    if st == ZZB_OK:
        return
    msg = last_error()
    if st == ZZB_E_BOUND:
        raise BoundError(st, msg)
    raise ZZBError(st, f"zzb200 error {st}: {msg}")


def init(device: int | None = None):
    """Load the driver, create the context on `device` (default: LOCAL_RANK or 0) and load the kernel image."""
    global _inited
    if _inited:
        return
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    ids = (C.c_int32 * 1)(device)
    check(lib().zzb_init(1, ids, CUBIN_PATH.encode()))
    _inited = True


def shutdown():
    global _inited
    if _inited:
        lib().zzb_shutdown()
        _inited = False


def device_info():
    init()
    sm = C.c_int32()
    mem = C.c_int64()
    name = C.create_string_buffer(128)
    check(lib().zzb_device_info(C.byref(sm), C.byref(mem), name, 128))
    return dict(sm_count=sm.value, total_mem=mem.value, name=name.value.decode())


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def f8(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def event_record(which: int):
    check(lib().zzb_event_record(int(which)))


def event_elapsed_ms() -> float:
    ms = C.c_float()
    check(lib().zzb_event_elapsed_ms(C.byref(ms)))
    return ms.value
