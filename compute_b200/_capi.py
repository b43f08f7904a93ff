"""ctypes binding of include/compute_b200.h.  Fails loudly when the CUDA library has not been built:
there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcompute_b200.so")

# every symbol include/compute_b200.h declares: name -> argtypes (restype is int unless noted)
_vp, _sz, _i = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
_pi = ctypes.POINTER(ctypes.c_int)
SIGNATURES = {
    "bcb_error_string": ([_i], ctypes.c_char_p),
    "bcb_version": ([], _i),
    "bcb_device_count": ([_pi], _i),
    "bcb_device_info": ([_i, ctypes.c_char_p, _sz, _pi, ctypes.POINTER(_sz), _pi, _pi], _i),
    "bcb_set_device": ([_i], _i),
    "bcb_get_device": ([_pi], _i),
    "bcb_stream_create": ([_i, ctypes.POINTER(_vp)], _i),
    "bcb_stream_destroy": ([_vp], _i),
    "bcb_stream_synchronize": ([_vp], _i),
    "bcb_malloc": ([ctypes.POINTER(_vp), _sz], _i),
    "bcb_free": ([_vp], _i),
    "bcb_host_alloc": ([ctypes.POINTER(_vp), _sz], _i),
    "bcb_host_free": ([_vp], _i),
    "bcb_host_register": ([_vp, _sz, ctypes.POINTER(_vp)], _i),
    "bcb_host_unregister": ([_vp], _i),
    "bcb_memcpy_h2d": ([_vp, _vp, _vp, _sz], _i),
    "bcb_memcpy_d2h": ([_vp, _vp, _vp, _sz], _i),
    "bcb_memcpy_d2d": ([_vp, _vp, _vp, _sz], _i),
    "bcb_fill": ([_vp, _vp, _sz, _vp, _sz], _i),
    "bcb_iota": ([_vp, _i, _vp, _sz, _vp], _i),
    "bcb_is_sorted": ([_vp, _i, _i, _vp, _sz, _pi], _i),
    "bcb_timing_enable": ([_vp, _i], _i),
    "bcb_timing_read": ([_vp, _i, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_ulonglong)], _i),
    "bcb_workspace_bytes": ([_vp, ctypes.POINTER(_sz)], _i),
    "bcb_workspace_release": ([_vp], _i),
    "bcb_radix_sort": ([_vp, _i, _i, _vp, _sz, _vp, _sz], _i),
    "bcb_radix_sort_copy": ([_vp, _i, _i, _vp, _vp, _sz, _vp, _vp, _sz], _i),
    "bcb_sort_speculation_stats": ([_vp, ctypes.POINTER(ctypes.c_ulonglong), ctypes.POINTER(ctypes.c_ulonglong)], _i),
    "bcb_is_sorted_by_radix_key": ([_vp, _i, _i, _vp, _sz, _pi], _i),
    "bcb_insertion_sort": ([_vp, _i, _i, _vp, _sz, _vp, _sz], _i),
    "bcb_sort_host": ([_vp, _i, _i, _vp, _sz], _i),
    "bcb_sort_by_field": ([_vp, _vp, _sz, _sz, _sz, _i, _i, _i], _i),
    "bcb_is_sorted_by_field": ([_vp, _vp, _sz, _sz, _sz, _i, _i, _i, _pi], _i),
    "bcb_partition_points": ([_vp, _i, _i, _vp, _sz, _vp, _sz, _vp], _i),
    "bcb_partition_by_splitters": ([_vp, _i, _i, _vp, _vp, _vp, _vp, _sz, _sz, _vp, _sz, _vp], _i),
    "bcb_radix_top_histogram": ([_vp, _i, _i, _vp, _sz, _vp], _i),
    "bcb_radix_exchange_scatter": ([_vp, _i, _i, _vp, _vp, _sz, _sz, _vp, _vp, _vp], _i),
    "bcb_radix_sort_segments": ([_vp, _i, _i, _vp, _vp, _sz, _vp, _vp, _vp, _vp, _sz], _i),
    "bcb_partition_counts": ([_vp, _i, _i, _vp, _sz, _vp, _sz, _vp], _i),
    "bcb_partition_scatter": ([_vp, _i, _i, _vp, _vp, _sz, _sz, _vp, _sz, _vp, _vp], _i),
    "bcb_ipc_export": ([_vp, _vp], _i),
    "bcb_ipc_open": ([_vp, ctypes.POINTER(_vp)], _i),
    "bcb_ipc_close": ([_vp], _i),
    "bcb_scan": ([_vp, _i, _i, _i, _i, _vp, _vp, _sz, _vp], _i),
    "bcb_scan_with_carry": ([_vp, _i, _i, _i, _i, _vp, _vp, _sz, _vp, _vp, _i], _i),
    "bcb_reduce": ([_vp, _i, _i, _i, _vp, _sz, _vp, _i], _i),
    "bcb_accumulate": ([_vp, _i, _i, _i, _i, _vp, _sz, _vp, _vp], _i),
    "bcb_transform_if": ([_vp, _i, _vp, _sz, _i, _vp, _vp, ctypes.POINTER(_sz)], _i),
    "bcb_count_if": ([_vp, _i, _vp, _sz, _vp, ctypes.POINTER(ctypes.c_ulonglong)], _i),
    "bcb_transform_reduce": ([_vp, _i, _vp, _vp, _sz, _i, _i, _vp, _i], _i),
    "bcb_reduce_by_key": ([_vp, _i, _i, _vp, _vp, _sz, _vp, _vp, _i, ctypes.POINTER(_sz)], _i),
    "bcb_set_operation": ([_vp, _i, _i, _vp, _sz, _vp, _sz, _vp, ctypes.POINTER(_sz)], _i),
    "bcb_find_extremum": ([_vp, _i, _vp, _sz, _i, ctypes.POINTER(_sz)], _i),
}

_lib = None


class ComputeError(RuntimeError):
    """Counterpart of boost::compute::opencl_error (exception/opencl_error.hpp:30-61)."""

    def __init__(self, code: int, what: str):
        super().__init__(f"{what} (code {code})")
        self.error_code = code


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m compute_b200.build` "
                "(nvcc, sm_100a).  compute_b200 has no CPU fallback."
            )
        L = ctypes.CDLL(LIB_PATH)
        for name, (argtypes, restype) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = L
    return _lib


def check(code: int) -> None:
    if code != 0:
        msg = lib().bcb_error_string(code)
        raise ComputeError(code, msg.decode() if msg else "unknown error")
