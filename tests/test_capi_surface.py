"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/compute_b200.h declares (no compute calls), and the Python mirror fails loudly without CUDA."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "compute_b200.h")).read()
    return sorted(set(re.findall(r"BCB_API\s+(?:const\s+char\s*\*|int)\s*(bcb_\w+)\s*\(", text)))


def test_header_declares_the_path():
    syms = _declared_symbols()
    for must in ("bcb_radix_sort", "bcb_insertion_sort", "bcb_sort_host", "bcb_scan", "bcb_reduce", "bcb_accumulate"):
        assert must in syms
    assert len(syms) >= 25


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from compute_b200 import _capi
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for s in _declared_symbols():
        assert hasattr(lib, s), f"{s} declared in compute_b200.h but not exported"
    assert set(_declared_symbols()) == set(_capi.SIGNATURES), "ctypes table and header disagree"
    # host-only entry points are callable without a device
    lib.bcb_error_string.restype = ctypes.c_char_p
    assert lib.bcb_error_string(0) == b"success"
    assert b"invalid" in lib.bcb_error_string(10001)
    assert lib.bcb_version() >= 100


def test_argument_checks_need_no_device():
    """Argument validation comes before any CUDA call: bad arguments are reported (never a crash) even on a box without a
    GPU -- the error convention of compute_b200.h."""
    from compute_b200 import _capi
    lib = _capi.lib()
    EINVAL, EUNSUPPORTED = 10001, 10002
    INT, USHORT, ULONG, DOUBLE = 4, 3, 7, 9
    buf = (ctypes.c_ubyte * 4096)()
    p = ctypes.addressof(buf)
    # field sorts: the field must lie inside the record, the projection must be identity or abs, the dtype must exist
    assert lib.bcb_sort_by_field(None, p, 100, 8, 6, INT, 0, 0) == EINVAL
    assert lib.bcb_sort_by_field(None, p, 100, 8, 0, INT, 3, 0) == EUNSUPPORTED
    assert lib.bcb_sort_by_field(None, p, 100, 8, 0, 99, 0, 0) == EINVAL
    assert lib.bcb_sort_by_field(None, p, 100, 0, 0, INT, 0, 0) == EINVAL
    res = ctypes.c_int(7)
    assert lib.bcb_is_sorted_by_field(None, p, 100, 8, 8, INT, 0, 0, ctypes.byref(res)) == EINVAL
    assert lib.bcb_is_sorted_by_field(None, p, 100, 8, 0, INT, 0, 0, None) == EINVAL
    # digit exchange: shapes outside the warp-specialised kernel are refused from the types alone
    ptrs = (ctypes.c_void_p * 256)(*([p] * 256))
    first = (ctypes.c_ulonglong * 256)()
    assert lib.bcb_radix_exchange_scatter(None, USHORT, 1, p, None, 0, 10, ptrs, None, first) == EUNSUPPORTED
    assert lib.bcb_radix_exchange_scatter(None, INT, 1, p, p, 3, 10, ptrs, ptrs, first) == EUNSUPPORTED
    assert lib.bcb_radix_exchange_scatter(None, ULONG, 1, p, p, 4, 10, ptrs, ptrs, first) == EUNSUPPORTED
    assert lib.bcb_radix_exchange_scatter(None, DOUBLE, 0, p, None, 0, 10, ptrs, None, first) == EUNSUPPORTED
    assert lib.bcb_radix_exchange_scatter(None, INT, 1, p, None, 0, 10, None, None, first) == EINVAL
    assert lib.bcb_radix_exchange_scatter(None, 99, 1, p, None, 0, 10, ptrs, None, first) == EINVAL
    seg = (ctypes.c_ulonglong * 2)(0, 50)
    ln = (ctypes.c_ulonglong * 2)(100, 200)
    assert lib.bcb_radix_sort_segments(None, INT, 1, p, None, 0, p, None, seg, ln, 2) == EINVAL        # overlapping / misaligned
    assert lib.bcb_radix_sort_segments(None, INT, 1, p, None, 0, p, None, seg, ln, 257) == EINVAL
    assert lib.bcb_radix_sort_segments(None, USHORT, 1, p, None, 0, p, None, seg, ln, 2) == EUNSUPPORTED
    assert lib.bcb_radix_sort_segments(None, INT, 1, p, None, 0, p, None, seg, ln, 0) == 0              # nothing to do
    # distributed scan: the gathered records are required
    init = ctypes.c_int(0)
    assert lib.bcb_scan_with_carry(None, INT, INT, 0, 1, p, p, 10, ctypes.byref(init), None, 1) == EINVAL
    assert lib.bcb_scan_with_carry(None, INT, INT, 0, 5, p, p, 10, ctypes.byref(init), p, 1) == EINVAL


def test_no_cpu_fallback_in_product_package():
    """The product never imports the oracle and refuses to run without CUDA."""
    pkg = os.path.join(ROOT, "compute_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f
    import torch
    if not torch.cuda.is_available():
        import compute_b200
        with pytest.raises(RuntimeError):
            compute_b200.command_queue()
