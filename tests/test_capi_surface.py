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
