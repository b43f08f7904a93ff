"""Pins the CPU oracle (oracle/oracle.c) against every golden vector transcribed from the
reference's own tests for the hot path (tests/golden/reference_vectors.json)."""
import numpy as np
import pytest

import oracle
from golden_util import case_id, check_case, load_cases

CASES = load_cases()


@pytest.mark.parametrize("case", CASES, ids=[case_id(c) for c in CASES])
def test_oracle_matches_reference_golden(case):
    check_case(case, oracle)


def test_golden_file_is_current():
    """reference_vectors.json is what make_goldens.py writes (no hand edits)."""
    import json, os, subprocess, sys, tempfile, shutil
    here = os.path.dirname(os.path.abspath(__file__))
    with tempfile.TemporaryDirectory() as d:
        shutil.copy(os.path.join(here, "golden", "make_goldens.py"), d)
        subprocess.check_call([sys.executable, os.path.join(d, "make_goldens.py")], stdout=subprocess.DEVNULL)
        a = json.load(open(os.path.join(d, "reference_vectors.json")))
    b = json.load(open(os.path.join(here, "golden", "reference_vectors.json")))
    assert a == b


# ---- properties of the restated key transform (radix_sort.hpp:100-127), SURVEY.md section 8a row R2
def _f32_bits(x):
    return int(np.array([x], dtype=np.float32).view(np.uint32)[0])


def test_radix_key_quirks():
    k = oracle.radix_key
    # ascending float: -0.0 < +0.0, -NaN first, +NaN last
    assert k("float", True, _f32_bits(-0.0)) < k("float", True, _f32_bits(0.0))
    assert k("float", True, 0xFFC00000) < k("float", True, _f32_bits(-np.inf))
    assert k("float", True, 0x7FC00000) > k("float", True, _f32_bits(np.inf))
    # descending signed: INT_MIN sorts first (negation overflow)
    assert k("int", False, 0x80000000) == 0
    assert k("int", False, 0x7FFFFFFF) == 1
    # descending float: -0.0 (0x7FFFFFFF) before +0.0 (0x80000000); ties with +-denorm_min
    assert k("float", False, _f32_bits(-0.0)) == 0x7FFFFFFF
    assert k("float", False, _f32_bits(0.0)) == 0x80000000
    assert k("float", False, 0x00000001) == k("float", False, _f32_bits(-0.0))
    assert k("float", False, 0x80000001) == k("float", False, _f32_bits(0.0))
    # descending unsigned: max - x
    assert k("uchar", False, 0x12) == 0xED
    assert k("ulong", False, 5) == 0xFFFFFFFFFFFFFFFF - 5


@pytest.mark.parametrize("dtype", oracle.DTYPES)
@pytest.mark.parametrize("descending", [False, True])
def test_radix_sort_equals_stable_sort_by_transformed_key(dtype, descending):
    """LSD radix sort == stable sort by the concatenated transformed key, independent of digit width."""
    rng = np.random.default_rng(hash((dtype, descending)) % 2**32)
    npdt = oracle.NP_DTYPES[dtype]
    n = 3000
    w = np.dtype(npdt).itemsize
    raw = rng.integers(0, 256, size=n * w, dtype=np.uint8)
    if w > 1:  # make duplicates and special patterns common
        raw[: (n // 4) * w] = np.tile(raw[:w], n // 4)
    keys = raw.view(npdt).copy()
    vals = np.arange(n, dtype=np.uint32)
    gk, gv = oracle.radix_sort(keys, descending, vals)
    ubits = keys.view({1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[w])
    tk = np.array([oracle.radix_key(dtype, not descending, int(b)) for b in ubits], dtype=np.uint64)
    order = np.argsort(tk, kind="stable")
    assert gk.tobytes() == keys[order].tobytes()
    np.testing.assert_array_equal(gv, vals[order])


def test_scan_reduce_wraparound_and_types():
    x = np.full(70000, 2**31 - 7, dtype=np.int32)
    inc = oracle.scan(x, "plus", False, 0)
    exp = np.cumsum(x.astype(np.int64)).astype(np.uint64).astype(np.uint32).view(np.int32)
    np.testing.assert_array_equal(inc, exp)
    assert oracle.reduce(x, "plus") == exp[-1]
    u8 = np.array([250, 250], dtype=np.uint8)
    assert oracle.reduce(u8, "plus") == np.uint8(244)
    assert oracle.reduce(u8, "plus", np.float32) == np.float32(500)
    assert oracle.accumulate(np.array([2, 8, 16], np.int32), np.int32(1024), "divides") == 4


def test_cpu_device_algorithms_match_serial_definitions():
    rng = np.random.default_rng(7)
    for n in (0, 1, 63, 64, 65, 513, 100_000, (1 << 21) + 12345):
        k = rng.integers(0, 2**32, size=n, dtype=np.uint32)
        ref = np.sort(k, kind="stable")
        for threads in (1, 3, 8):
            a = k.copy()
            oracle.merge_sort_on_cpu_u32(a, threads)
            np.testing.assert_array_equal(a, ref)
    x = rng.integers(0, 25, size=1_000_003, dtype=np.int32)
    for excl in (False, True):
        out = np.empty_like(x)
        oracle.scan_on_cpu_i32(x, out, excl, 5 if excl else 0, 8)
        np.testing.assert_array_equal(out, oracle.scan(x, "plus", excl, 5 if excl else 0))
    assert oracle.reduce_on_cpu_i32(x, 8) == int(oracle.reduce(x, "plus"))


def test_set_operations_two_pointer_equals_multiset_counts():
    """the two restatements of std::set_* (merge loop / multiset counts) agree, duplicates and empty ranges included"""
    rng = np.random.default_rng(3)
    for na, nb, hi in ((0, 0, 5), (0, 7, 5), (9, 0, 5), (50, 60, 8), (300, 200, 40), (257, 1000, 1000)):
        a = np.sort(rng.integers(0, hi, size=na).astype(np.int32))
        b = np.sort(rng.integers(0, hi, size=nb).astype(np.int32))
        for which in ("union", "intersection", "difference", "symmetric_difference"):
            x, y = oracle.set_operation(which, a, b), oracle.set_operation_counting(which, a, b)
            assert x.tobytes() == y.tobytes(), (which, na, nb)
