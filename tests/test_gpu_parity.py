"""GPU parity tests: the CUDA path (through the C ABI) vs the CPU oracle and the reference's golden vectors.

Bar: bit-exact (byte comparison) for every sort and for integer scan / reduce / accumulate; float sums within
|gpu - S| <= 4 * ceil(log2 N) * 2^-24 * sum|x_i| of a float64 left fold S (2^-53 for double), per element for
scans."""
import math

import numpy as np
import pytest

import oracle
from golden_util import case_id, check_case, load_cases

pytestmark = pytest.mark.gpu

CASES = load_cases()
ALL = oracle.DTYPES
NPD = oracle.NP_DTYPES


@pytest.fixture(scope="module")
def gpu():
    import gpu_api
    return gpu_api


def test_native_library_is_loaded(gpu):
    import compute_b200
    assert compute_b200.lib().bcb_version() >= 100
    with open("/proc/self/maps") as f:
        assert "libcompute_b200.so" in f.read()


@pytest.mark.parametrize("case", CASES, ids=[case_id(c) for c in CASES])
def test_gpu_matches_reference_golden(case, gpu):
    check_case(case, gpu)


def random_keys(dtype, n, seed, mode="bits"):
    rng = np.random.default_rng(seed)
    npdt = np.dtype(NPD[dtype])
    w = npdt.itemsize
    if mode == "bits":  # every bit pattern incl. NaN payloads, +-0, +-inf, denormals
        k = rng.integers(0, 256, size=n * w, dtype=np.uint8).view(npdt).copy()
        if npdt.kind == "f" and n >= 16:
            specials = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, -np.nan, np.finfo(npdt).tiny, -np.finfo(npdt).tiny,
                                 np.finfo(npdt).smallest_subnormal, -np.finfo(npdt).smallest_subnormal], dtype=npdt)
            pos = rng.integers(0, n, size=4 * len(specials))
            k[pos] = np.tile(specials, 4)
        return k
    if mode == "few":  # heavy duplicates: exercises stability and digit skew
        vals = rng.integers(0, 256, size=5 * w, dtype=np.uint8).view(npdt)
        return vals[rng.integers(0, 5, size=n)].copy()
    if mode == "equal":
        return np.full(n, rng.integers(0, 256, size=w, dtype=np.uint8).view(npdt)[0], dtype=npdt)
    if mode == "sorted":
        return np.sort(rng.integers(0, 256, size=n * w, dtype=np.uint8).view({1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[w])).view(npdt).copy()
    raise ValueError(mode)


SORT_SIZES = [2, 31, 32, 33, 100, 1000, 6143, 6144, 6145, 7679, 7680, 7681, 50_000, 300_001]


# small ranges are sorted by the one-launch kernel (small_sort_kernel); BCB_SORT_SMALL=0 (read per call) sends them through
# the multi-launch sort instead, so both families see every size
@pytest.fixture(params=["one-launch", "multi-launch"])
def small_path(request, monkeypatch):
    monkeypatch.setenv("BCB_SORT_SMALL", "1" if request.param == "one-launch" else "0")
    return request.param


@pytest.mark.parametrize("dtype", ALL)
@pytest.mark.parametrize("descending", [False, True], ids=["asc", "desc"])
def test_radix_sort_keys_bit_exact(dtype, descending, small_path, gpu):
    for n in SORT_SIZES:
        for mode in (["bits"] if n not in (1000, 50_000) else ["bits", "few", "equal", "sorted"]):
            k = random_keys(dtype, n, seed=n * 7 + len(mode), mode=mode)
            got = gpu.radix_sort(k, descending)
            exp = oracle.radix_sort(k, descending)
            assert got.tobytes() == exp.tobytes(), (dtype, descending, n, mode)


@pytest.mark.parametrize("dtype", ALL)
def test_public_sort_dispatch_bit_exact(dtype, gpu):
    """sort(): n<=32 insertion sort (native compare: +-0 ties keep input order), else radix."""
    for n in (0, 1, 2, 17, 32, 33, 500):
        for desc in (False, True):
            k = random_keys(dtype, n, seed=n + 99, mode="bits") if n else np.empty(0, NPD[dtype])
            if np.dtype(NPD[dtype]).kind == "f" and n:
                k[np.isnan(k)] = 1.5  # insertion sort with NaN is order-dependent garbage in both; keep it defined
                k[: min(n, 4)] = [0.0, -0.0, 0.0, -0.0][: min(n, 4)]
            assert gpu.sort(k, desc).tobytes() == oracle.sort(k, desc).tobytes(), (dtype, n, desc)
            assert gpu.stable_sort(k, desc).tobytes() == oracle.stable_sort(k, desc).tobytes(), (dtype, n, desc)


@pytest.mark.parametrize("key_dtype", ["uchar", "short", "int", "uint", "float", "long", "ulong", "double"])
@pytest.mark.parametrize("value_bytes", [1, 2, 4, 8, 16, 12, 3])
def test_radix_sort_pairs_bit_exact_and_stable(key_dtype, value_bytes, small_path, gpu):
    for n, mode in ((40, "few"), (5000, "few"), (7681, "bits"), (100_003, "few")):
        for desc in (False, True):
            k = random_keys(key_dtype, n, seed=n + value_bytes, mode=mode)
            rng = np.random.default_rng(n)
            v = rng.integers(0, 256, size=(n, value_bytes), dtype=np.uint8)
            v[:, :min(3, value_bytes)] = (np.arange(n)[:, None] >> (8 * np.arange(min(3, value_bytes)))) & 0xFF  # original index
            gk, gv = gpu.radix_sort(k, desc, v)
            ek, ev = oracle.radix_sort(k, desc, v)
            assert gk.tobytes() == ek.tobytes(), (key_dtype, value_bytes, n, desc)
            assert gv.tobytes() == ev.tobytes(), (key_dtype, value_bytes, n, desc)


@pytest.mark.parametrize("dtype,n", [("uint", 294_912), ("uint", 294_913), ("ulong", 196_608), ("uchar", 290_003), ("ushort", 12_289), ("float", 250_001)])
def test_one_launch_sort_near_its_size_limit(dtype, n, gpu):
    """The one-launch small sort at (and one key past) its largest size of 48 tiles, every key width (1 / 2 / 4 / 8 passes;
    8-bit keys end in the temporary buffer and are copied back), keys and a payload, repeated on the same queue (the grid
    barrier counter is monotonic across launches)."""
    k = random_keys(dtype, n, seed=n, mode="bits")
    v = np.arange(n, dtype=np.uint32)
    for desc in (False, True):
        assert gpu.radix_sort(k, desc).tobytes() == oracle.radix_sort(k, desc).tobytes(), (dtype, n, desc)
        m = min(n, 8 * 6144)   # (with a payload the one-launch sort stops at 8 tiles)
        gk, gv = gpu.radix_sort(k[:m], desc, v[:m])
        ek, ev = oracle.radix_sort(k[:m], desc, v[:m])
        assert gk.tobytes() == ek.tobytes() and gv.tobytes() == ev.tobytes(), (dtype, n, desc)


def test_sort_by_key_dispatch(gpu):
    for n in (0, 1, 5, 31, 32, 64):
        k = random_keys("int", n, seed=n, mode="few") if n else np.empty(0, np.int32)
        v = np.arange(n, dtype=np.uint32)
        for desc in (False, True):
            gk, gv = gpu.sort_by_key(k, v, desc)
            ek, ev = oracle.sort_by_key(k, v, desc)
            assert gk.tobytes() == ek.tobytes() and gv.tobytes() == ev.tobytes(), (n, desc)


@pytest.mark.parametrize("dtype", ["uchar", "ushort", "uint", "float", "ulong"])
def test_sort_unaligned_sub_range(dtype, gpu):
    """buffer_iterator offsets are arbitrary: sort [lo, hi) and leave the rest untouched."""
    n = 20_011
    k = random_keys(dtype, n, seed=5, mode="bits")
    for lo, hi in ((1, n - 1), (3, 10_000), (7, 7 + 7680), (13, 13 + 40)):
        got = gpu.sort_sub_range("radix_sort", k, lo, hi, False)
        exp = k.copy()
        exp[lo:hi] = oracle.radix_sort(k[lo:hi], False)
        assert got.tobytes() == exp.tobytes(), (dtype, lo, hi)


def test_sort_host_range(gpu):
    for n in (1, 20, 33, 100_000):
        k = random_keys("uint", n, seed=n, mode="bits")
        assert gpu.sort_host(k).tobytes() == oracle.sort(k).tobytes()
        assert gpu.sort_host(k, True).tobytes() == oracle.sort(k, True).tobytes()


def test_repeat_calls_reuse_workspace(gpu):
    """program-cache style repeat (test_radix_sort.cpp:194-201) + shrinking / growing sizes on one stream."""
    for n in (100_000, 50, 7681, 1_000_000, 33, 100_000):
        k = random_keys("uint", n, seed=n, mode="bits")
        assert gpu.radix_sort(k).tobytes() == oracle.radix_sort(k).tobytes()
        x = np.random.default_rng(n).integers(-100, 100, size=n).astype(np.int32)
        np.testing.assert_array_equal(gpu.scan(x, "plus", True, 3), oracle.scan(x, "plus", True, 3))


# ------------------------------------------------------------------------------------------ scan
SCAN_SIZES = [1, 2, 31, 255, 256, 257, 4095, 4096, 4097, 12_289, 100_000, 1_000_003]
INT_TYPES = ["char", "uchar", "short", "ushort", "int", "uint", "long", "ulong"]


def scan_input(dtype, n, seed):
    rng = np.random.default_rng(seed)
    npdt = np.dtype(NPD[dtype])
    if npdt.kind == "f":
        return rng.uniform(-1.0, 1.0, size=n).astype(npdt)
    w = npdt.itemsize
    return rng.integers(0, 256, size=n * w, dtype=np.uint8).view(npdt).copy()  # full range -> wrap-around matters


@pytest.mark.parametrize("dtype", INT_TYPES)
@pytest.mark.parametrize("op", ["plus", "multiplies", "min", "max", "bit_and", "bit_or", "bit_xor"])
def test_scan_integer_bit_exact(dtype, op, gpu):
    for n in SCAN_SIZES:
        x = scan_input(dtype, n, seed=n)
        if op == "multiplies":
            x = (x | 1).astype(x.dtype)  # keep products from collapsing to 0
        for excl, init in ((False, 0), (True, 0), (True, 7)):
            for in_place in (False, True):
                got = gpu.scan(x, op, excl, init, in_place=in_place)
                exp = oracle.scan(x, op, excl, init)
                assert got.tobytes() == exp.tobytes(), (dtype, op, n, excl, init, in_place)


@pytest.mark.parametrize("dtype", ["float", "double"])
def test_scan_float_plus_within_tolerance_and_deterministic(dtype, gpu):
    eps = 2.0 ** -24 if dtype == "float" else 2.0 ** -53
    for n in SCAN_SIZES + [5_000_000]:
        x = scan_input(dtype, n, seed=n)
        for excl in (False, True):
            got = gpu.scan(x, "plus", excl, 0.25 if excl else 0).astype(np.float64)
            pref, apref = oracle.prefix_f64(x)
            if dtype == "double":  # extended-precision yardstick: the float64 fold's own error is ~sqrt(n)*eps
                pref = np.cumsum(x.astype(np.longdouble)).astype(np.float64)
            if excl:
                ref = np.concatenate([[0.25], 0.25 + pref[:-1]])
                aref = np.concatenate([[0.25], 0.25 + apref[:-1]])
            else:
                ref, aref = pref, apref
            tol = 4 * max(1, math.ceil(math.log2(max(n, 2)))) * eps * aref + 1e-300
            assert np.all(np.abs(got - ref) <= tol), (dtype, n, excl, float(np.max(np.abs(got - ref) / tol)))
            again = gpu.scan(x, "plus", excl, 0.25 if excl else 0)
            assert again.astype(np.float64).tobytes() == got.tobytes(), "float scan must be run-to-run deterministic"


@pytest.mark.parametrize("dtype", ["float", "double"])
@pytest.mark.parametrize("op", ["min", "max"])
def test_scan_float_minmax_exact(dtype, op, gpu):
    for n in (1, 1000, 4097, 100_000):
        x = scan_input(dtype, n, seed=n)
        for excl, init in ((False, 0), (True, 0.5)):
            assert gpu.scan(x, op, excl, init).tobytes() == oracle.scan(x, op, excl, init).tobytes()


def test_scan_mixed_types(gpu):
    """arithmetic in the OUTPUT type (exclusive_scan.hpp:80-85)."""
    x = scan_input("uchar", 10_000, 3)
    for out in ("int", "float", "long"):
        got = gpu.scan(x, "plus", True, 1, out_dtype=NPD[out])
        exp = oracle.scan(x, "plus", True, 1, out_dtype=NPD[out])
        assert got.tobytes() == exp.tobytes(), out  # uchar sums < 2^24: exact in float too


# ranges of >= 1 MB take the warp-specialised kernel (scan_ws.cuh); here many rounds of it: ragged and exact tile
# multiples, every element width
@pytest.mark.parametrize("dtype,op", [("int", "plus"), ("uint", "max"), ("uchar", "plus"), ("short", "bit_xor"), ("long", "plus"),
                                     ("ulong", "min"), ("int", "multiplies")])
def test_scan_large_warp_specialised_integer_bit_exact(dtype, op, gpu):
    w = np.dtype(NPD[dtype]).itemsize
    for nbytes in ((1 << 24) + (1 << 22) + 13 * w, 9 * 24576 * 148 * 2):
        n = nbytes // w
        x = scan_input(dtype, n, seed=n % 1000)
        if op == "multiplies":
            x = (x | 1).astype(x.dtype)
        for excl, init, in_place in ((False, 0, False), (True, 7, False), (True, 3, True)):
            got = gpu.scan(x, op, excl, init, in_place=in_place)
            assert got.tobytes() == oracle.scan(x, op, excl, init).tobytes(), (dtype, op, n, excl, init, in_place)


@pytest.mark.parametrize("dtype", ["float", "double"])
def test_scan_large_warp_specialised_float(dtype, gpu):
    eps = 2.0 ** -24 if dtype == "float" else 2.0 ** -53
    w = np.dtype(NPD[dtype]).itemsize
    n = ((1 << 25) + 4444) // w
    x = scan_input(dtype, n, seed=5)
    for excl in (False, True):
        got = gpu.scan(x, "plus", excl, 0.25 if excl else 0).astype(np.float64)
        pref, apref = oracle.prefix_f64(x)
        if dtype == "double":
            pref = np.cumsum(x.astype(np.longdouble)).astype(np.float64)
        if excl:
            ref = np.concatenate([[0.25], 0.25 + pref[:-1]])
            aref = np.concatenate([[0.25], 0.25 + apref[:-1]])
        else:
            ref, aref = pref, apref
        tol = 4 * math.ceil(math.log2(n)) * eps * aref + 1e-300
        assert np.all(np.abs(got - ref) <= tol), (dtype, excl, float(np.max(np.abs(got - ref) / tol)))
        again = gpu.scan(x, "plus", excl, 0.25 if excl else 0)
        assert again.astype(np.float64).tobytes() == got.tobytes(), "float scan must be run-to-run deterministic"
    for op, init in (("min", 0.5), ("max", -0.5)):
        assert gpu.scan(x, op, True, init).tobytes() == oracle.scan(x, op, True, init).tobytes()


# ------------------------------------------------------------------------------------------ reduce / accumulate
RED_SIZES = [1, 2, 3, 33, 1023, 1024, 1025, 4099, 65_537, 1_000_003, 5_000_011]


@pytest.mark.parametrize("dtype", INT_TYPES)
@pytest.mark.parametrize("op", ["plus", "multiplies", "min", "max", "bit_and", "bit_or", "bit_xor"])
def test_reduce_integer_bit_exact(dtype, op, gpu):
    for n in RED_SIZES:
        x = scan_input(dtype, n, seed=n + 1)
        if op == "multiplies":
            x = (x | 1).astype(x.dtype)
        got = gpu.reduce(x, op)
        exp = oracle.reduce(x, op)
        assert NPD[dtype](got).tobytes() == NPD[dtype](exp).tobytes(), (dtype, op, n)


def test_reduce_unaligned_sub_range(gpu):
    import torch
    import compute_b200 as cb
    x = scan_input("int", 100_001, 9)
    d = gpu.to_dev(x)
    for lo, hi in ((1, 100_001), (3, 77_777), (5, 6), (2, 2)):
        got = cb.reduce(d[lo:hi], None, "plus")
        if lo == hi:
            assert got is None
        else:
            assert np.int32(got) == oracle.reduce(x[lo:hi], "plus")
    for dt in ("uchar", "short", "double"):
        y = scan_input(dt, 50_001, 4)
        dy = gpu.to_dev(y)
        got = cb.reduce(dy[1:], None, "max")
        assert NPD[dt](got) == oracle.reduce(y[1:], "max")


@pytest.mark.parametrize("dtype", ["float", "double"])
def test_reduce_float_within_tolerance_and_deterministic(dtype, gpu):
    eps = 2.0 ** -24 if dtype == "float" else 2.0 ** -53
    for n in RED_SIZES:
        x = np.random.default_rng(n).uniform(0, 1, size=n).astype(NPD[dtype])
        got = float(gpu.reduce(x, "plus"))
        s, a = oracle.sum_f64(x)
        if dtype == "double":
            s = math.fsum(x.tolist())  # a float64 left fold is itself off by ~sqrt(n)*eps: use the exact sum
        tol = 4 * max(1, math.ceil(math.log2(max(n, 2)))) * eps * a
        assert abs(got - s) <= tol, (dtype, n, got, s, tol)
        assert float(gpu.reduce(x, "plus")) == got
        assert NPD[dtype](gpu.reduce(x, "min")) == oracle.reduce(x, "min")
        assert NPD[dtype](gpu.reduce(x, "max")) == oracle.reduce(x, "max")
        acc = float(gpu.accumulate(x, NPD[dtype](1.5), "plus"))
        assert abs(acc - (s + 1.5)) <= tol + eps * 4


def test_reduce_mixed_result_type(gpu):
    x = scan_input("uchar", 100_000, 2)
    assert gpu.reduce(x, "plus", np.float32) == oracle.reduce(x, "plus", np.float32)  # < 2^24: exact
    assert gpu.reduce(x, "plus", np.int64) == oracle.reduce(x, "plus", np.int64)
    assert gpu.reduce(x, "plus", np.uint8) == oracle.reduce(x, "plus", np.uint8)


def test_accumulate_paths(gpu):
    x = scan_input("int", 100_003, 11)
    for init in (0, 5, -7):
        for op in ("plus", "multiplies", "min", "max", "bit_xor"):
            xx = (x | 1) if op == "multiplies" else x
            assert np.int32(gpu.accumulate(xx, np.int32(init), op)) == oracle.accumulate(xx, np.int32(init), op), (init, op)
    small = np.array([2, 8, 16], np.int32)
    assert gpu.accumulate(small, np.int32(1024), "divides") == 4
    assert gpu.accumulate(small, np.int32(100), "minus") == 74
    f = np.random.default_rng(0).uniform(0, 3, size=2000).astype(np.float32)
    # int-typed init over float data: strict serial fold with truncation each step (test_accumulate.cpp:258-268)
    assert gpu.accumulate(f, np.int32(0), "plus", op_dtype=np.float32, acc_dtype=np.int32) == \
        oracle.accumulate(f, np.int32(0), "plus", op_dtype=np.float32, acc_dtype=np.int32)
    assert gpu.accumulate(np.empty(0, np.int32), np.int32(9), "plus") == 9


# ------------------------------------------------------------------------------------------ speculative keys-only sorts
@pytest.mark.parametrize("dtype", ["uint", "int", "float", "ulong", "double"])
def test_large_keys_only_sort_speculative_path_bit_exact(dtype, gpu):
    """n >= 2^22 keys-only sorts take the speculative (verified) pass kernel; results must stay byte-exact."""
    n = (1 << 22) + 12345
    for desc, mode in ((False, "bits"), (True, "few")):
        k = random_keys(dtype, n, seed=17, mode=mode)
        assert gpu.radix_sort(k, desc).tobytes() == oracle.radix_sort(k, desc).tobytes(), (dtype, desc, mode)
    import ctypes
    import compute_b200 as cb
    runs, falls = ctypes.c_ulonglong(), ctypes.c_ulonglong()
    cb.lib().bcb_sort_speculation_stats(cb.command_queue().handle, ctypes.byref(runs), ctypes.byref(falls))
    assert runs.value >= 1 and falls.value == 0, (runs.value, falls.value)


def test_speculation_fallback_path_in_subprocess():
    """BCB_SORT_FORCE_FALLBACK=1 makes every verification 'fail': the deterministic re-sort must give the same bytes."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, ctypes, numpy as np\n"
        f"sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'tests')!r})\n"
        "import gpu_api, oracle, compute_b200 as cb\n"
        "rng = np.random.default_rng(3)\n"
        "k = rng.integers(0, 2**32, size=(1 << 22) + 777, dtype=np.uint32)\n"
        "k[::3] = k[0]\n"
        "assert gpu_api.radix_sort(k).tobytes() == oracle.radix_sort(k).tobytes()\n"
        "r, f = ctypes.c_ulonglong(), ctypes.c_ulonglong()\n"
        "cb.lib().bcb_sort_speculation_stats(cb.command_queue().handle, ctypes.byref(r), ctypes.byref(f))\n"
        "assert r.value == 1 and f.value == 1, (r.value, f.value)\n"
        "print('FALLBACK_OK')\n"
    )
    env = dict(os.environ, BCB_SORT_FORCE_FALLBACK="1")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0 and "FALLBACK_OK" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("dtype", ["uint", "float", "ulong", "short"])
def test_radix_key_sortedness_check_detects_every_seam(dtype, gpu):
    """The speculative sort trusts bcb_is_sorted_by_radix_key on its own output: one swapped pair anywhere -- inside a
    128-bit vector, across vectors, across a warp's 32 vectors, into the scalar tail, at either end -- must be seen."""
    import ctypes
    import compute_b200 as cb
    from compute_b200.core import dtype_code
    n = 70_003
    k = oracle.radix_sort(random_keys(dtype, n, seed=5, mode="bits"), False)
    q = cb.command_queue()

    def check(arr, asc=True):
        d = gpu.to_dev(arr)
        res = ctypes.c_int(-1)
        cb._capi.check(cb.lib().bcb_is_sorted_by_radix_key(q.handle, dtype_code(arr.dtype), int(asc), d.data_ptr(), arr.size,
                                                             ctypes.byref(res)))
        return bool(res.value)

    assert check(k)
    assert check(oracle.radix_sort(k, True), asc=False)
    vec = 16 // k.dtype.itemsize
    positions = [0, 1, vec - 1, vec, 32 * vec - 1, 32 * vec, 64 * vec - 1, 128 * vec - 1, 128 * vec, 256 * vec - 1, 1000, n // 2, (n // vec) * vec - 1, (n // vec) * vec, n - 2]
    bits = k.view({2: np.uint16, 4: np.uint32, 8: np.uint64}[k.dtype.itemsize])
    for p in positions:
        if p + 1 >= n:
            continue
        if oracle.radix_key(dtype, True, int(bits[p])) == oracle.radix_key(dtype, True, int(bits[p + 1])):
            continue  # swapping equal keys keeps the range sorted
        bad = k.copy()
        bad[p], bad[p + 1] = k[p + 1], k[p]
        assert not check(bad), (dtype, p)


# ------------------------------------------------------------------------------ multi-GPU building blocks, on one GPU
@pytest.mark.parametrize("dtype,value_bytes,n", [("uint", 0, 1 << 22), ("uint", 0, 100_001), ("float", 0, 7681), ("int", 8, 300_001),
                                                  ("ulong", 4, 50_000), ("short", 0, 9999), ("uchar", 0, 70_000), ("double", 12, 5000)])
def test_radix_sort_copy_leaves_source_untouched(dtype, value_bytes, n, gpu):
    import compute_b200 as cb
    from compute_b200.core import dtype_code
    k = random_keys(dtype, n, seed=21, mode="bits")
    v = np.arange(n * max(1, value_bytes), dtype=np.uint8).reshape(n, max(1, value_bytes)) if value_bytes else None
    for desc in (False, True):
        dk, dv = gpu.to_dev(k), (gpu.to_dev(v) if v is not None else None)
        ok = gpu.to_dev(np.zeros_like(k))
        ov = gpu.to_dev(np.zeros_like(v)) if v is not None else None
        q = cb.command_queue()
        cb._capi.check(cb.lib().bcb_radix_sort_copy(q.handle, dtype_code(k.dtype), int(not desc), dk.data_ptr(), ok.data_ptr(), n,
                                                    None if v is None else dv.data_ptr(), None if v is None else ov.data_ptr(),
                                                    value_bytes))
        q.finish()
        assert gpu.to_host(dk, k.dtype).tobytes() == k.tobytes()
        if v is None:
            assert gpu.to_host(ok, k.dtype).tobytes() == oracle.radix_sort(k, desc).tobytes()
        else:
            ek, ev = oracle.radix_sort(k, desc, v)
            assert gpu.to_host(ok, k.dtype).tobytes() == ek.tobytes()
            assert gpu.to_host(ov, v.dtype).reshape(v.shape).tobytes() == ev.tobytes()
            assert gpu.to_host(dv, v.dtype).reshape(v.shape).tobytes() == v.tobytes()


@pytest.mark.parametrize("dtype,value_bytes,nsplit,n", [("uint", 0, 1, 1 << 21), ("uint", 0, 7, 1_000_003), ("float", 4, 3, 200_000),
                                                         ("long", 8, 7, 77_777), ("short", 0, 2, 50_001), ("int", 0, 5, 100),
                                                         # >= 2^24 keys: the (opt-in) warp-specialised exchange kernel
                                                         ("uint", 0, 7, (1 << 24) + 12_345), ("float", 4, 3, (1 << 24) + 5),
                                                         ("int", 8, 7, 1 << 24), ("ulong", 0, 1, (1 << 24) + 3), ("double", 0, 6, (1 << 24) + 777),
                                                         ("uint", 0, 0, (1 << 24) + 1)])
def test_partition_counts_and_scatter_to_separate_buffers(dtype, value_bytes, nsplit, n, gpu, monkeypatch):
    """The multi-GPU exchange on one GPU: every bucket goes to its own destination buffer (at every 16-byte phase, like a
    slice inside a peer's receive buffer), stable, with exact counts, nothing written outside the bucket.  Shards of
    >= 2^24 keys go through the opt-in warp-specialised exchange kernel (BCB_SPLIT_WS=1), the others through the default."""
    monkeypatch.setenv("BCB_SPLIT_WS", "1" if n >= (1 << 24) else "0")
    import ctypes
    import torch
    import compute_b200 as cb
    from compute_b200 import distributed as cbd
    from compute_b200.core import dtype_code
    k = random_keys(dtype, n, seed=33, mode="bits" if dtype != "short" else "few")
    w = k.dtype.itemsize
    bits = k.view({1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[w])
    for desc in (False, True):
        tk = cbd.transformed_keys(bits, dtype_code(k.dtype), not desc)
        sp = np.sort(tk[np.random.default_rng(3).integers(0, n, size=nsplit)]).astype(np.uint64)
        if nsplit >= 3:
            sp[1] = sp[0]  # an empty bucket
        bucket = np.searchsorted(sp, tk, side="right")
        exp_counts = np.bincount(bucket, minlength=nsplit + 1)
        q = cb.command_queue()
        dk = gpu.to_dev(k)
        v = (np.arange(n * value_bytes, dtype=np.uint8).reshape(n, value_bytes)) if value_bytes else None
        dv = gpu.to_dev(v) if v is not None else None
        counts = np.zeros(nsplit + 1, dtype=np.uint64)
        cb._capi.check(cb.lib().bcb_partition_counts(q.handle, dtype_code(k.dtype), int(not desc), dk.data_ptr(), n, sp.ctypes.data, nsplit,
                                                     counts.ctypes.data))
        np.testing.assert_array_equal(counts.astype(np.int64), exp_counts)
        # bucket b starts `ph` elements into its buffer: keys and values share the element phase, as slices of one
        # receive buffer do
        phase = [(b * 3 + 1) % max(1, 16 // w) for b in range(nsplit + 1)]
        outs_k = [torch.zeros(int(c) * w + 64, dtype=torch.uint8, device="cuda") for c in exp_counts]
        outs_v = [torch.zeros(int(c) * value_bytes + 128, dtype=torch.uint8, device="cuda") for c in exp_counts]
        pk = (ctypes.c_void_p * (nsplit + 1))(*[t.data_ptr() + ph * w for t, ph in zip(outs_k, phase)])
        pv = (ctypes.c_void_p * (nsplit + 1))(*[t.data_ptr() + ph * value_bytes for t, ph in zip(outs_v, phase)])
        cb._capi.check(cb.lib().bcb_partition_scatter(q.handle, dtype_code(k.dtype), int(not desc), dk.data_ptr(),
                                                      None if v is None else dv.data_ptr(), value_bytes, n, sp.ctypes.data, nsplit, pk,
                                                      pv if v is not None else None))
        q.finish()
        for b in range(nsplit + 1):
            sel = np.flatnonzero(bucket == b)
            got = outs_k[b].cpu().numpy()
            lo = phase[b] * w
            assert got[lo: lo + sel.size * w].tobytes() == k[sel].tobytes(), (dtype, desc, b)
            assert not got[:lo].any() and not got[lo + sel.size * w:].any()  # nothing written outside the bucket
            if v is not None:
                gv = outs_v[b].cpu().numpy()
                lo = phase[b] * value_bytes
                assert gv[lo: lo + sel.size * value_bytes].tobytes() == v[sel].tobytes(), (dtype, desc, b)
                assert not gv[:lo].any() and not gv[lo + sel.size * value_bytes:].any()


@pytest.mark.parametrize("kind", ["small_ints", "two_values", "float_range", "one_hot_rest_uniform", "constant", "small_u64", "small_u16_in_u32"])
def test_hot_digit_values_are_ranked_by_ballot_bit_exact(kind, gpu):
    """Keys whose digits are far from uniform (constant high bytes of small integers, the exponent byte of floats, a
    handful of distinct keys): the warp-specialised pass ranks the lanes of the one or two most frequent digit values of
    a pass by ballot instead of same-address shared atomics.  >= 2^23 keys so that kernel runs (tests/conftest.py);
    A digit value that holds ALL keys turns the pass into a streaming copy (both kernel families); byte-exact against the
    oracle."""
    n = (1 << 23) + 7777
    rng = np.random.default_rng(5)
    if kind == "small_ints":
        k = rng.integers(0, 3000, size=n).astype(np.uint32)          # two constant digits, one with 12 values
    elif kind == "two_values":
        k = np.where(rng.random(n) < 0.7, np.uint32(0x01020304), np.uint32(0xF1F2F3F4)).astype(np.uint32)
    elif kind == "float_range":
        k = ((rng.random(n, dtype=np.float32) - np.float32(0.5)) * np.float32(1e5)).astype(np.float32)  # perf_sort_float's keys
    elif kind == "one_hot_rest_uniform":
        k = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
        k[rng.random(n) < 0.3] = np.uint32(0xABABABAB)                 # one hot value in every digit, 70 % uniform
    elif kind == "small_u64":
        k = rng.integers(0, 70000, size=(1 << 22) + 333).astype(np.uint64)   # five constant digits: streaming copies (r01 kernels)
    elif kind == "small_u16_in_u32":
        k = rng.integers(0, 65536, size=n).astype(np.uint32)                   # two constant digits (warp-specialised kernel)
    else:
        k = np.full(n, 12345, dtype=np.int32)
    for desc in (False, True):
        if kind == "float_range" and desc:
            continue  # (descending floats take the deterministic kernel: covered elsewhere)
        assert gpu.radix_sort(k, desc).tobytes() == oracle.radix_sort(k, desc).tobytes(), (kind, desc)


@pytest.mark.parametrize("record_bytes,offset,dtype,unary,n", [
    (8, 0, "int", "identity", 100_003), (8, 4, "int", "identity", 50_000), (8, 0, "int", "abs", 70_001), (16, 4, "float", "identity", 33_333),
    (12, 8, "uint", "identity", 20_011), (5, 1, "char", "abs", 9_999), (24, 8, "long", "abs", 12_345), (6, 2, "ushort", "identity", 40_000),
    (32, 16, "double", "identity", 25_000), (8, 0, "int", "identity", 2), (8, 0, "int", "identity", 33), (64, 60, "float", "abs", 4_097)])
def test_sort_by_field_matches_the_stable_comparator_sort(record_bytes, offset, dtype, unary, n, gpu):
    """sort / stable_sort / merge_sort_on_gpu with a comparator f(a.field) < f(b.field) (and ">"): records of several sizes
    (native 8 / 16-byte payloads, odd sizes through the index + gather path, an unaligned field), byte-exact against the
    stable comparator sort of the oracle; is_sorted with the same comparator before and after."""
    rng = np.random.default_rng(23)
    rec = rng.integers(0, 256, size=(n, record_bytes), dtype=np.uint8)
    npdt = np.dtype(NPD[dtype])
    if npdt.kind == "f":   # finite, non-zero floats with duplicates (the radix order of +-0 / NaN is not the comparator's)
        vals = (rng.integers(1, 500, size=n) * rng.choice([-0.5, 0.25, 3.0], size=n)).astype(npdt)
    else:
        info = np.iinfo(npdt)
        vals = rng.integers(max(info.min, -300), min(info.max, 300) + 1, size=n).astype(npdt)   # many ties: stability matters
        if npdt.kind == "i" and n > 10:
            vals[3] = info.min                                                                    # abs(INT_MIN) is 2^(w-1), unsigned
    rec[:, offset:offset + npdt.itemsize] = vals.view(np.uint8).reshape(n, npdt.itemsize)
    for desc in (False, True):
        exp = oracle.sort_by_field(rec, offset, dtype, unary, desc)
        got = gpu.sort_by_field(rec, offset, dtype, unary, desc)
        assert got.tobytes() == exp.tobytes(), (record_bytes, offset, dtype, unary, desc)
        assert gpu.is_sorted_by_field(got, offset, dtype, unary, desc)
        assert gpu.is_sorted_by_field(rec, offset, dtype, unary, desc) == oracle.is_sorted_by_field(rec, offset, dtype, unary, desc)


@pytest.mark.parametrize("dtype,op", [("int", "plus"), ("uint", "max"), ("long", "plus"), ("float", "plus"), ("double", "min"), ("int", "bit_xor"), ("short", "multiplies")])
def test_scan_with_carry_folds_the_partials_on_the_device(dtype, op, gpu):
    """bcb_scan_with_carry: one block of a block-distributed scan, seeded on the device with the partials of the blocks
    before it (records as an all-gather leaves them; an empty block in between is skipped).  Integer results equal the
    oracle's scan of the concatenated range bit for bit; float results within the scan tolerance."""
    import torch
    import compute_b200 as cb
    from compute_b200.core import dtype_code, op_code
    npdt = np.dtype(NPD[dtype])
    rng = np.random.default_rng(12)
    sizes = [5000, 0, 70_001, (1 << 20) + 17]        # blocks of four "ranks"; the second is empty
    n = sum(sizes)
    if npdt.kind == "f":
        full = rng.random(n).astype(npdt)
    elif op == "multiplies":
        full = rng.choice(np.array([1, 1, 1, -1, 3], dtype=npdt), size=n)
    else:
        full = rng.integers(-1000 if npdt.kind == "i" else 0, 1000, size=n).astype(npdt)
    cuts = np.concatenate([[0], np.cumsum(sizes)])
    lib, q = cb.lib(), cb.command_queue()
    records = np.zeros((len(sizes), 16), dtype=np.uint8)
    for r in range(len(sizes)):
        blk = full[cuts[r]:cuts[r + 1]]
        if blk.size:
            part = np.sum(blk, dtype=np.float64) if (npdt.kind == "f" and op == "plus") else oracle.reduce(blk, op)
            records[r, :npdt.itemsize] = np.array([part], dtype=npdt).view(np.uint8)
            records[r, 8] = 1
    drec = gpu.to_dev(records.reshape(-1))
    for exclusive, init in ((1, 7), (0, None)):
        exp = oracle.scan(full, op, bool(exclusive), init if init is not None else 0)
        for r in range(len(sizes)):
            blk = full[cuts[r]:cuts[r + 1]]
            if blk.size == 0:
                continue
            dx = gpu.to_dev(blk)
            out = torch.empty_like(dx)
            init_arr = np.array([init if init is not None else 0]).astype(npdt)
            cb._capi.check(lib.bcb_scan_with_carry(q.handle, dtype_code(npdt), dtype_code(npdt), op_code(op), exclusive, dx.data_ptr(),
                                                   out.data_ptr(), blk.size, init_arr.ctypes.data, drec.data_ptr(), r))
            q.finish()
            got = gpu.to_host(out, npdt)
            want = exp[cuts[r]:cuts[r + 1]]
            if npdt.kind == "f" and op == "plus":   # positive data: the bound 4 ceil(log2 n) eps sum|x| is relative to the prefix
                ref = (np.cumsum(full, dtype=np.float64) - (full if exclusive else 0) + (init or 0))[cuts[r]:cuts[r + 1]]
                np.testing.assert_allclose(got, ref, rtol=4 * 21 * np.finfo(npdt).eps, atol=1e-6, err_msg=f"{dtype} {op} rank {r}")
            else:
                assert got.tobytes() == want.tobytes(), (dtype, op, exclusive, r)


def test_exchange_and_field_entry_points_reject_bad_arguments(gpu):
    """Error behaviour of the round-2 entry points: codes, never an exception or a crash (compute_b200.h conventions)."""
    import ctypes
    import torch
    import compute_b200 as cb
    from compute_b200.core import dtype_code
    lib, q = cb.lib(), cb.command_queue()
    EINVAL, EUNSUPPORTED = 10001, 10002
    keys = torch.zeros(1024, dtype=torch.int32, device="cuda")
    buf = torch.zeros(4096, dtype=torch.int32, device="cuda")
    ptrs = np.full(256, buf.data_ptr(), dtype=np.uint64)
    first = np.zeros(256, dtype=np.uint64)
    u32, u16, u64, f64 = dtype_code(np.uint32), dtype_code(np.uint16), dtype_code(np.uint64), dtype_code(np.float64)
    ok = lib.bcb_radix_exchange_scatter(q.handle, u32, 1, keys.data_ptr(), None, 0, 1024, ptrs.ctypes.data, None, first.ctypes.data)
    assert ok == 0
    # 16-bit keys, a 3-byte payload, 64-bit keys with a payload, descending doubles: not covered -> the same answer on every rank
    assert lib.bcb_radix_exchange_scatter(q.handle, u16, 1, keys.data_ptr(), None, 0, 1024, ptrs.ctypes.data, None, first.ctypes.data) == EUNSUPPORTED
    assert lib.bcb_radix_exchange_scatter(q.handle, u32, 1, keys.data_ptr(), buf.data_ptr(), 3, 1024, ptrs.ctypes.data, ptrs.ctypes.data, first.ctypes.data) == EUNSUPPORTED
    assert lib.bcb_radix_exchange_scatter(q.handle, u64, 1, keys.data_ptr(), buf.data_ptr(), 4, 512, ptrs.ctypes.data, ptrs.ctypes.data, first.ctypes.data) == EUNSUPPORTED
    assert lib.bcb_radix_exchange_scatter(q.handle, f64, 0, keys.data_ptr(), None, 0, 512, ptrs.ctypes.data, None, first.ctypes.data) == EUNSUPPORTED
    bad = ptrs.copy(); bad[7] += 4   # a destination that is not 16-byte aligned
    assert lib.bcb_radix_exchange_scatter(q.handle, u32, 1, keys.data_ptr(), None, 0, 1024, bad.ctypes.data, None, first.ctypes.data) == EINVAL
    assert lib.bcb_radix_exchange_scatter(q.handle, u32, 1, keys.data_ptr() + 4, None, 0, 1000, ptrs.ctypes.data, None, first.ctypes.data) == EINVAL
    assert lib.bcb_radix_exchange_scatter(q.handle, u32, 1, keys.data_ptr(), None, 0, 1024, None, None, first.ctypes.data) == EINVAL
    out = torch.zeros(4096, dtype=torch.int32, device="cuda")
    sb, sl = np.array([0, 512], dtype=np.uint64), np.array([100, 200], dtype=np.uint64)
    assert lib.bcb_radix_sort_segments(q.handle, u32, 1, buf.data_ptr(), None, 0, out.data_ptr(), None, sb.ctypes.data, sl.ctypes.data, 2) == 0
    sb2 = np.array([0, 50], dtype=np.uint64)    # overlapping, and not 16-byte aligned
    assert lib.bcb_radix_sort_segments(q.handle, u32, 1, buf.data_ptr(), None, 0, out.data_ptr(), None, sb2.ctypes.data, sl.ctypes.data, 2) == EINVAL
    sb3 = np.array([0, 513], dtype=np.uint64)
    assert lib.bcb_radix_sort_segments(q.handle, u32, 1, buf.data_ptr(), None, 0, out.data_ptr(), None, sb3.ctypes.data, sl.ctypes.data, 2) == EINVAL
    assert lib.bcb_radix_sort_segments(q.handle, u32, 1, buf.data_ptr(), None, 0, None, None, sb.ctypes.data, sl.ctypes.data, 2) == EINVAL
    assert lib.bcb_radix_sort_segments(q.handle, u16, 1, buf.data_ptr(), None, 0, out.data_ptr(), None, sb.ctypes.data, sl.ctypes.data, 2) == EUNSUPPORTED
    assert lib.bcb_radix_sort_segments(q.handle, u32, 1, buf.data_ptr(), None, 0, out.data_ptr(), None, sb.ctypes.data, sl.ctypes.data, 0) == 0  # nothing to do
    # field sorts: the field must lie inside the record; only identity / abs
    rec = torch.zeros((100, 8), dtype=torch.uint8, device="cuda")
    assert lib.bcb_sort_by_field(q.handle, rec.data_ptr(), 100, 8, 6, dtype_code(np.int32), 0, 0) == EINVAL
    assert lib.bcb_sort_by_field(q.handle, rec.data_ptr(), 100, 8, 0, dtype_code(np.int32), 3, 0) == EUNSUPPORTED   # square
    assert lib.bcb_sort_by_field(q.handle, rec.data_ptr(), 100, 8, 0, 99, 0, 0) == EINVAL
    assert lib.bcb_sort_by_field(q.handle, None, 1, 8, 0, dtype_code(np.int32), 0, 0) == 0                           # n < 2: nothing to do
    res = ctypes.c_int(0)
    assert lib.bcb_is_sorted_by_field(q.handle, rec.data_ptr(), 100, 8, 4, dtype_code(np.int32), 2, 1, ctypes.byref(res)) == 0 and res.value == 1
    q.finish()


@pytest.mark.parametrize("dtype,n", [("uint", (1 << 23) + 12345), ("double", (1 << 22) + 99), ("short", (1 << 24) + 7)])
def test_sort_host_on_a_large_pageable_range(dtype, n, gpu, monkeypatch):
    """sort(host_first, host_last) on plain malloc'ed memory of >= 32 MB: the library stages the copies itself (several
    host threads through pinned slots, runtime.cu) -- same bytes as the oracle, also with the staging switched off, and
    for a range that is not a multiple of the slot size."""
    k = random_keys(dtype, n, seed=17, mode="bits")
    exp = oracle.sort(k)
    assert gpu.sort_host(k).tobytes() == exp.tobytes()
    assert gpu.sort_host(k, True).tobytes() == oracle.sort(k, True).tobytes()
    monkeypatch.setenv("BCB_STAGED_COPY", "0")
    assert gpu.sort_host(k).tobytes() == exp.tobytes()


@pytest.mark.parametrize("value_bytes", [0, 4, 8])
def test_hot_digit_values_with_the_deterministic_ranking(value_bytes, gpu):
    """The same for the kernels that rank deterministically: key-value sorts (stability of the payload inside the long
    runs of equal keys) and descending float keys (non-injective transform), >= 2^23 elements."""
    n = (1 << 23) + 4099
    rng = np.random.default_rng(6)
    if value_bytes:
        k = rng.integers(0, 700, size=n).astype(np.int32)              # three constant digits
        if value_bytes == 4:
            k[rng.random(n) < 0.01] = -5                                # ... with a few stragglers in another value (no copy pass)
        v = np.arange(n * (value_bytes // 4), dtype=np.uint32).reshape(n, value_bytes // 4)
        for desc in (False, True):
            gk, gv = gpu.radix_sort(k, desc, v)
            ek, ev = oracle.radix_sort(k, desc, v)
            assert gk.tobytes() == ek.tobytes() and gv.tobytes() == ev.tobytes(), (value_bytes, desc)
    else:
        k = rng.integers(0, 40, size=n).astype(np.float32)              # a handful of exponents, constant low mantissa bytes
        k[::97] = -0.0
        k[1::97] = np.finfo(np.float32).smallest_subnormal             # collides with -0.0 when descending
        assert gpu.radix_sort(k, True).tobytes() == oracle.radix_sort(k, True).tobytes()


@pytest.mark.parametrize("dtype,value_bytes,world,n,desc", [
    ("uint", 0, 2, 300_001, False), ("uint", 0, 8, (1 << 22) + 4321, False), ("int", 0, 3, 1_000_003, True), ("float", 0, 4, 777_777, False),
    ("float", 0, 2, 500_000, True), ("uint", 4, 4, 600_001, False), ("int", 8, 2, 400_003, True), ("float", 4, 3, 250_000, True),
    ("ulong", 0, 2, 500_009, False), ("long", 0, 4, 300_000, True), ("double", 0, 3, 400_001, False),
    ("uint", 0, 2, 1000, False), ("uint", 8, 5, 5000, False)])
def test_digit_exchange_on_one_gpu_matches_oracle(dtype, value_bytes, world, n, desc, gpu):
    """The multi-GPU sort whose exchange is its pass over the most significant digit, with all ranks emulated on one GPU:
    every shard's bcb_radix_exchange_scatter writes its digit runs into the owners' receive buffers (placement from
    digit_exchange_plan), every owner's bcb_radix_sort_segments sorts its segments by the remaining digits.  The
    concatenation must equal the oracle's stable sort of the whole range byte for byte (keys and payload), and nothing
    may be written outside the segments."""
    import ctypes
    import torch
    import compute_b200 as cb
    from compute_b200 import distributed as cbd
    from compute_b200.core import dtype_code
    k = random_keys(dtype, n, seed=91, mode="bits")
    if dtype in ("float", "double") and desc:
        k[::50] = -0.0
        k[1::50] = np.finfo(k.dtype).smallest_subnormal  # collides with -0.0 in the reference's descending transform
    w = k.dtype.itemsize
    code = dtype_code(k.dtype)
    v = np.arange(n * (value_bytes // 4), dtype=np.uint32).reshape(n, value_bytes // 4) if value_bytes else None
    cuts = [n * r // world for r in range(world + 1)]
    cuts[1] = min(cuts[1], 7) if world > 2 else cuts[1]   # a tiny shard
    bits = k.view({4: np.uint32, 8: np.uint64}[w])
    tk = cbd.transformed_keys(bits, code, not desc)
    allh = np.stack([np.bincount((tk[cuts[r]:cuts[r + 1]] >> np.uint64(8 * w - 8)).astype(np.int64), minlength=256) for r in range(world)])
    plan = cbd.digit_exchange_plan(allh, world, max_imbalance=1e9)
    owner, first, seg_begin, seg_len, recv, span, _ = plan
    q = cb.command_queue()
    lib = cb.lib()
    guard = 64
    rk = [torch.full((int(span[d]) * w + 2 * guard,), 0xAB, dtype=torch.uint8, device="cuda") for d in range(world)]
    rv = [torch.full((int(span[d]) * value_bytes + 2 * guard,), 0xAB, dtype=torch.uint8, device="cuda") for d in range(world)]
    for r in range(world):
        dk = gpu.to_dev(k[cuts[r]:cuts[r + 1]])
        dv = gpu.to_dev(v[cuts[r]:cuts[r + 1]]) if v is not None else None
        pk = (ctypes.c_void_p * 256)(*[rk[int(owner[g])].data_ptr() + guard for g in range(256)])
        pv = (ctypes.c_void_p * 256)(*[rv[int(owner[g])].data_ptr() + guard for g in range(256)])
        df = np.ascontiguousarray(first[r], dtype=np.uint64)
        before = dk.clone()
        cb._capi.check(lib.bcb_radix_exchange_scatter(q.handle, code, int(not desc), dk.data_ptr(), None if v is None else dv.data_ptr(),
                                                      value_bytes, dk.shape[0], pk, pv if v is not None else None, df.ctypes.data))
        q.finish()
        assert torch.equal(dk.view(torch.uint8), before.view(torch.uint8))  # the shard is not modified
    # the receive buffers: guards and the alignment gaps between segments untouched
    for d in range(world):
        got = rk[d].cpu().numpy()
        used = np.zeros(got.size, dtype=bool)
        for g in np.flatnonzero(owner == d):
            used[guard + int(seg_begin[g]) * w: guard + int(seg_begin[g] + seg_len[g]) * w] = True
        assert np.all(got[~used] == 0xAB), (d, "keys written outside the segments")
    outs_k, outs_v = [], []
    for d in range(world):
        ok = torch.empty(int(recv[d]) * w, dtype=torch.uint8, device="cuda")
        ov = torch.empty(int(recv[d]) * value_bytes, dtype=torch.uint8, device="cuda")
        mine = owner == d
        sb = np.ascontiguousarray(seg_begin[mine], dtype=np.uint64)
        sl = np.ascontiguousarray(seg_len[mine], dtype=np.uint64)
        cb._capi.check(lib.bcb_radix_sort_segments(q.handle, code, int(not desc), rk[d].data_ptr() + guard,
                                                   (rv[d].data_ptr() + guard) if v is not None else None, value_bytes, ok.data_ptr(),
                                                   ov.data_ptr() if v is not None else None, sb.ctypes.data, sl.ctypes.data, sb.size))
        q.finish()
        outs_k.append(ok.cpu().numpy())
        outs_v.append(ov.cpu().numpy())
    if v is None:
        exp_k = oracle.radix_sort(k, desc)
    else:
        exp_k, exp_v = oracle.radix_sort(k, desc, v)
        assert np.concatenate(outs_v).tobytes() == exp_v.tobytes(), (dtype, value_bytes, world, "payload")
    assert np.concatenate(outs_k).tobytes() == exp_k.tobytes(), (dtype, value_bytes, world, "keys")


@pytest.mark.parametrize("dtype", ["float", "double"])
def test_large_descending_float_sort_with_colliding_keys_stays_deterministic(dtype, gpu):
    """Descending float keys: -0.0 / +denorm_min (and +0.0 / -denorm_min) share a transformed key in the reference
    (radix_sort.hpp:100-127), so their output order is decided by stability alone and a sortedness check could not
    see a swap.  Such sorts must not take the speculative path, and must still match the oracle byte for byte."""
    import ctypes
    import compute_b200 as cb
    npdt = np.dtype(NPD[dtype])
    tiny = np.finfo(npdt).smallest_subnormal
    pool = np.array([-0.0, tiny, 0.0, -tiny, 1.0, -1.0], dtype=npdt)
    n = (1 << 22) + 77
    k = pool[np.random.default_rng(9).integers(0, pool.size, size=n)]
    runs0, falls = ctypes.c_ulonglong(), ctypes.c_ulonglong()
    q = cb.command_queue()
    cb.lib().bcb_sort_speculation_stats(q.handle, ctypes.byref(runs0), ctypes.byref(falls))
    assert gpu.radix_sort(k, True).tobytes() == oracle.radix_sort(k, True).tobytes()
    runs1 = ctypes.c_ulonglong()
    cb.lib().bcb_sort_speculation_stats(q.handle, ctypes.byref(runs1), ctypes.byref(falls))
    assert runs1.value == runs0.value  # no speculative run for the descending order
    assert gpu.radix_sort(k, False).tobytes() == oracle.radix_sort(k, False).tobytes()
    cb.lib().bcb_sort_speculation_stats(q.handle, ctypes.byref(runs1), ctypes.byref(falls))
    assert runs1.value == runs0.value + 1 and falls.value == 0  # ascending: injective transform, speculative + verified


@pytest.mark.parametrize("dtype", ["uchar", "char", "short", "ushort"])
def test_large_narrow_key_sorts_use_the_column_histogram(dtype, gpu):
    """8- and 16-bit keys never speculate, but from 2^22 keys on they share the lane-column histogram kernel."""
    n = (1 << 22) + 5
    k = random_keys(dtype, n, seed=23, mode="bits")
    for desc in (False, True):
        assert gpu.radix_sort(k, desc).tobytes() == oracle.radix_sort(k, desc).tobytes(), (dtype, desc)


# ------------------------------------------------------------------ round 2: warp-specialised pass, async contract, arenas
@pytest.mark.parametrize("dtype", ["uint", "int", "float", "ulong", "double"])
def test_warp_specialised_pass_bit_exact(dtype, gpu):
    """n >= 2^23 keys-only sorts of 32/64-bit keys take onesweep_ws (bulk-copy stores, ticketed tiles, look-back in helper
    warps): several tiles per CTA, a partial last tile, heavy duplicates (runs shorter than a 16-byte chunk / empty runs)."""
    n = (1 << 23) + 4321
    for desc, mode in ((False, "bits"), (True, "few"), (False, "sorted")):
        if desc and dtype in ("float", "double"):
            continue  # descending float keys never speculate (covered by the colliding-keys test)
        k = random_keys(dtype, n, seed=23, mode=mode)
        assert gpu.radix_sort(k, desc).tobytes() == oracle.radix_sort(k, desc).tobytes(), (dtype, desc, mode)


def test_multi_round_persistent_sort_2_26_keys_and_pairs(gpu):
    """2^26 elements: ~10 tiles per CTA in the warp-specialised kernel (keys only) and ~37 rounds of the persistent
    deterministic kernel (pairs): ticket wrap-around inside a launch, long look-back chains, epoch reuse across passes."""
    n = (1 << 26) + 7
    rng = np.random.default_rng(26)
    k = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    assert gpu.radix_sort(k).tobytes() == oracle.radix_sort(k).tobytes()
    v = np.arange(n, dtype=np.uint32)
    k &= np.uint32(0x000fffff)  # many equal keys: the payload order is decided by stability alone
    gk, gv = gpu.radix_sort(k, False, v)
    ok, ov = oracle.radix_sort(k, False, v)
    assert gk.tobytes() == ok.tobytes() and gv.tobytes() == ov.tobytes()


def test_default_thresholds_pick_the_measured_kernels():
    """Without the test hooks: 2^23 keys sort deterministically (no speculation), 2^24 keys speculate (two-sweep), and a
    2^27-key sort goes through the warp-specialised kernel; all verified against torch.sort."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, ctypes, torch\n"
        f"sys.path.insert(0, {root!r})\n"
        "import compute_b200 as cb\n"
        "def runs():\n"
        "    r, f = ctypes.c_ulonglong(), ctypes.c_ulonglong()\n"
        "    cb.lib().bcb_sort_speculation_stats(cb.command_queue().handle, ctypes.byref(r), ctypes.byref(f))\n"
        "    return r.value, f.value\n"
        "for lg, want in ((23, 0), (24, 1), (27, 2)):\n"
        "    k = torch.randint(0, 2**31 - 1, (1 << lg,), dtype=torch.int32, device='cuda')\n"
        "    ref = torch.sort(k).values\n"
        "    cb.radix_sort(k)\n"
        "    torch.cuda.synchronize()\n"
        "    assert torch.equal(k, ref), lg\n"
        "    assert runs() == (want, 0), (lg, runs())\n"
        "print('THRESHOLDS_OK')\n"
    )
    env = {k: v for k, v in os.environ.items() if k not in ("BCB_SORT_SPEC_MIN_LOG2", "BCB_SORT_WS_MIN_LOG2")}
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0 and "THRESHOLDS_OK" in out.stdout, out.stdout + out.stderr


def test_sort_is_enqueue_and_return(gpu):
    """perf_sort.cpp:38-39 enqueues the sort and calls queue.finish(): bcb_radix_sort must not wait for the device, even
    on the speculative path (the verification and the gated fallback stay on the device)."""
    import torch
    import compute_b200 as cb
    n = 1 << 27
    k = torch.randint(0, 2**31 - 1, (n,), dtype=torch.int32, device="cuda").view(torch.uint32)
    cb.radix_sort(k)  # warm-up: scratch allocation, function attributes
    torch.cuda.synchronize()
    k2 = torch.randint(0, 2**31 - 1, (n,), dtype=torch.int32, device="cuda").view(torch.uint32)
    torch.cuda.synchronize()
    cb.radix_sort(k2)
    still_running = not torch.cuda.current_stream().query()
    torch.cuda.synchronize()
    assert still_running, "bcb_radix_sort blocked until the sort was finished"
    assert cb.is_sorted(k2.view(torch.int32))  # values < 2^31: same order as unsigned


def test_two_queues_sort_concurrently_from_two_threads(gpu):
    """Forward progress must not depend on the whole grid of one sort being resident: two command queues of one device,
    driven from two host threads, each sorting 2^26 keys (warp-specialised kernel) and 2^22 pairs (persistent
    deterministic kernel) at the same time."""
    import threading
    import torch
    import compute_b200 as cb
    results = {}

    def work(idx):
        stream = torch.cuda.Stream()
        q = cb.command_queue(stream)
        with torch.cuda.stream(stream):
            g = torch.Generator(device="cuda").manual_seed(100 + idx)
            k = torch.randint(0, 2**31 - 1, (1 << 26,), dtype=torch.int32, device="cuda", generator=g)
            pk = torch.randint(0, 1000, (1 << 22,), dtype=torch.int32, device="cuda", generator=g)
            pv = torch.arange(1 << 22, dtype=torch.int32, device="cuda")
            ref = torch.sort(k).values
            pref = torch.sort(pk, stable=True)
            for _ in range(3):
                kk, pkk, pvv = k.clone(), pk.clone(), pv.clone()
                cb.radix_sort(kk, True, q)
                cb.radix_sort_by_key(pkk, pvv, True, q)
            stream.synchronize()
            results[idx] = bool(torch.equal(kk, ref)) and bool(torch.equal(pkk, pref.values)) and bool(torch.equal(pvv.long(), pref.indices))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not any(t.is_alive() for t in threads), "concurrent sorts did not finish (look-back waiting for a CTA that is not resident?)"
    assert results == {0: True, 1: True}


def test_scan_descriptor_arenas_do_not_alias(gpu):
    """Round-1 advisor findings: (a) two 8-byte scans of different data on a FRESH queue (the second launch used to draw
    the first one's epoch again after a look-back reallocation), (b) an 8-byte scan right after a u32 sort on the same
    stream (its status words used to overlay the sort's digit counts)."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'tests')!r})\n"
        "import gpu_api, oracle\n"
        "rng = np.random.default_rng(5)\n"
        "for i in range(3):\n"
        "    x = rng.integers(-1000, 1000, size=(1 << 20) + 5 * i).astype(np.float64)\n"
        "    assert np.array_equal(gpu_api.scan(x, 'plus', False), np.cumsum(x)), ('double scan', i)\n"
        "for i in range(4):\n"
        "    k = rng.integers(0, 64, size=1 << 21, dtype=np.uint32)\n"
        "    assert np.array_equal(gpu_api.radix_sort(k), np.sort(k))\n"
        "    y = rng.integers(-2**40, 2**40, size=(1 << 21) + 3, dtype=np.int64)\n"
        "    assert np.array_equal(gpu_api.scan(y, 'plus', False), np.cumsum(y)), ('long scan after sort', i)\n"
        "print('ARENAS_OK')\n"
    )
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ARENAS_OK" in out.stdout, out.stdout + out.stderr


# ------------------------------------------------------------------ round 2: callers of scan / reduce
PREDS = [("none", 0, "lt", 5), ("mul", 2, "ge", 10), ("mod", 2, "eq", 1), ("mod", 3, "ne", 0), ("and", 5, "eq", 4), ("add", 3, "gt", 0),
         ("sub", 7, "le", -2), ("none", 0, "true", 0)]


@pytest.mark.parametrize("dtype", ["char", "uchar", "short", "int", "uint", "long", "ulong", "float", "double"])
def test_transform_if_copy_if_count_if_bit_exact(dtype, gpu):
    """One-pass stable compaction vs the serial definition (transform_if.hpp:42-117): empty, one tile, many tiles
    (look-back), nothing / everything selected, the output tail past the count stays untouched."""
    npdt = np.dtype(NPD[dtype])
    rng = np.random.default_rng(31)
    for n in (0, 1, 31, 4096, 4097, 100_003, 1_500_001):
        if npdt.kind == "f":
            x = (rng.integers(-40, 40, size=n) / 2.0).astype(npdt)
        else:
            lo = -20 if npdt.kind == "i" else 0
            x = rng.integers(lo, 20, size=n).astype(npdt)
        for pred in PREDS:
            if npdt.kind == "f" and pred[0] in ("mod", "and"):
                continue
            if npdt.kind == "u" and pred[3] < 0:
                continue
            fn = ["identity", "negate", "abs", "square"][(n + len(pred[0])) % 4]
            exp = oracle.transform_if(x, fn, pred)
            out, count = gpu.transform_if(x, fn, pred, fill=77)
            assert count == exp.size == gpu.count_if(x, pred) == oracle.count_if(x, pred), (dtype, n, pred)
            assert out[:count].tobytes() == exp.tobytes(), (dtype, n, pred, fn)
            assert np.all(out[count:] == npdt.type(77)), "wrote past the selected count"


@pytest.mark.parametrize("dtype", ["char", "ushort", "int", "uint", "long", "ulong"])
def test_transform_reduce_inner_product_integer_bit_exact(dtype, gpu):
    npdt = np.dtype(NPD[dtype])
    rng = np.random.default_rng(32)
    for n in (1, 255, 100_003, 3_000_001):
        x = rng.integers(0, 2**(8 * npdt.itemsize), size=n, dtype=np.uint64).astype({1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[npdt.itemsize]).view(npdt)
        y = rng.integers(0, 2**(8 * npdt.itemsize), size=n, dtype=np.uint64).astype({1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[npdt.itemsize]).view(npdt)
        for t in ("identity", "negate", "abs", "square"):
            for op in ("plus", "min", "max"):
                assert gpu.transform_reduce(x, t, op) == oracle.transform_reduce(x, t, op), (dtype, n, t, op)
        for t in ("multiplies", "plus", "minus"):
            assert gpu.transform_reduce(x, t, "plus", y) == oracle.transform_reduce(x, t, "plus", y), (dtype, n, t)
        assert gpu.inner_product(x, y, 5) == oracle.inner_product(x, y, 5)
    assert gpu.transform_reduce(np.empty(0, npdt), "abs", "plus") is None
    assert gpu.inner_product(np.empty(0, npdt), np.empty(0, npdt), 9) == npdt.type(9)


@pytest.mark.parametrize("dtype", ["float", "double"])
def test_transform_reduce_inner_product_float_tolerance(dtype, gpu):
    npdt = np.dtype(NPD[dtype])
    eps = 2.0**-24 if dtype == "float" else 2.0**-53
    rng = np.random.default_rng(33)
    n = 2_000_003
    x = rng.uniform(-1, 1, size=n).astype(npdt)
    y = rng.uniform(-1, 1, size=n).astype(npdt)
    tol = 4 * math.ceil(math.log2(n)) * eps
    exact = float(np.sum(np.abs(x.astype(np.float64))))
    assert abs(float(gpu.transform_reduce(x, "abs", "plus")) - exact) <= tol * exact
    prod = x.astype(np.float64) * y.astype(np.float64)
    got = float(gpu.inner_product(x, y, 0.5))
    assert abs(got - (0.5 + float(prod.sum()))) <= tol * float(np.abs(prod).sum()) + eps
    assert float(gpu.transform_reduce(x, "square", "max")) == float(np.max(x * x))
    assert gpu.transform_reduce(x, "identity", "plus", None) == gpu.transform_reduce(x, "identity", "plus", None)  # deterministic


@pytest.mark.parametrize("key_dtype", ["uchar", "int", "ulong", "float"])
@pytest.mark.parametrize("val_dtype", ["int", "uint", "long", "float", "double"])
def test_reduce_by_key_matches_serial_definition(key_dtype, val_dtype, gpu):
    """Segmented one-pass reduction vs the serial definition: runs inside a thread, across threads / warps / tiles
    (carry through the look-back, including tiles without any segment head), single run, all runs of length 1."""
    kd, vd = np.dtype(NPD[key_dtype]), np.dtype(NPD[val_dtype])
    rng = np.random.default_rng(34)
    for n, mean_run in ((1, 1), (2047, 3), (2048, 1), (2049, 5000), (300_001, 40), (1_000_003, 30_000), (500_000, 1)):
        runs = np.maximum(1, rng.geometric(1.0 / mean_run, size=n // mean_run + 2))
        ids = np.repeat(np.arange(runs.size), runs)[:n]
        palette = rng.permutation(251)  # key values repeat along the range (non-adjacent equal keys stay separate runs)
        keys = palette[ids % 251].astype(kd)
        if vd.kind == "f":
            vals = (rng.integers(-8, 9, size=n) / 4.0).astype(vd)  # exactly representable: sums are order independent
        else:
            vals = rng.integers(-9 if vd.kind == "i" else 0, 10, size=n).astype(vd)
        for op in ("plus", "min", "max"):
            ek, ev = oracle.reduce_by_key(keys, vals, op)
            gk, gv = gpu.reduce_by_key(keys, vals, op)
            assert gk.tobytes() == ek.tobytes(), (key_dtype, val_dtype, n, mean_run, op)
            assert gv.tobytes() == ev.tobytes(), (key_dtype, val_dtype, n, mean_run, op)
    gk, gv = gpu.reduce_by_key(np.empty(0, kd), np.empty(0, vd))
    assert gk.size == 0 and gv.size == 0


@pytest.mark.parametrize("dtype", ["char", "ushort", "int", "ulong", "float"])
def test_sort_callers_is_permutation_and_sort_by_transform(dtype, gpu):
    """is_permutation.hpp:43-67 and experimental/sort_by_transform.hpp:26-63 on top of the sort path."""
    npdt = np.dtype(NPD[dtype])
    rng = np.random.default_rng(41)
    for n in (1, 33, 5000, 300_001):
        x = rng.integers(-100 if npdt.kind != "u" else 0, 100, size=n).astype(npdt)
        y = rng.permutation(x)
        assert gpu.is_permutation(x, y) is True
        if n > 1:
            z = y.copy()
            z[n // 2] = z[n // 2] + npdt.type(1)
            assert gpu.is_permutation(x, z) is oracle.is_permutation(x, z) is False
        for fn in ("abs", "negate", "square"):
            assert gpu.sort_by_transform(x, fn).tobytes() == oracle.sort_by_transform(x, fn).tobytes(), (dtype, n, fn)
        assert gpu.sort_by_transform(x, "abs", True).tobytes() == oracle.sort_by_transform(x, "abs", True).tobytes()


# ------------------------------------------------------------------ round 2: set operations on sorted ranges, extrema
@pytest.mark.parametrize("which", ["union", "intersection", "difference", "symmetric_difference"])
@pytest.mark.parametrize("dtype", ["int", "uchar", "ulong", "float", "short"])
def test_set_operations_match_oracle(which, dtype, gpu):
    rng = np.random.default_rng(11)
    npdt = NPD[dtype]
    for na, nb, hi in ((0, 0, 5), (0, 9, 5), (7, 0, 5), (1, 1, 2), (100, 120, 10), (5_000, 3_000, 200), (70_001, 50_003, 30_011), (300_000, 1, 1000)):
        if np.dtype(npdt).kind == "f":
            a = np.sort((rng.integers(0, hi, size=na) * 0.5 - hi / 4).astype(npdt))
            b = np.sort((rng.integers(0, hi, size=nb) * 0.5 - hi / 4).astype(npdt))
        else:
            top = min(hi, int(np.iinfo(npdt).max))
            a = np.sort(rng.integers(0, top, size=na).astype(npdt))
            b = np.sort(rng.integers(0, top, size=nb).astype(npdt))
        got = gpu.set_operation(which, a, b)
        exp = oracle.set_operation(which, a, b) if na + nb <= 10_000 else oracle.set_operation_counting(which, a, b)
        assert got.tobytes() == exp.tobytes(), (which, dtype, na, nb)


def test_set_union_large_disjoint_and_identical(gpu):
    n = 3_000_000
    a = np.arange(0, 2 * n, 2, dtype=np.int32)
    b = np.arange(1, 2 * n, 2, dtype=np.int32)
    assert np.array_equal(gpu.set_operation("union", a, b), np.arange(2 * n, dtype=np.int32))
    assert gpu.set_operation("intersection", a, b).size == 0
    assert np.array_equal(gpu.set_operation("symmetric_difference", a, a), a[:0])
    assert np.array_equal(gpu.set_operation("difference", a, a[::2].copy()), a[1::2])


@pytest.mark.parametrize("dtype", ["char", "uchar", "short", "int", "uint", "long", "ulong", "float", "double"])
def test_extrema_first_occurrence(dtype, gpu):
    rng = np.random.default_rng(5)
    npdt = NPD[dtype]
    for n in (1, 2, 33, 1000, 4097, 65_537, 1_000_003, 5_000_011):
        if np.dtype(npdt).kind == "f":
            x = rng.integers(-50, 50, size=n).astype(npdt) / 4
        else:
            info = np.iinfo(npdt)
            x = rng.integers(max(info.min, -50), min(info.max, 50) + 1, size=n).astype(npdt)  # few distinct values: ties everywhere
        assert gpu.min_element(x) == oracle.min_element(x), (dtype, n)
        assert gpu.max_element(x) == oracle.max_element(x), (dtype, n)
    x = np.zeros(100_000, dtype=npdt)
    assert gpu.min_element(x) == 0 and gpu.max_element(x) == 0
    x[77_777] = 1
    assert gpu.max_element(x) == 77_777 and gpu.min_element(x) == 0


@pytest.mark.parametrize("dtype", ["uchar", "short", "int", "uint", "float", "long", "double"])
def test_top_digit_histogram_matches_transformed_keys(dtype, gpu):
    """bcb_radix_top_histogram (first step of the multi-GPU sort's histogram plan): 256 bins of the most significant digit
    of the TRANSFORMED key, both orders, small (plain kernel) and large (lane-column kernel) ranges."""
    import ctypes
    import compute_b200 as cb
    from compute_b200 import distributed as cbd
    from compute_b200.core import dtype_code
    npdt = np.dtype(NPD[dtype])
    w = npdt.itemsize
    for n in (0, 1, 1000, 100_003, (1 << 22) + 77):
        k = random_keys(dtype, n, seed=n % 97, mode="bits") if n else np.zeros(0, npdt)
        d = gpu.to_dev(k) if n else None
        bits = k.view({1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[w])
        for desc in (False, True):
            counts = np.zeros(256, dtype=np.uint64)
            cb._capi.check(cb.lib().bcb_radix_top_histogram(cb.command_queue().handle, dtype_code(npdt), int(not desc),
                                                            d.data_ptr() if n else None, n, counts.ctypes.data))
            tk = cbd.transformed_keys(bits, dtype_code(npdt), not desc)
            exp = np.bincount((tk >> np.uint64(8 * w - 8)).astype(np.int64), minlength=256)
            np.testing.assert_array_equal(counts.astype(np.int64), exp, err_msg=f"{dtype} n={n} desc={desc}")
