"""ctypes front-end of the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, ``__graft_entry__.smoke()`` and
bench.py's ``cpu_baseline`` / ``--impl reference`` legs.  The product package
``compute_b200`` never imports this module.

Parity pinning: pinned against the reference's own golden vectors
(tests/golden/reference_vectors.json); the reference cannot be compiled in this
image (no Boost, no OpenCL), see oracle/Makefile.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

DTYPES = ["char", "uchar", "short", "ushort", "int", "uint", "long", "ulong", "float", "double"]
NP_DTYPES = {
    "char": np.int8, "uchar": np.uint8, "short": np.int16, "ushort": np.uint16,
    "int": np.int32, "uint": np.uint32, "long": np.int64, "ulong": np.uint64,
    "float": np.float32, "double": np.float64,
}
OPS = ["plus", "multiplies", "min", "max", "bit_and", "bit_or", "bit_xor", "minus", "divides"]


def dtype_code(name_or_np) -> int:
    if isinstance(name_or_np, str):
        return DTYPES.index(name_or_np)
    dt = np.dtype(name_or_np)
    for i, n in enumerate(DTYPES):
        if np.dtype(NP_DTYPES[n]) == dt:
            return i
    raise ValueError(f"unsupported dtype {dt}")


def op_code(name: str) -> int:
    return OPS.index(name)


_STL_LIB_PATH = os.path.join(_HERE, "liboracle_stl.so")


def build(force: bool = False) -> str:
    stale = False
    for lib_path, src_name in ((_LIB_PATH, "oracle.c"), (_STL_LIB_PATH, "stl_baseline.cpp")):
        src = os.path.join(_HERE, src_name)
        stale |= not os.path.exists(lib_path) or os.path.getmtime(lib_path) < os.path.getmtime(src)
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _LIB_PATH


_stl = None


def stl_lib():
    """Single-threaded STL baselines (perf/perf_stl_sort.cpp, perf_stl_partial_sum.cpp, perf_stl_accumulate.cpp)."""
    global _stl
    if _stl is None:
        build()
        L = ctypes.CDLL(_STL_LIB_PATH)
        vp, sz = ctypes.c_void_p, ctypes.c_size_t
        L.orc_stl_sort_u32.restype = ctypes.c_double
        L.orc_stl_sort_u32.argtypes = [vp, sz]
        L.orc_stl_partial_sum_i32.restype = ctypes.c_double
        L.orc_stl_partial_sum_i32.argtypes = [vp, vp, sz]
        L.orc_stl_accumulate_i32.restype = ctypes.c_double
        L.orc_stl_accumulate_i32.argtypes = [vp, sz, vp]
        _stl = L
    return _stl


def stl_sort_u32(keys: np.ndarray) -> float:
    """std::sort in place; returns seconds."""
    assert keys.dtype == np.uint32 and keys.flags.c_contiguous
    return float(stl_lib().orc_stl_sort_u32(_ptr(keys), keys.size))


def stl_partial_sum_i32(x: np.ndarray, out: np.ndarray) -> float:
    assert x.dtype == np.int32 and out.dtype == np.int32
    return float(stl_lib().orc_stl_partial_sum_i32(_ptr(x), _ptr(out), x.size))


def stl_accumulate_i32(x: np.ndarray):
    assert x.dtype == np.int32
    r = ctypes.c_int32()
    t = float(stl_lib().orc_stl_accumulate_i32(_ptr(x), x.size, ctypes.byref(r)))
    return t, int(r.value)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, sz, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
        L.orc_radix_key.restype = ctypes.c_uint64
        L.orc_radix_key.argtypes = [i32, i32, ctypes.c_uint64]
        L.orc_radix_sort.argtypes = [i32, i32, vp, sz, vp, sz]
        L.orc_insertion_sort.argtypes = [i32, i32, vp, sz, vp, sz]
        L.orc_sort.argtypes = [i32, i32, vp, sz]
        L.orc_sort_by_key.argtypes = [i32, i32, vp, sz, vp, sz]
        L.orc_stable_sort.argtypes = [i32, i32, vp, sz]
        L.orc_stable_sort_by_key.argtypes = [i32, i32, vp, sz, vp, sz]
        L.orc_scan.argtypes = [i32, i32, i32, i32, vp, vp, sz, vp]
        L.orc_reduce.argtypes = [i32, i32, i32, vp, sz, vp]
        L.orc_accumulate.argtypes = [i32, i32, i32, i32, vp, sz, vp, vp]
        L.orc_sum_f64.argtypes = [i32, vp, sz, vp, vp]
        L.orc_prefix_f64.argtypes = [i32, vp, sz, vp, vp]
        L.orc_is_sorted.argtypes = [i32, i32, vp, sz]
        L.orc_merge_sort_on_cpu_u32.argtypes = [vp, sz, i32]
        L.orc_scan_on_cpu_i32.argtypes = [vp, vp, sz, i32, ctypes.c_int32, i32]
        L.orc_reduce_on_cpu_i32.argtypes = [vp, sz, vp, i32]
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return ctypes.c_void_p(a.ctypes.data)


def _check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"oracle {what} failed with code {rc}")


def radix_key(dtype: str, ascending: bool, bits: int) -> int:
    return int(lib().orc_radix_key(dtype_code(dtype), int(ascending), ctypes.c_uint64(bits)))


def _sort_call(fn_name, keys, descending, values=None, flag_is_ascending=False):
    keys = np.ascontiguousarray(keys).copy()
    code = dtype_code(keys.dtype)
    flag = int(not descending) if flag_is_ascending else int(descending)
    fn = getattr(lib(), fn_name)
    if values is None:
        if fn_name in ("orc_sort", "orc_stable_sort"):
            _check(fn(code, flag, _ptr(keys), keys.size), fn_name)
        else:
            _check(fn(code, flag, _ptr(keys), keys.size, None, 0), fn_name)
        return keys
    values = np.ascontiguousarray(values).copy()
    vb = values.nbytes // max(1, keys.size) if keys.size else values.dtype.itemsize
    _check(fn(code, flag, _ptr(keys), keys.size, _ptr(values), vb), fn_name)
    return keys, values


def radix_sort(keys, descending=False, values=None):
    """detail::radix_sort / radix_sort_by_key (radix_sort.hpp:428-462)."""
    return _sort_call("orc_radix_sort", keys, descending, values, flag_is_ascending=True)


def insertion_sort(keys, descending=False, values=None):
    return _sort_call("orc_insertion_sort", keys, descending, values)


def sort(keys, descending=False):
    return _sort_call("orc_sort", keys, descending)


def sort_by_key(keys, values, descending=False):
    return _sort_call("orc_sort_by_key", keys, descending, values)


def stable_sort(keys, descending=False):
    return _sort_call("orc_stable_sort", keys, descending)


def stable_sort_by_key(keys, values, descending=False):
    return _sort_call("orc_stable_sort_by_key", keys, descending, values)


def scan(x, op="plus", exclusive=False, init=None, out_dtype=None):
    x = np.ascontiguousarray(x)
    out_np = np.dtype(out_dtype) if out_dtype is not None else x.dtype
    out = np.empty(x.shape, dtype=out_np)
    init_arr = np.zeros(1, dtype=out_np) if init is None else np.array([init]).astype(out_np)
    _check(lib().orc_scan(dtype_code(x.dtype), dtype_code(out_np), op_code(op), int(exclusive),
                          _ptr(x), _ptr(out), x.size, _ptr(init_arr)), "scan")
    return out


def reduce(x, op="plus", result_dtype=None):
    x = np.ascontiguousarray(x)
    res_np = np.dtype(result_dtype) if result_dtype is not None else x.dtype
    out = np.zeros(1, dtype=res_np)
    _check(lib().orc_reduce(dtype_code(x.dtype), dtype_code(res_np), op_code(op), _ptr(x), x.size, _ptr(out)), "reduce")
    return out[0]


def accumulate(x, init, op="plus", op_dtype=None, acc_dtype=None):
    """init's dtype (acc_dtype) is the return type; op_dtype is the functor's type (defaults to x.dtype)."""
    x = np.ascontiguousarray(x)
    acc_np = np.dtype(acc_dtype) if acc_dtype is not None else np.asarray(init).dtype
    op_np = np.dtype(op_dtype) if op_dtype is not None else x.dtype
    init_arr = np.array([init]).astype(acc_np)
    out = np.zeros(1, dtype=acc_np)
    _check(lib().orc_accumulate(dtype_code(x.dtype), dtype_code(op_np), dtype_code(acc_np), op_code(op),
                                _ptr(x), x.size, _ptr(init_arr), _ptr(out)), "accumulate")
    return out[0]


def sum_f64(x):
    x = np.ascontiguousarray(x)
    s = ctypes.c_double()
    a = ctypes.c_double()
    _check(lib().orc_sum_f64(dtype_code(x.dtype), _ptr(x), x.size, ctypes.byref(s), ctypes.byref(a)), "sum_f64")
    return s.value, a.value


def prefix_f64(x):
    x = np.ascontiguousarray(x)
    p = np.empty(x.size, dtype=np.float64)
    a = np.empty(x.size, dtype=np.float64)
    _check(lib().orc_prefix_f64(dtype_code(x.dtype), _ptr(x), x.size, _ptr(p), _ptr(a)), "prefix_f64")
    return p, a


def is_sorted(keys, descending=False) -> bool:
    keys = np.ascontiguousarray(keys)
    return bool(lib().orc_is_sorted(dtype_code(keys.dtype), int(descending), _ptr(keys), keys.size))


def merge_sort_on_cpu_u32(keys: np.ndarray, threads: int) -> None:
    """In-place; keys must be a contiguous uint32 array (CPU-device sort path, sort.hpp:117-121)."""
    assert keys.dtype == np.uint32 and keys.flags.c_contiguous
    _check(lib().orc_merge_sort_on_cpu_u32(_ptr(keys), keys.size, threads), "merge_sort_on_cpu")


def scan_on_cpu_i32(x: np.ndarray, out: np.ndarray, exclusive: bool, init: int, threads: int) -> None:
    assert x.dtype == np.int32 and out.dtype == np.int32
    _check(lib().orc_scan_on_cpu_i32(_ptr(x), _ptr(out), x.size, int(exclusive), init, threads), "scan_on_cpu")


def reduce_on_cpu_i32(x: np.ndarray, threads: int) -> int:
    assert x.dtype == np.int32
    out = np.zeros(1, dtype=np.int32)
    _check(lib().orc_reduce_on_cpu_i32(_ptr(x), x.size, _ptr(out), threads), "reduce_on_cpu")
    return int(out[0])


# ---------------------------------------------------------------------------------------------------------------
# callers of scan / reduce: numpy restatements of the SERIAL definitions the
# reference's algorithms implement (test infrastructure, like everything in this package)
# ---------------------------------------------------------------------------------------------------------------
ARITH_NAMES = ["none", "mul", "mod", "add", "sub", "and"]
CMP_NAMES = ["eq", "ne", "lt", "le", "gt", "ge", "true"]


def _promote(x: np.ndarray) -> np.ndarray:
    """OpenCL C integer promotion: operands narrower than int are computed in int."""
    if x.dtype.kind in "iu" and x.dtype.itemsize < 4:
        return x.astype(np.int32)
    return x


def eval_predicate(x: np.ndarray, pred) -> np.ndarray:
    """((x ARITH a) CMP b) -- the expressions the reference's tests build from lambda placeholders
    (test_copy_if.cpp:35-66, test_transform_if.cpp:34, test_count.cpp:69)."""
    arith, a, cmp_, b = pred
    arith = ARITH_NAMES[arith] if isinstance(arith, int) else arith
    cmp_ = CMP_NAMES[cmp_] if isinstance(cmp_, int) else cmp_
    xv = _promote(x)
    av = _promote(np.array([a]).astype(x.dtype))[0]
    bv = _promote(np.array([b]).astype(x.dtype))[0]
    with np.errstate(over="ignore", divide="ignore", invalid="ignore"):
        if arith == "mul":
            y = xv * av
        elif arith == "add":
            y = xv + av
        elif arith == "sub":
            y = xv - av
        elif arith == "mod":  # C remainder: sign of the dividend
            y = np.where(av == 0, 0, np.fmod(xv, av if av != 0 else 1)).astype(xv.dtype)
        elif arith == "and":
            y = xv & av
        else:
            y = xv
    y = y.astype(xv.dtype)
    return {"eq": y == bv, "ne": y != bv, "lt": y < bv, "le": y <= bv, "gt": y > bv, "ge": y >= bv,
            "true": np.ones(x.shape, bool)}[cmp_]


def apply_unary(x: np.ndarray, name: str) -> np.ndarray:
    with np.errstate(over="ignore"):
        if name == "negate":
            return (-x).astype(x.dtype) if x.dtype.kind != "u" else (np.zeros_like(x) - x)
        if name == "abs":
            return np.abs(x).astype(x.dtype)
        if name == "square":
            return (x * x).astype(x.dtype)
    return x.copy()


def transform_if(x: np.ndarray, function: str, pred) -> np.ndarray:
    """transform_if.hpp:42-117 semantics: function(x_i) for every i with predicate(x_i), in input order."""
    x = np.ascontiguousarray(x)
    return apply_unary(x[eval_predicate(x, pred)], function)


def copy_if(x: np.ndarray, pred) -> np.ndarray:
    return transform_if(x, "identity", pred)


def count_if(x: np.ndarray, pred) -> int:
    """count_if_with_reduce.hpp:27-80: sum of predicate(x_i) in ulong."""
    return int(np.count_nonzero(eval_predicate(np.ascontiguousarray(x), pred)))


def _fold(values: np.ndarray, op: str):
    """Left fold in the array's own type (wrap-around for integers), serial_reduce.hpp:45-50."""
    if values.size == 0:
        return None
    if values.dtype.kind in "iu" and op in ("plus", "multiplies"):
        w = values.astype({1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[values.dtype.itemsize])
        with np.errstate(over="ignore"):
            r = np.add.reduce(w, dtype=w.dtype) if op == "plus" else np.multiply.reduce(w, dtype=w.dtype)
        return np.array([r], dtype=w.dtype).view(values.dtype)[0]
    if op == "plus":
        return values.dtype.type(np.add.reduce(values.astype(np.float64))) if values.dtype.kind == "f" else None
    if op == "multiplies":
        return values.dtype.type(np.multiply.reduce(values.astype(np.float64)))
    if op == "min":
        return values.min()
    if op == "max":
        return values.max()
    raise ValueError(op)


def transform_reduce(x: np.ndarray, transform: str, reduce_op: str = "plus", y: np.ndarray | None = None):
    """transform_reduce.hpp:40-90: reduce(transform_iterator(first, transform) ...).  Floating-point sums are returned
    from a float64 fold (compare with a tolerance)."""
    x = np.ascontiguousarray(x)
    if y is None:
        t = apply_unary(x, transform)
    else:
        y = np.ascontiguousarray(y)[: x.size]
        with np.errstate(over="ignore"):
            t = {"plus": x + y, "minus": x - y, "multiplies": x * y, "min": np.minimum(x, y), "max": np.maximum(x, y)}[transform].astype(x.dtype)
    return _fold(t, reduce_op)


def inner_product(x: np.ndarray, y: np.ndarray, init):
    """inner_product.hpp:40-64: init + sum of x_i * y_i, in the value type."""
    r = transform_reduce(x, "multiplies", "plus", y)
    init_v = np.array([init]).astype(x.dtype)[0]
    if r is None:
        return init_v
    with np.errstate(over="ignore"):
        return x.dtype.type(init_v + r)


def reduce_by_key(keys: np.ndarray, values: np.ndarray, op: str = "plus"):
    """reduce_by_key.hpp:60-118 (serial definition, detail/serial_reduce_by_key.hpp): every run of consecutive equal
    keys -> (key, left fold of its values).  Float sums fold in float64 and are rounded once (compare with tolerance)."""
    keys = np.ascontiguousarray(keys)
    values = np.ascontiguousarray(values)[: keys.size]
    if keys.size == 0:
        return keys[:0].copy(), values[:0].copy()
    heads = np.ones(keys.size, bool)
    heads[1:] = ~(keys[1:] == keys[:-1])
    starts = np.flatnonzero(heads)
    ends = np.append(starts[1:], keys.size)
    out_v = np.empty(starts.size, dtype=values.dtype)
    for j, (s0, e0) in enumerate(zip(starts, ends)):
        out_v[j] = _fold(values[s0:e0], op)
    return keys[starts].copy(), out_v


def is_permutation(a: np.ndarray, b: np.ndarray) -> bool:
    """is_permutation.hpp:43-67: equal after sorting both (compared on the bit patterns, as the GPU path does)."""
    if a.size != b.size:
        return False
    return sort(np.ascontiguousarray(a), False).tobytes() == sort(np.ascontiguousarray(b), False).tobytes()


def sort_by_transform(x: np.ndarray, function: str, descending: bool = False) -> np.ndarray:
    """experimental/sort_by_transform.hpp:26-63: sort_by_key(transform(x), x, compare) -- stable by the transformed key."""
    x = np.ascontiguousarray(x)
    if x.size < 2:
        return x.copy()
    return sort_by_key(apply_unary(x, function), x.copy(), descending)[1]


# ---- sorts with a custom comparator: the comparator family f(a.field) < f(b.field) ----
def project_field(records: np.ndarray, offset: int, dtype: str, unary: str = "identity") -> np.ndarray:
    """The value a comparator of the family looks at: the scalar field of type `dtype` at byte `offset` of every record
    (records: uint8[n, record_bytes]), through `unary` (identity or abs).  abs of a signed integer is UNSIGNED, as
    OpenCL's abs() is (the reference's abs_sort comparator, test_merge_sort_gpu.cpp:239-242, compares those)."""
    npdt = np.dtype(NP_DTYPES[dtype])
    rec = np.ascontiguousarray(records).view(np.uint8).reshape(records.shape[0], -1)
    f = np.ascontiguousarray(rec[:, offset:offset + npdt.itemsize]).view(npdt).reshape(-1)
    if unary == "identity":
        return f
    if unary != "abs":
        raise ValueError(unary)
    if npdt.kind == "i":
        u = np.dtype("u%d" % npdt.itemsize)
        return np.where(f < 0, (0 - f.astype(u)).astype(u), f.astype(u)).astype(u)
    if npdt.kind == "f":
        return np.where(f < 0, -f, f).astype(npdt)
    return f


def sort_by_field(records: np.ndarray, offset: int, dtype: str, unary: str = "identity", descending: bool = False) -> np.ndarray:
    """stable_sort(first, last, compare) with compare(a, b) = f(a.field) < f(b.field) (">" when descending):
    stable_sort.hpp:34-50 -> detail/merge_sort_on_gpu.hpp:523-572 with stable = true, restated as the definition of a
    stable sort by that comparator (native compare of the projected values; equal elements keep their input order).
    sort() with the same comparator (sort.hpp:83-106) may return any order of equal elements; this is one of them."""
    key = project_field(records, offset, dtype, unary)
    n = key.shape[0]
    if descending:  # stable for ">": descending keys, ties in input order
        order = (n - 1 - np.argsort(key[::-1], kind="stable"))[::-1]
    else:
        order = np.argsort(key, kind="stable")
    return np.ascontiguousarray(np.ascontiguousarray(records)[order])


def is_sorted_by_field(records: np.ndarray, offset: int, dtype: str, unary: str = "identity", descending: bool = False) -> bool:
    """is_sorted(first, last, compare) (is_sorted.hpp:39-68): no adjacent pair with compare(x[i+1], x[i])."""
    key = project_field(records, offset, dtype, unary)
    if key.shape[0] < 2:
        return True
    return not bool(np.any(key[1:] > key[:-1])) if descending else not bool(np.any(key[1:] < key[:-1]))


# ---- second batch of callers: set operations on sorted ranges, extrema ----
def set_operation(which: str, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """std::set_union / set_intersection / set_difference / set_symmetric_difference restated as the two-pointer merge
    the standard (and set_union.hpp:120-199 etc. of the reference, through its serial definitions) specifies: multiset
    semantics, equal elements taken from the first range first.  Pure Python loop: small inputs only."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    out = []
    i = j = 0
    na, nb = a.size, b.size
    while i < na and j < nb:
        if a[i] < b[j]:
            if which in ("union", "difference", "symmetric_difference"):
                out.append(a[i])
            i += 1
        elif b[j] < a[i]:
            if which in ("union", "symmetric_difference"):
                out.append(b[j])
            j += 1
        else:
            if which in ("union", "intersection"):
                out.append(a[i])
            i += 1
            j += 1
    if which in ("union", "difference", "symmetric_difference"):
        out.extend(a[i:])
    if which in ("union", "symmetric_difference"):
        out.extend(b[j:])
    return np.array(out, dtype=a.dtype)


def set_operation_counting(which: str, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """The same results from the multiset-count definition (value v appears max / min / a-b / |a-b| times), vectorised
    with numpy for large inputs; cross-checked against set_operation() by the CPU tests."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    vals = np.union1d(a, b)
    ca = np.searchsorted(a, vals, "right") - np.searchsorted(a, vals, "left")
    cb = np.searchsorted(b, vals, "right") - np.searchsorted(b, vals, "left")
    if which == "union":
        c = np.maximum(ca, cb)
    elif which == "intersection":
        c = np.minimum(ca, cb)
    elif which == "difference":
        c = np.maximum(ca - cb, 0)
    else:
        c = np.abs(ca - cb)
    return np.repeat(vals, c).astype(a.dtype)


def min_element(x: np.ndarray) -> int:
    """min_element.hpp / detail/find_extrema_with_reduce.hpp:156-158: index of the FIRST smallest element (0 if empty)."""
    return int(np.argmin(x)) if x.size else 0


def max_element(x: np.ndarray) -> int:
    """max_element.hpp: index of the FIRST largest element (ties: the smaller index, find_extrema_with_reduce.hpp:156-158)."""
    return int(np.argmax(x)) if x.size else 0
