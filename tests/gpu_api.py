"""numpy-in / numpy-out adapter with the oracle module's function names, routed through the product:
compute_b200's Python mirror -> C ABI (include/compute_b200.h) -> sm_100a kernels.  Used by the -m gpu tests so
the same golden / parity checks run against the oracle and against the GPU path."""
from __future__ import annotations

import numpy as np
import torch

import compute_b200 as cb
from compute_b200.core import NP_OF_CODE, dtype_code

_TORCH = {
    np.dtype(np.int8): torch.int8, np.dtype(np.uint8): torch.uint8, np.dtype(np.int16): torch.int16,
    np.dtype(np.uint16): torch.uint16, np.dtype(np.int32): torch.int32, np.dtype(np.uint32): torch.uint32,
    np.dtype(np.int64): torch.int64, np.dtype(np.uint64): torch.uint64, np.dtype(np.float32): torch.float32,
    np.dtype(np.float64): torch.float64,
}


def to_dev(a: np.ndarray) -> torch.Tensor:
    """Upload bit-exactly (via a byte view: torch has no arithmetic on uint16/32/64, we only need storage)."""
    a = np.ascontiguousarray(a)
    if a.size == 0:
        return torch.empty(a.shape, dtype=_TORCH[a.dtype], device="cuda")
    t = torch.from_numpy(a.reshape(-1).view(np.uint8).copy()).cuda()
    return t.view(_TORCH[a.dtype]).reshape(a.shape)


def to_host(t: torch.Tensor, npdt) -> np.ndarray:
    if t.numel() == 0:
        return np.empty(tuple(t.shape), dtype=npdt)
    return t.contiguous().reshape(-1).view(torch.uint8).cpu().numpy().view(npdt).reshape(tuple(t.shape))


def _keys_values(fn, keys, values, *args):
    k = to_dev(keys)
    if values is None:
        fn(k, *args)
        torch.cuda.synchronize()
        return to_host(k, keys.dtype)
    v = to_dev(values)
    fn(k, v, *args)
    torch.cuda.synchronize()
    return to_host(k, keys.dtype), to_host(v, values.dtype)


def radix_sort(keys, descending=False, values=None):
    if values is None:
        return _keys_values(lambda k: cb.radix_sort(k, not descending), keys, None)
    return _keys_values(lambda k, v: cb.radix_sort_by_key(k, v, not descending), keys, values)


def insertion_sort(keys, descending=False, values=None):
    if values is None:
        return _keys_values(lambda k: cb.insertion_sort(k, None, descending), keys, None)
    return _keys_values(lambda k, v: cb.insertion_sort(k, v, descending), keys, values)


def sort(keys, descending=False):
    return _keys_values(lambda k: cb.sort(k, descending), keys, None)


def stable_sort(keys, descending=False):
    return _keys_values(lambda k: cb.stable_sort(k, descending), keys, None)


def sort_by_key(keys, values, descending=False):
    return _keys_values(lambda k, v: cb.sort_by_key(k, v, descending), keys, values)


def stable_sort_by_key(keys, values, descending=False):
    return _keys_values(lambda k, v: cb.stable_sort_by_key(k, v, descending), keys, values)


def sort_sub_range(fn, x, lo, hi, descending):
    """Sort [lo, hi) of a device vector in place; the rest must stay untouched (test_radix_sort.cpp:530-541)."""
    t = to_dev(x)
    sub = t[lo:hi]
    {"radix_sort": lambda: cb.radix_sort(sub, not descending), "sort": lambda: cb.sort(sub, descending),
     "stable_sort": lambda: cb.stable_sort(sub, descending), "insertion_sort": lambda: cb.insertion_sort(sub, None, descending)}[fn]()
    torch.cuda.synchronize()
    return to_host(t, x.dtype)


def sort_host(x, descending=False):
    y = np.ascontiguousarray(x).copy()
    cb.sort_host(y, descending)
    return y


def scan(x, op="plus", exclusive=False, init=None, out_dtype=None, in_place=False):
    x = np.ascontiguousarray(x)
    out_np = np.dtype(out_dtype) if out_dtype is not None else x.dtype
    d_in = to_dev(x)
    if in_place:
        assert out_np == x.dtype
        d_out = d_in
    else:
        d_out = torch.empty(x.shape, dtype=_TORCH[out_np], device="cuda")
        if x.size:
            d_out.view(torch.uint8).fill_(0xCD)
    if exclusive:
        rest = cb.exclusive_scan(d_in, d_out, 0 if init is None else init, op)
    else:
        rest = cb.inclusive_scan(d_in, d_out, op)
    assert rest.numel() == 0
    torch.cuda.synchronize()
    if not in_place and x.size:
        np.testing.assert_array_equal(to_host(d_in, x.dtype).view(np.uint8), x.view(np.uint8))  # input untouched
    return to_host(d_out, out_np)


def reduce(x, op="plus", result_dtype=None):
    x = np.ascontiguousarray(x)
    return cb.reduce(to_dev(x), None, op, result_dtype if result_dtype is not None else x.dtype)


def reduce_into(x, op, result_dtype, preset):
    """reduce into a host slot holding `preset`; an empty range must leave it untouched (test_reduce.cpp:42-49)."""
    host = np.array([preset], dtype=result_dtype)
    cb.reduce(to_dev(np.ascontiguousarray(x)), host, op)
    return host[0]


def reduce_to_device(x, op, result_dtype):
    """reduce into a device iterator (test_reduce.cpp:80-88)."""
    res = to_dev(np.zeros(2, dtype=result_dtype))
    cb.reduce(to_dev(np.ascontiguousarray(x)), res[1:], op)
    torch.cuda.synchronize()
    out = to_host(res, result_dtype)
    assert out[0] == 0
    return out[1]


def accumulate(x, init, op="plus", op_dtype=None, acc_dtype=None):
    x = np.ascontiguousarray(x)
    acc_np = np.dtype(acc_dtype) if acc_dtype is not None else np.asarray(init).dtype
    return cb.accumulate(to_dev(x), init, op, op_dtype if op_dtype is not None else x.dtype, acc_np)


def is_sorted(keys, descending=False):
    return cb.is_sorted(to_dev(np.ascontiguousarray(keys)), descending)


_ = (NP_OF_CODE, dtype_code)


# ---- callers of scan / reduce ----
def _pred(pred):
    arith, a, cmp_, b = pred
    return cb.predicate(cmp_ if isinstance(cmp_, str) else cb.algorithm.CMP_NAMES[cmp_], b,
                        arith if isinstance(arith, str) else cb.algorithm.ARITH_NAMES[arith], a)


def transform_if(x, function, pred, fill=0):
    """Returns (whole output buffer -- pre-filled with `fill`, so an overrun past the count shows --, count)."""
    x = np.ascontiguousarray(x)
    d_in = to_dev(x)
    d_out = to_dev(np.full(x.size, fill, dtype=x.dtype))
    count = cb.transform_if(d_in, d_out, function, _pred(pred))
    torch.cuda.synchronize()
    return to_host(d_out, x.dtype), count


def copy_if(x, pred, fill=0):
    return transform_if(x, "identity", pred, fill)


def count_if(x, pred):
    return cb.count_if(to_dev(np.ascontiguousarray(x)), _pred(pred))


def transform_reduce(x, transform, reduce_op="plus", y=None):
    x = np.ascontiguousarray(x)
    return cb.transform_reduce(to_dev(x), transform, reduce_op, None if y is None else to_dev(np.ascontiguousarray(y)))


def inner_product(x, y, init):
    return cb.inner_product(to_dev(np.ascontiguousarray(x)), to_dev(np.ascontiguousarray(y)), init)


def reduce_by_key(keys, values, op="plus"):
    keys, values = np.ascontiguousarray(keys), np.ascontiguousarray(values)
    dk, dv = to_dev(keys), to_dev(values)
    ok = to_dev(np.zeros(keys.size, dtype=keys.dtype))
    ov = to_dev(np.zeros(keys.size, dtype=values.dtype))
    m = cb.reduce_by_key(dk, dv, ok, ov, op)
    torch.cuda.synchronize()
    return to_host(ok, keys.dtype)[:m], to_host(ov, values.dtype)[:m]


def is_permutation(a, b):
    return cb.is_permutation(to_dev(np.ascontiguousarray(a)), to_dev(np.ascontiguousarray(b)))


def sort_by_transform(x, function, descending=False):
    d = to_dev(np.ascontiguousarray(x))
    cb.sort_by_transform(d, function, descending)
    torch.cuda.synchronize()
    return to_host(d, x.dtype)


def sort_by_field(records, offset, dtype, unary="identity", descending=False):
    rec = np.ascontiguousarray(records)
    d = to_dev(rec.view(np.uint8).reshape(rec.shape[0], -1))
    cb.sort_by_field(d, offset, dtype, unary, descending)
    torch.cuda.synchronize()
    return d.cpu().numpy().view(rec.dtype).reshape(rec.shape)


def is_sorted_by_field(records, offset, dtype, unary="identity", descending=False):
    rec = np.ascontiguousarray(records)
    d = to_dev(rec.view(np.uint8).reshape(rec.shape[0], -1))
    return cb.is_sorted_by_field(d, offset, dtype, unary, descending)


def set_operation(which, a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    d_out = to_dev(np.full((a.size + b.size + 3) * a.dtype.itemsize, 0x5A, dtype=np.uint8).view(a.dtype))
    n = getattr(cb, "set_" + which)(to_dev(a), to_dev(b), d_out)
    torch.cuda.synchronize()
    out = to_host(d_out, a.dtype)
    assert np.all(out[n:].view(np.uint8) == 0x5A), "set operation wrote past its count"
    return out[:n]


def min_element(x):
    return cb.min_element(to_dev(np.ascontiguousarray(x)))


def max_element(x):
    return cb.max_element(to_dev(np.ascontiguousarray(x)))
