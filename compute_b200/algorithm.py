"""Host-side mirror of the reference's algorithm dispatch for the sort / scan / reduce path.

Each function keeps the name, argument meaning and edge-case behaviour of its Boost.Compute counterpart
(cited per function; paths relative to include/boost/compute/) and forwards to the C ABI of
include/compute_b200.h.  Ranges are 1-D contiguous CUDA tensors (``first``..``last`` = the tensor, a slice
of a tensor is a sub-range); the queue argument plays the role of ``command_queue &queue``.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from ._capi import check, lib
from .core import NP_OF_CODE, TORCH_OF_CODE, command_queue, default_queue, dtype_code, op_code


def _q(queue):
    return queue if queue is not None else default_queue()


def _range(t: torch.Tensor, what="range"):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{what} must be a CUDA tensor (is_device_iterator static assert in the reference)")
    if t.dim() != 1 and t.numel() > 0 and what == "range":
        raise ValueError("ranges are 1-D")
    if not t.is_contiguous():
        raise ValueError(f"{what} must be contiguous")
    return t


def _values(values: torch.Tensor, n: int):
    _range(values, "values")
    if values.shape[0] != n:
        raise ValueError("values range is shorter than the key range")
    vb = values.element_size() * (values.numel() // n if n else 1)
    return values, vb


def _host_scalar(value, code: int):
    return np.array([value]).astype(NP_OF_CODE[code])


# ---------------------------------------------------------------------------------------------------------
# sort family
# ---------------------------------------------------------------------------------------------------------
def radix_sort(keys: torch.Tensor, ascending: bool = True, queue: command_queue | None = None) -> None:
    """detail::radix_sort(first, last[, ascending], queue) -- algorithm/detail/radix_sort.hpp:428-452."""
    _range(keys)
    check(lib().bcb_radix_sort(_q(queue).handle, dtype_code(keys.dtype), int(ascending), keys.data_ptr(),
                               keys.numel(), None, 0))


def radix_sort_by_key(keys: torch.Tensor, values: torch.Tensor, ascending: bool = True,
                      queue: command_queue | None = None) -> None:
    """detail::radix_sort_by_key -- algorithm/detail/radix_sort.hpp:436-461 (stable, any payload size)."""
    _range(keys)
    n = keys.numel()
    values, vb = _values(values, n)
    check(lib().bcb_radix_sort(_q(queue).handle, dtype_code(keys.dtype), int(ascending), keys.data_ptr(), n,
                               values.data_ptr(), vb))


def insertion_sort(keys: torch.Tensor, values: torch.Tensor | None = None, descending: bool = False,
                   queue: command_queue | None = None) -> None:
    """detail::serial_insertion_sort(_by_key) -- algorithm/detail/insertion_sort.hpp:25-159."""
    _range(keys)
    n = keys.numel()
    vptr, vb = None, 0
    if values is not None:
        values, vb = _values(values, n)
        vptr = values.data_ptr()
    check(lib().bcb_insertion_sort(_q(queue).handle, dtype_code(keys.dtype), int(descending), keys.data_ptr(), n, vptr, vb))


def sort(keys: torch.Tensor, descending: bool = False, queue: command_queue | None = None) -> None:
    """sort(first, last, less<T>() | greater<T>(), queue) -- algorithm/sort.hpp:182-202 via dispatch_gpu_sort
    :34-81: n < 2 nothing, n <= 32 insertion sort, else radix sort."""
    _range(keys)
    n = keys.numel()
    if n < 2:
        return
    if n <= 32:
        insertion_sort(keys, None, descending, queue)
    else:
        radix_sort(keys, not descending, queue)


def sort_host(host_keys: np.ndarray, descending: bool = False, queue: command_queue | None = None) -> None:
    """sort(host_first, host_last, queue) -- algorithm/sort.hpp:125-148 (maps the host range, sorts, copies back)."""
    if not isinstance(host_keys, np.ndarray) or not host_keys.flags.c_contiguous:
        raise TypeError("sort_host needs a contiguous numpy array")
    check(lib().bcb_sort_host(_q(queue).handle, dtype_code(host_keys.dtype), int(descending),
                              host_keys.ctypes.data, host_keys.size))


def sort_by_key(keys: torch.Tensor, values: torch.Tensor, descending: bool = False,
                queue: command_queue | None = None) -> None:
    """sort_by_key -- algorithm/sort_by_key.hpp:135-163 via dispatch_gpu_sort_by_key :33-86:
    n < 32 insertion sort by key, else radix_sort_by_key."""
    _range(keys)
    n = keys.numel()
    if n < 32:
        insertion_sort(keys, values, descending, queue)
    else:
        radix_sort_by_key(keys, values, not descending, queue)


def stable_sort(keys: torch.Tensor, descending: bool = False, queue: command_queue | None = None) -> None:
    """stable_sort with less / greater -- algorithm/stable_sort.hpp:52-71: straight to radix sort."""
    radix_sort(keys, not descending, queue)


def stable_sort_by_key(keys: torch.Tensor, values: torch.Tensor, descending: bool = False,
                       queue: command_queue | None = None) -> None:
    """stable_sort_by_key with less / greater -- algorithm/stable_sort_by_key.hpp:29-80: radix_sort_by_key."""
    radix_sort_by_key(keys, values, not descending, queue)


def is_sorted(keys: torch.Tensor, descending: bool = False, queue: command_queue | None = None) -> bool:
    """is_sorted(first, last[, greater<T>()], queue) -- algorithm/is_sorted.hpp:39-68."""
    _range(keys)
    res = ctypes.c_int(1)
    check(lib().bcb_is_sorted(_q(queue).handle, dtype_code(keys.dtype), int(descending), keys.data_ptr(), keys.numel(),
                              ctypes.byref(res)))
    return bool(res.value)


# ---------------------------------------------------------------------------------------------------------
# scan family
# ---------------------------------------------------------------------------------------------------------
def _scan(first: torch.Tensor, result: torch.Tensor, exclusive: bool, init, op, queue):
    _range(first)
    _range(result)
    n = first.numel()
    if result.numel() < n:
        raise ValueError("result range is shorter than the input range")
    out_code = dtype_code(result.dtype)
    init_arr = _host_scalar(init, out_code)
    check(lib().bcb_scan(_q(queue).handle, dtype_code(first.dtype), out_code, op_code(op), int(exclusive),
                         first.data_ptr(), result.data_ptr(), n, init_arr.ctypes.data))
    return result[n:] if result.numel() > n else result[n:n]  # "result + n" (scan_on_gpu.hpp:266)


def exclusive_scan(first: torch.Tensor, result: torch.Tensor, init=0, op="plus", queue: command_queue | None = None):
    """exclusive_scan(first, last, result[, init[, binary_op]], queue) -- algorithm/exclusive_scan.hpp:55-104.
    out[i] = init op x0 op ... op x(i-1), arithmetic in result's value type; first may alias result."""
    return _scan(first, result, True, init, op, queue)


def inclusive_scan(first: torch.Tensor, result: torch.Tensor, op="plus", queue: command_queue | None = None):
    """inclusive_scan(first, last, result[, binary_op], queue) -- algorithm/inclusive_scan.hpp:53-87."""
    return _scan(first, result, False, 0, op, queue)


def partial_sum(first: torch.Tensor, result: torch.Tensor, queue: command_queue | None = None):
    """partial_sum == inclusive_scan with plus -- algorithm/partial_sum.hpp:31-41."""
    return inclusive_scan(first, result, "plus", queue)


# ---------------------------------------------------------------------------------------------------------
# reduce / accumulate
# ---------------------------------------------------------------------------------------------------------
def reduce(first: torch.Tensor, result=None, op="plus", result_dtype=None, queue: command_queue | None = None):
    """reduce(first, last, result[, function], queue) -- algorithm/reduce.hpp:275-305.

    ``result`` may be a CUDA tensor (device iterator: the value is written to result[0], enqueue-and-return)
    or None / a 1-element numpy array (host pointer: blocks and returns the value).  ``result_dtype`` is the
    functor's type U for ``plus<U>`` over a T range (defaults to the input type).  An empty range leaves the
    result untouched (reduce.hpp:283-285) and returns None for the host form.
    """
    _range(first)
    n = first.numel()
    in_code = dtype_code(first.dtype)
    if isinstance(result, torch.Tensor):
        _range(result, "result")
        check(lib().bcb_reduce(_q(queue).handle, in_code, dtype_code(result.dtype), op_code(op), first.data_ptr(), n,
                               result.data_ptr(), 1))
        return None
    res_code = dtype_code(result_dtype) if result_dtype is not None else (
        dtype_code(result.dtype) if isinstance(result, np.ndarray) else in_code)
    host = result if isinstance(result, np.ndarray) else np.zeros(1, dtype=NP_OF_CODE[res_code])
    check(lib().bcb_reduce(_q(queue).handle, in_code, res_code, op_code(op), first.data_ptr(), n, host.ctypes.data, 0))
    if n == 0 and not isinstance(result, np.ndarray):
        return None
    return host[0]


def accumulate(first: torch.Tensor, init, op="plus", op_dtype=None, acc_dtype=None, queue: command_queue | None = None):
    """accumulate(first, last, init[, function], queue) -- algorithm/accumulate.hpp:165-188.

    Returns ``init op x0 op x1 ...`` as a host value of init's type T (``acc_dtype``; default: init's numpy
    dtype, or the input type for plain Python numbers).  ``op_dtype`` is the functor's type (default: the input
    value type, accumulate.hpp:185-187)."""
    _range(first)
    in_code = dtype_code(first.dtype)
    if acc_dtype is None:
        acc_code = dtype_code(init.dtype) if isinstance(init, np.generic) else in_code
    else:
        acc_code = dtype_code(acc_dtype)
    opd_code = dtype_code(op_dtype) if op_dtype is not None else in_code
    init_arr = _host_scalar(init, acc_code)
    out = np.zeros(1, dtype=NP_OF_CODE[acc_code])
    check(lib().bcb_accumulate(_q(queue).handle, in_code, opd_code, acc_code, op_code(op), first.data_ptr(),
                               first.numel(), init_arr.ctypes.data, out.ctypes.data))
    return out[0]


# ---------------------------------------------------------------------------------------------------------
# callers of scan / reduce: closed functor set, see include/compute_b200.h
# ---------------------------------------------------------------------------------------------------------
ARITH_NAMES = ["none", "mul", "mod", "add", "sub", "and"]
CMP_NAMES = ["eq", "ne", "lt", "le", "gt", "ge", "true"]
UNARY_NAMES = ["identity", "negate", "abs", "square"]


class _BcbPred(ctypes.Structure):
    _fields_ = [("arith", ctypes.c_int), ("cmp", ctypes.c_int), ("a_bits", ctypes.c_ulonglong), ("b_bits", ctypes.c_ulonglong)]


def predicate(cmp: str, b=0, arith: str = "none", a=0):
    """((x ARITH a) CMP b): what the reference's lambda placeholders `_1 < 5`, `_1 * 2 >= 10`, `_1 % 2 == 1` denote."""
    return (ARITH_NAMES.index(arith), a, CMP_NAMES.index(cmp), b)


def _pred_struct(pred, code: int) -> _BcbPred:
    arith, a, cmp_, b = pred
    bits = []
    for v in (a, b):
        raw = np.zeros(8, dtype=np.uint8)
        arr = _host_scalar(v, code)
        raw[: arr.itemsize] = arr.view(np.uint8)
        bits.append(int(raw.view(np.uint64)[0]))
    return _BcbPred(arith, cmp_, bits[0], bits[1])


def transform_if(first: torch.Tensor, result: torch.Tensor, function, pred, queue: command_queue | None = None) -> int:
    """transform_if(first, last, result, function, predicate, queue) -- algorithm/transform_if.hpp:42-117.
    Returns the number of elements written (the reference returns result + count)."""
    _range(first)
    _range(result, "result")
    code = dtype_code(first.dtype)
    if dtype_code(result.dtype) != code:
        raise TypeError("transform_if: result and input value types differ")
    ps = _pred_struct(pred, code)
    count = ctypes.c_size_t()
    check(lib().bcb_transform_if(_q(queue).handle, code, first.data_ptr(), first.numel(), UNARY_NAMES.index(function),
                                 ctypes.byref(ps), result.data_ptr(), ctypes.byref(count)))
    return int(count.value)


def copy_if(first: torch.Tensor, result: torch.Tensor, pred, queue: command_queue | None = None) -> int:
    """copy_if(first, last, result, predicate, queue) -- algorithm/copy_if.hpp:28-52 (transform_if with identity)."""
    return transform_if(first, result, "identity", pred, queue)


def count_if(first: torch.Tensor, pred, queue: command_queue | None = None) -> int:
    """count_if(first, last, predicate, queue) -- algorithm/count_if.hpp:31-58, detail/count_if_with_reduce.hpp:27-80."""
    _range(first)
    code = dtype_code(first.dtype)
    ps = _pred_struct(pred, code)
    count = ctypes.c_ulonglong()
    check(lib().bcb_count_if(_q(queue).handle, code, first.data_ptr(), first.numel(), ctypes.byref(ps), ctypes.byref(count)))
    return int(count.value)


def count(first: torch.Tensor, value, queue: command_queue | None = None) -> int:
    """count(first, last, value, queue) -- algorithm/count.hpp:32-59: count_if(_1 == value)."""
    return count_if(first, predicate("eq", value), queue)


def transform_reduce(first: torch.Tensor, transform, reduce_op="plus", first2: torch.Tensor | None = None, result=None,
                     queue: command_queue | None = None):
    """transform_reduce(first, last, result, transform, reduce, queue) -- algorithm/transform_reduce.hpp:40-90; with
    ``first2`` the binary form (first1, last1, first2, result, transform, reduce).  ``transform`` is a unary name
    (identity / negate / abs / square) or, for the binary form, an operator name.  ``result``: CUDA tensor (device
    iterator) or None (host value returned; None for an empty range)."""
    _range(first)
    code = dtype_code(first.dtype)
    n = first.numel()
    p2 = None
    if first2 is not None:
        _range(first2)
        if dtype_code(first2.dtype) != code or first2.numel() < n:
            raise ValueError("transform_reduce: second range must have the same value type and at least n elements")
        p2 = first2.data_ptr()
        t = op_code(transform)
    else:
        t = UNARY_NAMES.index(transform)
    if isinstance(result, torch.Tensor):
        check(lib().bcb_transform_reduce(_q(queue).handle, code, first.data_ptr(), p2, n, t, op_code(reduce_op), result.data_ptr(), 1))
        return None
    host = np.zeros(1, dtype=NP_OF_CODE[code])
    check(lib().bcb_transform_reduce(_q(queue).handle, code, first.data_ptr(), p2, n, t, op_code(reduce_op), host.ctypes.data, 0))
    return None if n == 0 else host[0]


def inner_product(first1: torch.Tensor, first2: torch.Tensor, init, queue: command_queue | None = None):
    """inner_product(first1, last1, first2, init, queue) -- algorithm/inner_product.hpp:40-64:
    accumulate(transform(multiplies)(zip(first1, first2)), init, plus).  Returns init's type = the value type."""
    code = dtype_code(first1.dtype)
    r = transform_reduce(first1, "multiplies", "plus", first2, None, queue)
    init_v = _host_scalar(init, code)[0]
    if r is None:
        return init_v
    with np.errstate(over="ignore"):
        return NP_OF_CODE[code].type(init_v + r)


def reduce_by_key(keys: torch.Tensor, values: torch.Tensor, keys_result: torch.Tensor, values_result: torch.Tensor,
                  op="plus", queue: command_queue | None = None) -> int:
    """reduce_by_key(keys_first, keys_last, values_first, keys_result, values_result[, function], queue) --
    algorithm/reduce_by_key.hpp:60-118.  Returns the number of (key, reduced value) pairs written."""
    _range(keys)
    _range(values, "values")
    _range(keys_result, "keys_result")
    _range(values_result, "values_result")
    n = keys.numel()
    if values.numel() < n:
        raise ValueError("values range is shorter than the key range")
    count = ctypes.c_size_t()
    check(lib().bcb_reduce_by_key(_q(queue).handle, dtype_code(keys.dtype), dtype_code(values.dtype), keys.data_ptr(), values.data_ptr(), n,
                                  keys_result.data_ptr(), values_result.data_ptr(), op_code(op), ctypes.byref(count)))
    return int(count.value)


def transform(first: torch.Tensor, result: torch.Tensor, function, queue: command_queue | None = None) -> int:
    """transform(first, last, result, function, queue) -- algorithm/transform.hpp:30-75, unary form, closed function set."""
    return transform_if(first, result, function, predicate("true"), queue)


def equal(first1: torch.Tensor, first2: torch.Tensor, queue: command_queue | None = None) -> bool:
    """equal(first1, last1, first2, queue) -- algorithm/equal.hpp:30-47, on the bit patterns of same-typed ranges."""
    _range(first1)
    _range(first2)
    n = first1.numel()
    if dtype_code(first1.dtype) != dtype_code(first2.dtype) or first2.numel() < n:
        raise ValueError("equal: ranges must have the same value type and the second at least n elements")
    if n == 0:
        return True
    code = {1: 1, 2: 3, 4: 5, 8: 7}[first1.element_size()]  # unsigned type of the same width
    host = np.zeros(1, dtype=NP_OF_CODE[code])
    check(lib().bcb_transform_reduce(_q(queue).handle, code, first1.data_ptr(), first2.data_ptr(), n, op_code("bit_xor"), op_code("bit_or"),
                                     host.ctypes.data, 0))
    return int(host[0]) == 0


def is_permutation(first1: torch.Tensor, first2: torch.Tensor, queue: command_queue | None = None) -> bool:
    """is_permutation(first1, last1, first2, last2, queue) -- algorithm/is_permutation.hpp:43-67: sort copies, compare."""
    _range(first1)
    _range(first2)
    if first1.numel() != first2.numel():
        return False
    a, b = first1.clone(), first2.clone()
    sort(a, False, queue)
    sort(b, False, queue)
    return equal(a, b, queue)


def sort_by_transform(first: torch.Tensor, function, descending: bool = False, queue: command_queue | None = None) -> None:
    """experimental::sort_by_transform(first, last, transform, compare, queue) -- experimental/sort_by_transform.hpp:26-63:
    keys = transform(range); sort_by_key(keys, range, compare)."""
    _range(first)
    if first.numel() < 2:
        return
    keys = torch.empty_like(first)
    transform(first, keys, function, queue)
    sort_by_key(keys, first, descending, queue)


def sort_by_field(records: torch.Tensor, field_offset: int, field_dtype, unary: str = "identity", descending: bool = False,
                  queue: command_queue | None = None) -> None:
    """sort / stable_sort / detail::merge_sort_on_gpu with a comparator of the family f(a.field) < f(b.field)
    (algorithm/sort.hpp:83-106, stable_sort.hpp:34-50, detail/merge_sort_on_gpu.hpp:523-572): ``records`` is a contiguous
    2-D tensor, one record per row; the scalar field of ``field_dtype`` sits at byte ``field_offset`` of every row;
    ``unary`` is "identity" or "abs".  Stable, in place."""
    _range(records, "records")
    if records.dim() != 2:
        raise ValueError("sort_by_field needs a 2-D tensor: one record per row")
    row = records.shape[1] * records.element_size()
    check(lib().bcb_sort_by_field(_q(queue).handle, records.data_ptr(), records.shape[0], row, int(field_offset), dtype_code(field_dtype),
                                  UNARY_NAMES.index(unary), int(descending)))


def is_sorted_by_field(records: torch.Tensor, field_offset: int, field_dtype, unary: str = "identity", descending: bool = False,
                       queue: command_queue | None = None) -> bool:
    """is_sorted(first, last, compare) (is_sorted.hpp:39-68) for the same comparator family; blocks."""
    _range(records, "records")
    if records.dim() != 2:
        raise ValueError("is_sorted_by_field needs a 2-D tensor: one record per row")
    row = records.shape[1] * records.element_size()
    out = ctypes.c_int(1)
    check(lib().bcb_is_sorted_by_field(_q(queue).handle, records.data_ptr(), records.shape[0], row, int(field_offset),
                                       dtype_code(field_dtype), UNARY_NAMES.index(unary), int(descending), ctypes.byref(out)))
    return bool(out.value)


SET_OPS = ("union", "intersection", "difference", "symmetric_difference")


def _set_operation(which: str, first1: torch.Tensor, first2: torch.Tensor, result: torch.Tensor, queue) -> int:
    _range(first1)
    _range(first2, "second range")
    _range(result, "result")
    code = dtype_code(first1.dtype)
    if dtype_code(first2.dtype) != code:
        raise ValueError("set operations need two ranges of the same value type")
    # the result may be another integer type of the same width (test_set_union.cpp:24-42 writes int_ into uint_):
    # same-width integer conversion keeps the bits
    if result.element_size() != first1.element_size() or result.is_floating_point() != first1.is_floating_point():
        raise ValueError("result must have the inputs' value type (or an integer type of the same width)")
    worst = {"union": first1.numel() + first2.numel(), "intersection": min(first1.numel(), first2.numel()),
             "difference": first1.numel(), "symmetric_difference": first1.numel() + first2.numel()}[which]
    if result.numel() < worst:
        # the C ABI writes at most `worst` elements; a shorter result is fine when the caller knows the count
        # (the reference's tests size it exactly) -- checked after the fact
        tmp = torch.empty(worst, dtype=first1.dtype, device=first1.device)
    else:
        tmp = None
    count = ctypes.c_size_t()
    dst = result if tmp is None else tmp
    check(lib().bcb_set_operation(_q(queue).handle, code, SET_OPS.index(which), first1.data_ptr(), first1.numel(), first2.data_ptr(),
                                  first2.numel(), dst.data_ptr(), ctypes.byref(count)))
    n = int(count.value)
    if tmp is not None:
        if n > result.numel():
            raise ValueError(f"result holds {result.numel()} elements, the set {which} has {n}")
        result.view(torch.uint8)[: n * result.element_size()].copy_(tmp.view(torch.uint8)[: n * result.element_size()])
    return n


def set_union(first1: torch.Tensor, first2: torch.Tensor, result: torch.Tensor, queue: command_queue | None = None) -> int:
    """set_union(first1, last1, first2, last2, result, queue) -- algorithm/set_union.hpp:120-199.  Returns the count."""
    return _set_operation("union", first1, first2, result, queue)


def set_intersection(first1: torch.Tensor, first2: torch.Tensor, result: torch.Tensor, queue: command_queue | None = None) -> int:
    """set_intersection(...) -- algorithm/set_intersection.hpp:104-175."""
    return _set_operation("intersection", first1, first2, result, queue)


def set_difference(first1: torch.Tensor, first2: torch.Tensor, result: torch.Tensor, queue: command_queue | None = None) -> int:
    """set_difference(...) -- algorithm/set_difference.hpp:112-186."""
    return _set_operation("difference", first1, first2, result, queue)


def set_symmetric_difference(first1: torch.Tensor, first2: torch.Tensor, result: torch.Tensor, queue: command_queue | None = None) -> int:
    """set_symmetric_difference(...) -- algorithm/set_symmetric_difference.hpp:121-199."""
    return _set_operation("symmetric_difference", first1, first2, result, queue)


def _find_extremum(first: torch.Tensor, want_max: bool, queue) -> int:
    _range(first)
    idx = ctypes.c_size_t()
    check(lib().bcb_find_extremum(_q(queue).handle, dtype_code(first.dtype), first.data_ptr(), first.numel(), int(want_max), ctypes.byref(idx)))
    return int(idx.value)


def min_element(first: torch.Tensor, queue: command_queue | None = None) -> int:
    """min_element(first, last, queue) -- algorithm/min_element.hpp:36-80 (less<T>): index of the first smallest element
    (0 for an empty range, like the reference's `first`)."""
    return _find_extremum(first, False, queue)


def max_element(first: torch.Tensor, queue: command_queue | None = None) -> int:
    """max_element(first, last, queue) -- algorithm/max_element.hpp:36-79: index of the first largest element."""
    return _find_extremum(first, True, queue)


def minmax_element(first: torch.Tensor, queue: command_queue | None = None):
    """minmax_element(first, last, queue) -- algorithm/minmax_element.hpp:33-66: (min_element, max_element)."""
    return min_element(first, queue), max_element(first, queue)


__all__ = [
    "radix_sort", "radix_sort_by_key", "insertion_sort", "sort", "sort_host", "sort_by_field", "is_sorted_by_field", "sort_by_key", "stable_sort",
    "stable_sort_by_key", "is_sorted", "exclusive_scan", "inclusive_scan", "partial_sum", "reduce", "accumulate",
    "predicate", "transform_if", "copy_if", "count_if", "count", "transform_reduce", "inner_product", "reduce_by_key",
    "transform", "equal", "is_permutation", "sort_by_transform", "set_union", "set_intersection", "set_difference",
    "set_symmetric_difference", "min_element", "max_element", "minmax_element",
]
_ = TORCH_OF_CODE
