"""compute_b200 -- B200-native sort / scan / reduce behind Boost.Compute's algorithm interface.

Python mirror of the reference's host-side dispatch (algorithm/sort.hpp, sort_by_key.hpp, stable_sort*.hpp,
exclusive_scan.hpp, inclusive_scan.hpp, reduce.hpp, accumulate.hpp) on top of the C ABI declared in
include/compute_b200.h.  Device memory and streams come from PyTorch (plumbing only); every algorithm call
lands in hand-written sm_100a CUDA kernels (compute_b200/csrc).  No CPU fallback.
"""
from ._capi import ComputeError, LIB_PATH, lib  # noqa: F401
from .algorithm import (  # noqa: F401
    accumulate,
    copy_if,
    max_element,
    min_element,
    minmax_element,
    set_difference,
    set_intersection,
    set_symmetric_difference,
    set_union,
    count,
    equal,
    is_permutation,
    sort_by_transform,
    transform,
    count_if,
    inner_product,
    predicate,
    reduce_by_key,
    transform_if,
    transform_reduce,
    exclusive_scan,
    inclusive_scan,
    insertion_sort,
    is_sorted,
    partial_sum,
    radix_sort,
    radix_sort_by_key,
    reduce,
    sort,
    sort_by_key,
    sort_host,
    sort_by_field,
    is_sorted_by_field,
    stable_sort,
    stable_sort_by_key,
)
from .core import DTYPE_NAMES, command_queue, dtype_code, op_code  # noqa: F401
