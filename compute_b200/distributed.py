"""Multi-GPU sort / scan / reduce: one process per GPU, ``torch.distributed`` (NCCL over NVLink / NVSwitch) for
the exchange steps, the single-GPU CUDA kernels for all local work (SURVEY.md section 8e -- the reference itself has
no multi-device path, so this is new functionality with the single-GPU results as its parity target).

Data model: a range is block-distributed -- rank r holds the r-th contiguous block of the global range.

* ``sort`` / ``sort_by_key`` (sample sort, one exchange step): regular samples of the UNSORTED shard's transformed
  keys -> all-gather, common splitters -> bucket sizes of the shard (``bcb_partition_counts``) -> all-gather of the
  P x P count matrix -> ONE stable partition pass (``bcb_partition_scatter``) whose stores go straight into the
  destination ranks' receive buffers, which every rank has mapped through CUDA IPC (NVLink / NVSwitch peer stores, no
  all-to-all) -> stream-ordered barrier -> local stable radix sort out of the receive buffer.  The runs lie in
  source-rank order there, so equal keys keep their global input order and the concatenation of the per-rank outputs
  in rank order is bit-identical to the single-GPU sort.  Fallback plans (chosen collectively): the same partition
  into a local buffer + ``all_to_all_single`` when peer mapping is unavailable; local sort + binary-search cuts +
  all-to-all for payload sizes / rank counts the partition kernel does not cover.
* scans: local reduce -> all-gather of P partials -> carry = init op partial_0 op ... op partial_(r-1), folded in
  rank order -> local single-pass scan seeded with the carry.  Integer results are bit-exact; float results are
  deterministic (fixed fold order).
* ``reduce`` / ``accumulate``: local reduce -> all-gather of P partials -> fold in rank order.

The local primitives are reached through a small ``LocalOps`` object so that the host-side protocol (sampling,
splitters, exchange plan, carries) can be exercised on CPU with the gloo backend by the tests, which plug in the
CPU oracle there.  The product only ever uses ``CudaLocalOps`` (C ABI -> sm_100a kernels); there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch
import torch.distributed as dist

from .core import NP_OF_CODE, dtype_code, op_code

_BITS_VIEW = {1: torch.int8, 2: torch.int16, 4: torch.int32, 8: torch.int64}
_NP_UINT = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}


# ------------------------------------------------------------------------------------------------------------
# host-side protocol helpers (pure numpy; unit-tested on CPU)
# ------------------------------------------------------------------------------------------------------------
def transformed_keys(raw_bits: np.ndarray, dtype_c: int, ascending: bool) -> np.ndarray:
    """The reference's order-preserving key transform (algorithm/detail/radix_sort.hpp:100-127) on raw key bit
    patterns (unsigned ints of the key width); returns uint64 keys whose unsigned order is the sort order."""
    w = raw_bits.dtype.itemsize * 8
    x = raw_bits.astype(np.uint64)
    ones = np.uint64((1 << w) - 1)
    sign = np.uint64(1 << (w - 1))
    is_float = dtype_c in (8, 9)
    is_signed = dtype_c in (0, 2, 4, 6)
    with np.errstate(over="ignore"):
        if ascending:
            if is_float:
                mask = (((np.uint64(0) - (x >> np.uint64(w - 1))) & ones) | sign)
                return (x ^ mask) & ones
            if is_signed:
                return (x ^ sign) & ones
            return x
        neg = (np.uint64(0) - x) & ones
        if is_float:
            mask = (((np.uint64(0) - (x >> np.uint64(w - 1))) & ones) | sign)
            return (neg ^ mask) & ones
        if is_signed:
            return (neg ^ sign) & ones
        return (ones - x) & ones


def select_splitters(all_samples: np.ndarray, world: int) -> np.ndarray:
    """P-1 splitters from the gathered, transformed samples (regular sampling: every rank contributes the same
    number of evenly spaced samples of its sorted shard).  Returns uint64[world-1], non-decreasing."""
    s = np.sort(all_samples.astype(np.uint64).reshape(-1))
    if world <= 1 or s.size == 0:
        return np.empty(0, dtype=np.uint64)
    pos = (np.arange(1, world, dtype=np.int64) * s.size) // world
    return s[np.minimum(pos, s.size - 1)]


def _digit_edges(tot: np.ndarray, world: int) -> np.ndarray:
    """Bin edges e[0..world] (e[0] = 0, e[world] = 256): rank d is dealt the digit values [e[d], e[d+1]); each edge is the
    one whose cumulative global count is closest to d * N / P."""
    n = int(tot.sum())
    cum = np.cumsum(tot)                                        # cum[b - 1] = keys with digit < b
    edges = np.zeros(world + 1, dtype=np.int64)
    edges[world] = 256
    for d in range(1, world):
        target = (d * n) // world
        b = int(np.searchsorted(cum, target, side="left")) + 1  # first edge with at least `target` keys below it
        below = int(cum[b - 2]) if b >= 2 else 0                # the edge before it
        if b >= 2 and target - below < int(cum[b - 1]) - target:
            b -= 1
        edges[d] = min(max(b, edges[d - 1]), 256)
    return edges


def histogram_plan(all_hist: np.ndarray, world: int, key_bits: int, max_imbalance: float = 1.06):
    """Splitters and the P x P count matrix from the all-gathered 256-bin histograms of the keys' most significant digit
    (all_hist[src][digit]).  Rank d is dealt the digit values [b_d, b_(d+1)); the boundaries are the bin edges whose
    cumulative global count is closest to d * N / P.  Returns (splitters uint64[P-1] in transformed-key space,
    counts int64[P][P] = counts[src][dst], imbalance = largest receive count * P / N), or None when whole digit values
    cannot be dealt evenly enough (skewed keys: the caller samples instead).  Deterministic: every rank computes the same."""
    h = np.asarray(all_hist, dtype=np.int64).reshape(world, 256)
    tot = h.sum(axis=0)
    n = int(tot.sum())
    if n == 0 or world <= 1:
        return None
    edges = _digit_edges(tot, world)
    recv = np.array([int(tot[edges[d]:edges[d + 1]].sum()) for d in range(world)], dtype=np.int64)
    imbalance = float(recv.max() * world / n)
    if imbalance > max_imbalance:
        return None
    counts = np.stack([[int(h[src, edges[d]:edges[d + 1]].sum()) for d in range(world)] for src in range(world)]).astype(np.int64)
    splitters = (edges[1:world].astype(np.uint64) << np.uint64(key_bits - 8))
    return splitters, counts, imbalance


def digit_exchange_plan(all_hist: np.ndarray, world: int, max_imbalance: float = 1.06, align: int = 32):
    """Placement of every (source rank, top digit value) run when the pass over the most significant digit is the exchange
    (bcb_radix_exchange_scatter / bcb_radix_sort_segments).  Whole digit values are dealt to the ranks as in
    histogram_plan; inside the owner's receive buffer every digit value g has a segment that starts on a multiple of
    ``align`` elements, and source s's run of g lies after the runs of the sources before it (equal keys keep their global
    input order).  Returns (owner int64[256], first int64[P][256] = element index of (src, g)'s run in the owner's receive
    buffer, seg_begin int64[256], seg_len int64[256] (= global count of g), recv int64[P], span int64[P] = elements of
    receive buffer in use, imbalance) or None when no even deal exists."""
    h = np.asarray(all_hist, dtype=np.int64).reshape(world, 256)
    tot = h.sum(axis=0)
    n = int(tot.sum())
    if n == 0 or world <= 1:
        return None
    edges = _digit_edges(tot, world)
    cum = np.concatenate([[0], np.cumsum(tot)])
    recv = cum[edges[1:]] - cum[edges[:-1]]
    imbalance = float(recv.max() * world / n)
    if imbalance > max_imbalance:
        return None
    owner = np.searchsorted(edges[1:], np.arange(256), side="right").astype(np.int64)   # edges[owner] <= g < edges[owner + 1]
    owner = np.minimum(owner, world - 1)
    padded = (tot + align - 1) // align * align
    before = np.cumsum(padded) - padded                             # aligned slots of the smaller digits, globally
    seg_begin = before - before[edges[owner]]                       # ... among the owner's digits
    first = seg_begin[None, :] + (np.cumsum(h, axis=0) - h)         # + the same digit on the sources before src
    last = np.maximum(edges[1:] - 1, 0)                             # the owner's last digit value ends its span
    span = np.where(edges[1:] > edges[:-1], (seg_begin + tot)[last], 0).astype(np.int64)
    return owner, first.astype(np.int64), seg_begin.astype(np.int64), tot.astype(np.int64), recv.astype(np.int64), span, imbalance


def exchange_plan(points: np.ndarray, n_local: int):
    """Send counts per destination from the partition points of the sorted shard (points[j] = first index whose
    transformed key is >= splitter j)."""
    edges = np.concatenate([[0], np.asarray(points, dtype=np.int64), [n_local]])
    return np.diff(edges).astype(np.int64)


def fold_carry(partials: np.ndarray, rank: int, op: str, init=None):
    """init op partial_0 op ... op partial_(rank-1) in rank order, in the partials' dtype (wrap-around integers)."""
    dt = partials.dtype
    fn = _NP_OPS[op]
    acc = None if init is None else dt.type(init)
    with np.errstate(over="ignore"):
        for r in range(rank):
            acc = partials[r] if acc is None else dt.type(fn(acc, partials[r]))
    return acc


_NP_OPS = {
    "plus": lambda a, b: a + b, "multiplies": lambda a, b: a * b,
    "min": lambda a, b: b if b < a else a, "max": lambda a, b: b if a < b else a,
    "bit_and": lambda a, b: a & b, "bit_or": lambda a, b: a | b, "bit_xor": lambda a, b: a ^ b,
}


# ------------------------------------------------------------------------------------------------------------
# local primitives
# ------------------------------------------------------------------------------------------------------------
class CudaLocalOps:
    """Local work on this rank's GPU through the C ABI (the only implementation the product ships)."""

    device_type = "cuda"

    def __init__(self):
        from . import algorithm, core
        from ._capi import check, lib
        self._alg, self._check, self._lib = algorithm, check, lib()
        self.queue = core.command_queue()

    def sort(self, keys, values, descending):
        if values is None:
            self._alg.radix_sort(keys, not descending, self.queue)
        else:
            self._alg.radix_sort_by_key(keys, values, not descending, self.queue)

    def partition_points(self, sorted_keys, splitters: np.ndarray, descending: bool) -> np.ndarray:
        out = np.zeros(max(1, splitters.size), dtype=np.uint64)
        sp = np.ascontiguousarray(splitters, dtype=np.uint64)
        self._check(self._lib.bcb_partition_points(self.queue.handle, dtype_code(sorted_keys.dtype), int(not descending),
                                                   sorted_keys.data_ptr(), sorted_keys.numel(), sp.ctypes.data, sp.size,
                                                   out.ctypes.data))
        return out[: splitters.size].astype(np.int64)

    def partition(self, keys, values, splitters: np.ndarray, descending: bool):
        """Stable partition of the unsorted shard into len(splitters)+1 buckets in one pass.  Returns
        (keys_out, values_out, counts) or None when the C ABI does not support the shape (then sort-and-cut is used)."""
        nb = splitters.size + 1
        vb = 0 if values is None else values.element_size() * (values.numel() // max(1, values.shape[0]))
        if splitters.size > 7 or vb not in (0, 4, 8):
            return None  # same decision on every rank (it depends on shapes only)
        if keys.shape[0] == 0:
            return keys, values, np.zeros(nb, dtype=np.int64)
        out_k = torch.empty_like(keys)
        out_v = torch.empty_like(values) if values is not None else None
        counts = np.zeros(nb, dtype=np.uint64)
        sp = np.ascontiguousarray(splitters, dtype=np.uint64)
        rc = self._lib.bcb_partition_by_splitters(self.queue.handle, dtype_code(keys.dtype), int(not descending), keys.data_ptr(),
                                                  out_k.data_ptr(), None if values is None else values.data_ptr(),
                                                  None if values is None else out_v.data_ptr(), vb, keys.shape[0],
                                                  sp.ctypes.data, sp.size, counts.ctypes.data)
        if rc == 10002:  # BCB_EUNSUPPORTED
            return None
        self._check(rc)
        return out_k, out_v, counts.astype(np.int64)

    # ---- peer-memory exchange (NVLink stores from inside the partition pass) --------------------------------
    @staticmethod
    def _row_bytes(values) -> int:
        return 0 if values is None else values.element_size() * (values.numel() // max(1, values.shape[0]))

    def top_histogram(self, keys, descending: bool) -> np.ndarray:
        """256-bin histogram of the most significant digit of the transformed keys (int64[256]); blocks."""
        counts = np.zeros(256, dtype=np.uint64)
        self._check(self._lib.bcb_radix_top_histogram(self.queue.handle, dtype_code(keys.dtype), int(not descending), keys.data_ptr(),
                                                      keys.shape[0], counts.ctypes.data))
        return counts.astype(np.int64)

    def exchange_scatter(self, keys, values, descending: bool, dst_keys, dst_values, dst_first) -> bool:
        """One stable pass over the most significant digit: digit g's run goes to the array at dst_keys[g] (possibly a
        peer's memory) from element dst_first[g] on.  Asynchronous; the shard is not modified.  False: shape not supported
        (decided from the types: the same answer on every rank)."""
        dk = np.ascontiguousarray(dst_keys, dtype=np.uint64)  # (void *const *: 256 device addresses)
        dv = np.ascontiguousarray(dst_values, dtype=np.uint64)
        df = np.ascontiguousarray(dst_first, dtype=np.uint64)
        rc = self._lib.bcb_radix_exchange_scatter(self.queue.handle, dtype_code(keys.dtype), int(not descending), keys.data_ptr(),
                                                  None if values is None else values.data_ptr(), self._row_bytes(values), keys.shape[0],
                                                  dk.ctypes.data, dv.ctypes.data if values is not None else None, df.ctypes.data)
        if rc == 10002:  # BCB_EUNSUPPORTED
            return False
        self._check(rc)
        return True

    def sort_segments(self, recv_keys_ptr: int, recv_values_ptr: int, out_keys, out_values, descending: bool, seg_begin, seg_len) -> None:
        """Every segment of the receive buffer sorted by the remaining digits (all segments in one launch per digit); the
        last pass writes them back to back into out_keys / out_values.  Asynchronous."""
        sb = np.ascontiguousarray(seg_begin, dtype=np.uint64)
        sl = np.ascontiguousarray(seg_len, dtype=np.uint64)
        self._check(self._lib.bcb_radix_sort_segments(self.queue.handle, dtype_code(out_keys.dtype), int(not descending), recv_keys_ptr,
                                                      None if out_values is None else recv_values_ptr, self._row_bytes(out_values),
                                                      out_keys.data_ptr(), None if out_values is None else out_values.data_ptr(),
                                                      sb.ctypes.data, sl.ctypes.data, sb.size))

    def partition_counts(self, keys, splitters: np.ndarray, descending: bool) -> np.ndarray:
        counts = np.zeros(splitters.size + 1, dtype=np.uint64)
        sp = np.ascontiguousarray(splitters, dtype=np.uint64)
        self._check(self._lib.bcb_partition_counts(self.queue.handle, dtype_code(keys.dtype), int(not descending), keys.data_ptr(),
                                                   keys.shape[0], sp.ctypes.data, sp.size, counts.ctypes.data))
        return counts.astype(np.int64)

    def partition_scatter(self, keys, values, splitters: np.ndarray, descending: bool, dst_keys, dst_values) -> None:
        """Bucket b of the stable partition goes to the device address dst_keys[b] (dst_values[b]); asynchronous."""
        nb = splitters.size + 1
        sp = np.ascontiguousarray(splitters, dtype=np.uint64)
        dk = (ctypes.c_void_p * nb)(*[int(a) for a in dst_keys])
        dv = (ctypes.c_void_p * nb)(*[int(a) for a in dst_values]) if values is not None else None
        self._check(self._lib.bcb_partition_scatter(self.queue.handle, dtype_code(keys.dtype), int(not descending), keys.data_ptr(),
                                                    None if values is None else values.data_ptr(), self._row_bytes(values),
                                                    keys.shape[0], sp.ctypes.data, sp.size, dk, dv))

    def sort_copy(self, src_keys_ptr: int, out_keys, src_values_ptr, out_values, descending: bool) -> None:
        self._check(self._lib.bcb_radix_sort_copy(self.queue.handle, dtype_code(out_keys.dtype), int(not descending), src_keys_ptr,
                                                  out_keys.data_ptr(), out_keys.shape[0],
                                                  None if out_values is None else src_values_ptr,
                                                  None if out_values is None else out_values.data_ptr(), self._row_bytes(out_values)))

    def peer_alloc(self, nbytes: int):
        """(local device pointer, 64-byte IPC handle) of a fresh buffer other ranks can map."""
        ptr = ctypes.c_void_p()
        self._check(self._lib.bcb_malloc(ctypes.byref(ptr), nbytes))
        handle = np.zeros(64, dtype=np.uint8)
        self._check(self._lib.bcb_ipc_export(ptr, handle.ctypes.data))
        return int(ptr.value), handle

    def peer_open(self, handle: np.ndarray) -> int:
        ptr = ctypes.c_void_p()
        h = np.ascontiguousarray(handle, dtype=np.uint8)
        self._check(self._lib.bcb_ipc_open(h.ctypes.data, ctypes.byref(ptr)))
        return int(ptr.value)

    def peer_close(self, ptr: int) -> None:
        self._check(self._lib.bcb_ipc_close(ctypes.c_void_p(ptr)))

    def peer_free(self, ptr: int) -> None:
        self._check(self._lib.bcb_free(ctypes.c_void_p(ptr)))

    def gather_bits(self, keys, positions: np.ndarray) -> np.ndarray:
        """Raw bit patterns of keys[positions] as unsigned ints on the host."""
        w = keys.element_size()
        idx = torch.from_numpy(positions.astype(np.int64)).to(keys.device)
        picked = keys.view(_BITS_VIEW[w]).index_select(0, idx)
        return picked.cpu().numpy().view(_NP_UINT[w])

    def reduce_to(self, x, op, result_dtype):
        """Local reduction as a 1-element tensor of result_dtype on this device."""
        out = torch.empty(1, dtype=result_dtype, device=x.device)
        self._alg.reduce(x, out, op, queue=self.queue)
        return out

    def scan(self, x, out, mode: int, init, op):
        out_code = dtype_code(out.dtype)
        init_arr = np.array([0 if init is None else init]).astype(NP_OF_CODE[out_code])
        self._check(self._lib.bcb_scan(self.queue.handle, dtype_code(x.dtype), out_code, op_code(op), mode, x.data_ptr(),
                                       out.data_ptr(), x.numel(), init_arr.ctypes.data))

    def scan_with_carry(self, x, out, exclusive: bool, init, op, records, rank: int) -> None:
        """This rank's block of a block-distributed scan, seeded on the device with the partials of the ranks before it
        (``records``: the all-gathered 16-byte records, a device tensor).  Asynchronous."""
        out_code = dtype_code(out.dtype)
        init_arr = np.array([0 if init is None else init]).astype(NP_OF_CODE[out_code])
        self._check(self._lib.bcb_scan_with_carry(self.queue.handle, dtype_code(x.dtype), out_code, op_code(op), int(exclusive),
                                                  x.data_ptr(), out.data_ptr(), x.numel(), init_arr.ctypes.data, records.data_ptr(), rank))

    def empty(self, n, like):
        return torch.empty((n,) + tuple(like.shape[1:]), dtype=like.dtype, device=like.device)


# ------------------------------------------------------------------------------------------------------------
# the distributed algorithms
# ------------------------------------------------------------------------------------------------------------
class PeerExchange:
    """One receive buffer per rank, mapped into every other rank's address space (CUDA IPC over NVLink / NVSwitch),
    so that the partition pass of the multi-GPU sort can store each key directly where it belongs on its destination
    GPU.  All methods are collective and take the same arguments on every rank."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.capacity = 0
        self.local = 0          # this rank's buffer
        self.peers = []         # device address of rank r's buffer in THIS process (peers[rank] == local)
        self.failed = False     # peer mapping is not possible here: callers use the NCCL all-to-all plan
        self._live = False      # an allocation round took place (same value on every rank)

    def release(self):
        ops, ctx = self.ctx.ops, self.ctx
        if not self._live:
            return
        for r, p in enumerate(self.peers):
            if r != ctx.rank and p:
                ops.peer_close(p)
        self.peers = []
        ctx._all_gather_np(np.zeros(1, np.int32))  # every rank has unmapped the buffers before any is freed
        if self.local:
            ops.peer_free(self.local)
        self.local, self.capacity, self._live = 0, 0, False

    def ensure(self, nbytes: int) -> bool:
        if self.failed:
            return False
        if nbytes <= self.capacity:
            return True
        ops, ctx = self.ctx.ops, self.ctx
        self.release()
        cap = (int(nbytes * 1.125) + (2 << 20)) & ~((2 << 20) - 1)
        ok, handle = 1, np.zeros(64, dtype=np.uint8)
        self._live = True
        try:
            self.local, handle = ops.peer_alloc(cap)
        except Exception:  # noqa: BLE001 -- any failure here only selects the NCCL plan
            ok = 0
        handles = ctx._all_gather_np(handle)
        oks = ctx._all_gather_np(np.array([ok], np.int32)).reshape(-1)
        peers = [0] * ctx.world
        if oks.all():
            for r in range(ctx.world):
                if r == ctx.rank:
                    peers[r] = self.local
                    continue
                try:
                    peers[r] = ops.peer_open(handles[r])
                except Exception:  # noqa: BLE001
                    ok = 0
                    break
        else:
            ok = 0
        self.peers = peers
        if not ctx._all_gather_np(np.array([ok], np.int32)).all():
            self.release()
            self.failed = True
            return False
        self.capacity = cap
        return True


class Context:
    def __init__(self, group=None, local_ops=None, samples_per_rank: int = 1024):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.ops = local_ops if local_ops is not None else CudaLocalOps()
        self.samples_per_rank = samples_per_rank
        self.last_stats = {}
        self.peer = PeerExchange(self)
        # BCB_DIST_PEER=0: always exchange through NCCL all-to-all; BCB_DIST_PROFILE=1: synchronise after every phase
        # and record last_stats["phases_ms"] (diagnostics only -- the synchronisation costs time)
        self.use_peer_memory = os.environ.get("BCB_DIST_PEER", "1") != "0"
        self.use_histogram_plan = os.environ.get("BCB_DIST_HISTOGRAM", "1") != "0"  # 0: always sample (A/B comparison)
        self.use_digit_exchange = os.environ.get("BCB_DIST_DIGIT_EXCHANGE", "1") != "0"  # 0: partition pass + local sort (A/B)
        self.digit_exchange_wide = os.environ.get("BCB_DIST_DIGIT_EXCHANGE", "1") == "2"  # also for 64-bit keys
        self.profile = os.environ.get("BCB_DIST_PROFILE", "0") == "1"
        self._flag = None

    def _phase(self, name):
        if not self.profile:
            return
        import time
        if self.ops.device_type == "cuda":
            torch.cuda.synchronize()
        now = time.perf_counter()
        if name is None:
            self._phases, self._t0 = {}, now
            return
        self._phases[name] = self._phases.get(name, 0.0) + (now - self._t0) * 1e3
        self._t0 = now

    def _stream_barrier(self):
        """Stream-ordered barrier: everything enqueued before it on every rank has completed before anything after it
        starts on any rank (a one-element all-reduce; no host synchronisation)."""
        if self._flag is None:
            self._flag = torch.zeros(1, dtype=torch.int32, device="cuda" if self.ops.device_type == "cuda" else "cpu")
        dist.all_reduce(self._flag, group=self.group)

    # -- helpers ----------------------------------------------------------------------------------------
    def _all_gather_np(self, arr: np.ndarray) -> np.ndarray:
        """all-gather a small host array (same shape on every rank) -> [world, ...]."""
        if self.world == 1:
            return arr[None]
        dev = "cuda" if self.ops.device_type == "cuda" else "cpu"
        t = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1).copy()).to(dev, non_blocking=True)
        out = torch.empty(self.world * t.numel(), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(out, t, group=self.group)
        return out.cpu().numpy().view(arr.dtype).reshape((self.world,) + arr.shape)  # one device->host copy

    def _gather_records(self, part, w: int):
        """All-gather of the 16-byte record {partial at byte 0, "shard not empty" at byte 8} -> device tensor uint8[world * 16]."""
        dev = "cuda" if self.ops.device_type == "cuda" else "cpu"
        packed = torch.zeros(16, dtype=torch.uint8, device=dev)
        if part is not None:
            packed[:w] = part.view(torch.uint8).reshape(-1)[:w]
            packed[8] = 1
        if self.world == 1:
            return packed
        out = torch.empty(self.world * 16, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(out, packed, group=self.group)
        return out

    def _gather_partials(self, part, np_dt):
        """One all-gather of (partial value, "shard not empty") -> (partials[world], present[world]).  ``part`` is the
        1-element DEVICE tensor a local reduction left behind, or None for an empty shard: the record is assembled and
        gathered on the device, so the call costs ONE device->host synchronisation (the partial never visits the host on
        its own)."""
        w = np.dtype(np_dt).itemsize
        allp = self._gather_records(part, w).cpu().numpy().reshape(self.world, 16)   # the one device->host copy
        return allp[:, :w].copy().view(np_dt).reshape(-1), allp[:, 8].astype(np.int32)

    def _all_to_all(self, src: torch.Tensor, send_counts: np.ndarray, recv_counts: np.ndarray) -> torch.Tensor:
        """variable-size all-to-all of contiguous row slices (bitwise; rows may be wider than one element)."""
        out = self.ops.empty(int(recv_counts.sum()), src)
        if self.world == 1:
            out.copy_(src)
            return out
        row = src.element_size() * (src.numel() // src.shape[0] if src.shape[0] else 1)
        s8 = src.contiguous().view(torch.uint8).reshape(-1)
        o8 = out.view(torch.uint8).reshape(-1)
        dist.all_to_all_single(o8, s8, output_split_sizes=[int(c) * row for c in recv_counts],
                               input_split_sizes=[int(c) * row for c in send_counts], group=self.group)
        return out

    # -- sort -------------------------------------------------------------------------------------------
    def sort(self, keys: torch.Tensor, values: torch.Tensor | None = None, descending: bool = False):
        """Globally sorts the block-distributed range.  ``keys`` (and ``values``) are this rank's shard and are used
        as scratch; returns this rank's slice of the sorted range (sizes differ slightly between ranks)."""
        P = self.world
        n_local = keys.shape[0]
        if P == 1:
            self.ops.sort(keys, values, descending)
            return keys if values is None else (keys, values)
        code = dtype_code(keys.dtype)
        s = self.samples_per_rank

        def sample_splitters(src):
            pos = (np.arange(s, dtype=np.int64) * max(n_local, 1)) // s if n_local else np.zeros(0, np.int64)
            bits = self.ops.gather_bits(src, pos) if n_local else np.zeros(0, _NP_UINT[src.element_size()])
            tk = transformed_keys(bits, code, not descending)
            if tk.size < s:  # empty shard: pad with the maximum so it does not pull splitters down
                tk = np.concatenate([tk, np.full(s - tk.size, np.iinfo(np.uint64).max, np.uint64)])
            return select_splitters(self._all_gather_np(tk), P)

        self._phase(None)
        vb = 0 if values is None else values.element_size() * (values.numel() // max(1, values.shape[0]))
        peer_ok = (self.use_peer_memory and hasattr(self.ops, "partition_scatter") and not self.peer.failed and P - 1 <= 7
                   and vb in (0, 4, 8))
        # Preferred plan: ONE stable partition pass (bucket = number of splitters <= transformed key) whose stores go
        # straight into the destination ranks' receive buffers over NVLink -> one local sort out of the receive buffer.
        # Splitters and send counts come, when the keys allow it, from ONE exchange: the all-gathered 256-bin
        # histograms of the most significant digit deal whole digit values to the ranks (exact counts, no sampling, no
        # count pass); skewed keys (a few digit values hold most of them) fall back to regular samples of the
        # unsorted shard + a count pass.
        if peer_ok and self.use_histogram_plan and hasattr(self.ops, "top_histogram"):
            # Best plan: the exchange is ONE of the sort's radix passes (digit_exchange_plan) -- the pass over the most
            # significant digit writes into the owners' receive buffers, the owners sort every digit value's segment by
            # the remaining digits: as many passes over the data as on one GPU.  Needs the warp-specialised pass kernel:
            # 32-bit keys (payload 0 / 4 / 8 bytes) or 64-bit keys alone with an injective transform (the decision depends
            # on shapes only: the same on every rank).
            ksize = keys.element_size()
            # (64-bit keys alone are covered by the kernels too, but measured slower than the partition pass + local sort:
            # the warp-specialised kernel's 21504-key tiles lose more per pass than the saved ninth pass gains --
            # 40.7 against 44.8 Gkeys/s on 2 GPUs; BCB_DIST_DIGIT_EXCHANGE=2 forces the plan for them)
            digit_ok = (self.use_digit_exchange and hasattr(self.ops, "exchange_scatter")
                        and (ksize == 4 or (self.digit_exchange_wide and ksize == 8 and vb == 0
                                            and not (keys.dtype == torch.float64 and descending))))
            hist = self.ops.top_histogram(keys, descending) if n_local else np.zeros(256, np.int64)
            all_hist = self._all_gather_np(hist)
            self._phase("histogram")
            if digit_ok:
                done = self._sort_digit_exchange(keys, values, descending, vb, all_hist)
                if done is not None:
                    return done
            plan = histogram_plan(all_hist, P, keys.element_size() * 8)
            if plan is not None:
                done = self._sort_peer(keys, values, descending, plan[0], vb, counts=plan[1])
                if done is not None:
                    return done
        splitters = sample_splitters(keys)
        self._phase("sample")
        if peer_ok and not self.peer.failed:
            done = self._sort_peer(keys, values, descending, splitters, vb)
            if done is not None:
                return done
        part = self.ops.partition(keys, values, splitters, descending) if hasattr(self.ops, "partition") else None
        if part is not None:
            src_keys, src_vals, send = part
            plan = "partition"
        else:
            # Fallback (payload sizes / rank counts the partition kernel does not cover): local stable sort, then
            # cut the sorted shard at the splitters by binary search.
            self.ops.sort(keys, values, descending)
            splitters = sample_splitters(keys)
            points = self.ops.partition_points(keys, splitters, descending) if n_local else np.zeros(P - 1, np.int64)
            send = exchange_plan(points, n_local)
            src_keys, src_vals = keys, values
            plan = "sort-and-cut"
        self._phase("partition")
        counts = self._all_gather_np(np.asarray(send, dtype=np.int64))          # counts[src][dst]
        recv = counts[:, self.rank].copy()
        # exchange: contiguous slices, received in source-rank order
        out_keys = self._all_to_all(src_keys, send, recv)
        out_vals = self._all_to_all(src_vals, send, recv) if values is not None else None
        self._phase("all_to_all")
        # 5. final local stable sort of the P received runs
        self.ops.sort(out_keys, out_vals, descending)
        self._phase("sort")
        self.last_stats = {"plan": plan, "phases_ms": dict(getattr(self, "_phases", {})) if self.profile else None, "sent": int(send.sum() - send[self.rank]), "received": int(recv.sum()),
                           "imbalance": float(counts.sum(axis=0).max() * P / max(1, counts.sum()))}
        return out_keys if values is None else (out_keys, out_vals)

    def _sort_digit_exchange(self, keys, values, descending, vb, all_hist):
        """The digit-exchange plan (see sort).  Returns None -- collectively -- when whole digit values cannot be dealt
        evenly, the buffers cannot be mapped or the kernels do not cover the shape; the shard is then still untouched."""
        P, me = self.world, self.rank
        plan = digit_exchange_plan(all_hist, P)
        if plan is None:
            return None
        owner, first, seg_begin, seg_len, recv_tot, span, imbalance = plan
        ksize = keys.element_size()
        max_span = int(span.max())
        val_off = (max_span * ksize + 255) & ~255                        # values region of every receive buffer
        # (the all-gather of the histograms also ordered this call after every rank's previous use of its receive buffer)
        if not self.peer.ensure(max(256, val_off + max_span * vb)):
            return None
        dst_k = np.asarray(self.peer.peers, dtype=np.uint64)[owner]
        dst_v = dst_k + np.uint64(val_off)
        self._phase("plan")
        if self.ops.device_type == "cuda":  # the pass moves whole 16-byte chunks: a misaligned view is copied first
            if keys.data_ptr() % 16:
                keys = keys.clone()
            if values is not None and values.data_ptr() % 16:
                values = values.clone()
        if not self.ops.exchange_scatter(keys, values, descending, dst_k, dst_v, first[me]):
            return None
        self._stream_barrier()                                           # all incoming runs have landed
        self._phase("exchange pass")
        n_out = int(recv_tot[me])
        out_keys = self.ops.empty(n_out, keys)
        out_vals = self.ops.empty(n_out, values) if values is not None else None
        mine = owner == me
        self.ops.sort_segments(self.peer.local, self.peer.local + val_off, out_keys, out_vals, descending, seg_begin[mine], seg_len[mine])
        self._phase("segment passes")
        sent = int(all_hist[me].sum() - all_hist[me][mine].sum())
        self.last_stats = {"plan": "digit-exchange", "splitters": "top-digit histogram",
                           "phases_ms": dict(self._phases) if self.profile else None,
                           "sent": sent, "received": n_out, "imbalance": imbalance}
        return out_keys if values is None else (out_keys, out_vals)

    def _sort_peer(self, keys, values, descending, splitters, vb, counts=None):
        """The peer-memory plan (see sort).  Returns None -- collectively -- when the buffers cannot be mapped.
        ``counts``: the P x P matrix when the caller already has it (histogram plan); else a count pass + all-gather."""
        P, me = self.world, self.rank
        ksize = keys.element_size()
        from_histogram = counts is not None
        if counts is None:
            send = self.ops.partition_counts(keys, splitters, descending)
            self._phase("counts")
            counts = self._all_gather_np(send)                           # counts[src][dst]
        else:
            send = counts[me]
        recv_tot = counts.sum(axis=0)
        max_recv = int(recv_tot.max())
        val_off = (max_recv * ksize + 255) & ~255                        # values region of every receive buffer
        # the all-gather above also orders this call after every rank's previous use of its receive buffer
        if not self.peer.ensure(max(256, val_off + max_recv * vb)):  # (at least a token buffer: an all-empty range still maps peers)
            return None
        offs = np.cumsum(counts, axis=0) - counts                        # offs[src][dst]: where src's slice starts at dst
        dst_k = [self.peer.peers[d] + int(offs[me][d]) * ksize for d in range(P)]
        dst_v = [self.peer.peers[d] + val_off + int(offs[me][d]) * vb for d in range(P)]
        self._phase("plan")
        self.ops.partition_scatter(keys, values, splitters, descending, dst_k, dst_v)
        self._stream_barrier()                                           # all incoming stores have landed
        self._phase("scatter")
        n_out = int(recv_tot[me])
        out_keys = self.ops.empty(n_out, keys)
        out_vals = self.ops.empty(n_out, values) if values is not None else None
        self.ops.sort_copy(self.peer.local, out_keys, self.peer.local + val_off, out_vals, descending)
        self._phase("sort")
        self.last_stats = {"plan": "peer-scatter", "splitters": "top-digit histogram" if from_histogram else "regular samples",
                           "phases_ms": dict(self._phases) if self.profile else None,
                           "sent": int(send.sum() - send[me]), "received": n_out,
                           "imbalance": float(recv_tot.max() * P / max(1, counts.sum()))}
        return out_keys if values is None else (out_keys, out_vals)

    # -- scan -------------------------------------------------------------------------------------------
    def _scan(self, x, out, exclusive, init, op):
        if self.world == 1:
            self.ops.scan(x, out, 1 if exclusive else 0, init, op)
            return out
        part = self.ops.reduce_to(x, op, out.dtype) if x.numel() else None
        np_dt = NP_OF_CODE[dtype_code(out.dtype)].type
        if hasattr(self.ops, "scan_with_carry"):
            # everything stays on the device: partial -> all-gather -> carry folded in rank order by a one-thread kernel ->
            # the scan reads its seed from device memory.  No host synchronisation: the call is enqueue-and-return.
            records = self._gather_records(part, np.dtype(np_dt).itemsize)
            if x.numel():
                self.ops.scan_with_carry(x, out, exclusive, init, op, records, self.rank)
            return out
        partials, present = self._gather_partials(part, np_dt)
        carry = None if not exclusive else np_dt(0 if init is None else init)
        fn = _NP_OPS[op]
        with np.errstate(over="ignore"):
            for r in range(self.rank):
                if present[r]:
                    carry = partials[r] if carry is None else np_dt(fn(carry, partials[r]))
        if x.numel():
            if carry is None:
                self.ops.scan(x, out, 0, None, op)      # inclusive, nothing before this rank
            else:
                self.ops.scan(x, out, 1 if exclusive else 2, carry, op)  # 2 = inclusive scan seeded with a carry
        return out

    def exclusive_scan(self, x, out, init=0, op="plus"):
        return self._scan(x, out, True, init, op)

    def inclusive_scan(self, x, out, op="plus"):
        return self._scan(x, out, False, None, op)

    # -- reduce -----------------------------------------------------------------------------------------
    def reduce(self, x, op="plus", result_dtype=None):
        """Global reduction; every rank returns the same host scalar (None if the global range is empty)."""
        rdt = result_dtype if result_dtype is not None else x.dtype
        np_dt = NP_OF_CODE[dtype_code(rdt)].type
        part = self.ops.reduce_to(x, op, rdt) if x.numel() else None
        partials, present = self._gather_partials(part, np_dt)
        acc = None
        fn = _NP_OPS[op]
        with np.errstate(over="ignore"):
            for r in range(self.world):
                if present[r]:
                    acc = partials[r] if acc is None else np_dt(fn(acc, partials[r]))
        return acc

    def accumulate(self, x, init, op="plus"):
        r = self.reduce(x, op)
        np_dt = NP_OF_CODE[dtype_code(x.dtype)].type
        if r is None:
            return np_dt(init)
        with np.errstate(over="ignore"):
            return np_dt(_NP_OPS[op](np_dt(init), r))

