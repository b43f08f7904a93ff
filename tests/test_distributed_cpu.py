"""Host-side protocol of the multi-GPU path (compute_b200/distributed.py) on CPU: unit tests of the numpy helpers
and a world_size-2 gloo run in which the local GPU primitives are replaced by the CPU oracle (test infrastructure
only -- the product ships CudaLocalOps alone).  The distributed result must equal the single-device oracle result
bit for bit (sorts, integer scans / reductions)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from compute_b200 import distributed as cbd
from compute_b200.core import dtype_code

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_transformed_keys_match_oracle():
    rng = np.random.default_rng(1)
    for name in oracle.DTYPES:
        npdt = np.dtype(oracle.NP_DTYPES[name])
        w = npdt.itemsize
        bits = rng.integers(0, 256, size=200 * w, dtype=np.uint8).view({1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[w])
        for asc in (True, False):
            got = cbd.transformed_keys(bits, dtype_code(npdt), asc)
            exp = np.array([oracle.radix_key(name, asc, int(b)) for b in bits], dtype=np.uint64)
            np.testing.assert_array_equal(got, exp, err_msg=f"{name} asc={asc}")


def test_splitters_and_plan():
    s = np.arange(4000, dtype=np.uint64)[::-1].copy()
    sp = cbd.select_splitters(s, 4)
    np.testing.assert_array_equal(sp, [1000, 2000, 3000])
    assert cbd.select_splitters(s, 1).size == 0
    np.testing.assert_array_equal(cbd.exchange_plan(np.array([3, 3, 10]), 12), [3, 0, 7, 2])
    p = np.array([5, 7, 250], dtype=np.uint8)
    assert cbd.fold_carry(p, 3, "plus", 10) == np.uint8((10 + 5 + 7 + 250) % 256)
    assert cbd.fold_carry(p, 0, "plus", None) is None
    assert cbd.fold_carry(p, 2, "max", None) == 7


def test_histogram_plan_deals_whole_digit_values():
    rng = np.random.default_rng(7)
    for world in (2, 3, 4, 8):
        keys = [rng.integers(0, 2**32, size=50_000 + 1000 * r, dtype=np.uint64) for r in range(world)]
        allh = np.stack([np.bincount((k >> np.uint64(24)).astype(np.int64), minlength=256) for k in keys])
        plan = cbd.histogram_plan(allh, world, 32)
        assert plan is not None
        splitters, counts, imbalance = plan
        assert splitters.size == world - 1 and np.all(np.diff(splitters.astype(np.int64)) >= 0) and imbalance < 1.06
        for src, k in enumerate(keys):  # the counts ARE the bucket sizes the exchange pass will produce
            bucket = np.searchsorted(splitters, k, side="right")
            np.testing.assert_array_equal(np.bincount(bucket, minlength=world), counts[src])
    # skewed keys (everything in three digit values): no even deal of whole digit values -> None (the caller samples)
    skew = np.zeros((4, 256), dtype=np.int64)
    skew[:, 7] = 1000
    skew[:, 8] = 10
    skew[:, 200] = 5
    assert cbd.histogram_plan(skew, 4, 32) is None
    assert cbd.histogram_plan(np.zeros((2, 256), np.int64), 2, 32) is None  # globally empty range


def test_digit_exchange_plan_places_every_run():
    rng = np.random.default_rng(11)
    for world in (2, 3, 8):
        keys = [rng.integers(0, 2**32, size=30_000 + 777 * r, dtype=np.uint64) for r in range(world)]
        allh = np.stack([np.bincount((k >> np.uint64(24)).astype(np.int64), minlength=256) for k in keys])
        owner, first, seg_begin, seg_len, recv, span, imbalance = cbd.digit_exchange_plan(allh, world)
        assert np.all(np.diff(owner) >= 0) and owner[0] == 0 and owner[-1] == world - 1 and imbalance < 1.06
        assert int(recv.sum()) == sum(k.size for k in keys) and np.all(seg_begin % 32 == 0)
        # emulate the exchange: every (src, digit) run copied to first[src][g] in the owner's receive buffer, then every
        # segment sorted on its own and the segments concatenated = the global sort
        bufs = [np.full(int(span[d]), -1, dtype=np.int64) for d in range(world)]
        for src, k in enumerate(keys):
            top = (k >> np.uint64(24)).astype(np.int64)
            for g in np.unique(top):
                run = k[top == g]
                o = bufs[int(owner[g])]
                assert np.all(o[first[src][g]:first[src][g] + run.size] == -1)   # runs never overlap
                o[first[src][g]:first[src][g] + run.size] = run
        outs = []
        for d in range(world):
            for g in np.flatnonzero(owner == d):
                seg = bufs[d][seg_begin[g]:seg_begin[g] + seg_len[g]]
                assert np.all(seg >= 0) and np.all(seg >> 24 == g)
                outs.append(np.sort(seg, kind="stable"))
            assert sum(int(seg_len[g]) for g in np.flatnonzero(owner == d)) == recv[d]
        np.testing.assert_array_equal(np.concatenate(outs), np.sort(np.concatenate(keys)).astype(np.int64))
    skew = np.zeros((4, 256), dtype=np.int64)
    skew[:, 7] = 1000
    assert cbd.digit_exchange_plan(skew, 4) is None
    assert cbd.digit_exchange_plan(np.zeros((2, 256), np.int64), 2) is None


def test_digit_exchange_plan_invariants_on_random_histograms():
    """Property check of the placement: for random (also ragged, sparse, partly empty) histograms every (source, digit)
    run lies inside its digit's segment, runs of one segment tile it exactly in source order, segments of one owner are
    disjoint, aligned and ascending, and the owners' receive counts add up."""
    rng = np.random.default_rng(2024)
    accepted = 0
    for trial in range(300):
        world = int(rng.integers(2, 9))
        kind = trial % 4
        if kind == 0:
            h = rng.integers(0, 5000, size=(world, 256))
        elif kind == 1:   # sparse: most digit values unused
            h = rng.integers(0, 5000, size=(world, 256)) * (rng.random((1, 256)) < 0.2)
        elif kind == 2:   # some ranks hold nothing
            h = rng.integers(0, 3000, size=(world, 256))
            h[rng.random(world) < 0.3] = 0
        else:             # tiny counts
            h = rng.integers(0, 3, size=(world, 256))
        h = h.astype(np.int64)
        plan = cbd.digit_exchange_plan(h, world, max_imbalance=1e9)
        if h.sum() == 0:
            assert plan is None
            continue
        accepted += 1
        owner, first, seg_begin, seg_len, recv, span, _ = plan
        tot = h.sum(axis=0)
        np.testing.assert_array_equal(seg_len, tot)
        assert np.all(np.diff(owner) >= 0) and owner.min() >= 0 and owner.max() < world
        assert np.all(seg_begin % 32 == 0)
        for d in range(world):
            mine = np.flatnonzero(owner == d)
            assert int(tot[mine].sum()) == int(recv[d])
            end = 0
            for g in mine:                                   # ascending, disjoint, inside the span
                assert seg_begin[g] >= end
                end = seg_begin[g] + seg_len[g]
            assert end <= span[d] or mine.size == 0
        # runs of a segment tile it exactly, in source order
        ends = first + h
        assert np.all(first[0] == seg_begin)
        assert np.all(first[1:] == ends[:-1])
        assert np.all(ends[-1] == seg_begin + seg_len)
    assert accepted > 250


class OracleLocalOps:
    """CPU stand-in for CudaLocalOps used ONLY by this test: same interface, oracle semantics, CPU tensors."""
    device_type = "cpu"
    _T = {torch.int8: np.int8, torch.uint8: np.uint8, torch.int16: np.int16, torch.int32: np.int32, torch.int64: np.int64,
          torch.float32: np.float32, torch.float64: np.float64}

    def _np(self, t):
        return t.numpy()

    def sort(self, keys, values, descending):
        k = self._np(keys)
        if values is None:
            k[:] = oracle.radix_sort(k, descending)
        else:
            v = self._np(values)
            sk, sv = oracle.radix_sort(k, descending, v)
            k[:] = sk
            v[:] = sv

    def partition_points(self, sorted_keys, splitters, descending):
        k = self._np(sorted_keys)
        w = k.dtype.itemsize
        bits = k.view({1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[w])
        tk = cbd.transformed_keys(bits, dtype_code(k.dtype), not descending)
        return np.searchsorted(tk, splitters, side="left").astype(np.int64)

    def partition(self, keys, values, splitters, descending):
        if getattr(self, "force_sort_and_cut", False):
            return None
        k = self._np(keys)
        bits = k.view({1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[k.dtype.itemsize])
        tk = cbd.transformed_keys(bits, dtype_code(k.dtype), not descending)
        bucket = np.searchsorted(splitters, tk, side="right")
        order = np.argsort(bucket, kind="stable")
        counts = np.bincount(bucket, minlength=splitters.size + 1).astype(np.int64)
        ok = torch.from_numpy(k[order].copy())
        ov = torch.from_numpy(self._np(values)[order].copy()) if values is not None else None
        return ok, ov, counts

    # ---- emulated peer memory: POSIX shared memory stands in for CUDA IPC, memmove for the NVLink stores ----
    def _buckets(self, keys, splitters, descending):
        k = self._np(keys)
        bits = k.view({1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[k.dtype.itemsize])
        tk = cbd.transformed_keys(bits, dtype_code(k.dtype), not descending)
        return np.searchsorted(splitters, tk, side="right")

    def top_histogram(self, keys, descending):
        k = self._np(keys)
        w = k.dtype.itemsize
        bits = k.view({1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[w])
        tk = cbd.transformed_keys(bits, dtype_code(k.dtype), not descending)
        return np.bincount((tk >> np.uint64(8 * w - 8)).astype(np.int64), minlength=256).astype(np.int64)

    def exchange_scatter(self, keys, values, descending, dst_keys, dst_values, dst_first):
        """Stand-in for bcb_radix_exchange_scatter: stable partition by the most significant digit of the transformed
        key; every digit value's run is stored where the plan says (memmove = the bulk copies over NVLink)."""
        import ctypes
        if getattr(self, "no_digit_exchange", False):
            return False
        k = self._np(keys)
        w = k.dtype.itemsize
        bits = k.view({1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[w])
        top = (cbd.transformed_keys(bits, dtype_code(k.dtype), not descending) >> np.uint64(8 * w - 8)).astype(np.int64)
        v = self._np(values) if values is not None else None
        row = 0 if v is None else v.dtype.itemsize * (v.size // max(1, v.shape[0]))
        for g in np.unique(top):
            sel = np.flatnonzero(top == g)  # ascending positions: stable
            kb = np.ascontiguousarray(k[sel])
            ctypes.memmove(int(dst_keys[g]) + int(dst_first[g]) * w, kb.ctypes.data, kb.nbytes)
            if v is not None:
                vb = np.ascontiguousarray(v[sel])
                ctypes.memmove(int(dst_values[g]) + int(dst_first[g]) * row, vb.ctypes.data, vb.nbytes)
        return True

    def sort_segments(self, recv_keys_ptr, recv_values_ptr, out_keys, out_values, descending, seg_begin, seg_len):
        """Stand-in for bcb_radix_sort_segments: every segment stably sorted on its own, results back to back."""
        import ctypes
        k = self._np(out_keys)
        v = self._np(out_values) if out_values is not None else None
        row = 0 if v is None else v.dtype.itemsize * (v.size // max(1, v.shape[0]))
        pos = 0
        for b, l in zip(seg_begin, seg_len):
            b, l = int(b), int(l)
            if l == 0:
                continue
            seg_k = k[pos:pos + l]
            ctypes.memmove(seg_k.ctypes.data, int(recv_keys_ptr) + b * k.dtype.itemsize, l * k.dtype.itemsize)
            if v is None:
                seg_k[:] = oracle.radix_sort(seg_k.copy(), descending)
            else:
                seg_v = v[pos:pos + l]
                ctypes.memmove(seg_v.ctypes.data, int(recv_values_ptr) + b * row, l * row)
                sk, sv = oracle.radix_sort(seg_k.copy(), descending, seg_v.copy())
                seg_k[:] = sk
                seg_v[:] = sv
            pos += l

    def partition_counts(self, keys, splitters, descending):
        return np.bincount(self._buckets(keys, splitters, descending), minlength=splitters.size + 1).astype(np.int64)

    def partition_scatter(self, keys, values, splitters, descending, dst_keys, dst_values):
        import ctypes
        bucket = self._buckets(keys, splitters, descending)
        for b in range(splitters.size + 1):
            sel = np.flatnonzero(bucket == b)  # ascending positions: stable
            if sel.size == 0:
                continue
            kb = np.ascontiguousarray(self._np(keys)[sel])
            ctypes.memmove(int(dst_keys[b]), kb.ctypes.data, kb.nbytes)
            if values is not None:
                vb = np.ascontiguousarray(self._np(values)[sel])
                ctypes.memmove(int(dst_values[b]), vb.ctypes.data, vb.nbytes)

    def sort_copy(self, src_keys_ptr, out_keys, src_values_ptr, out_values, descending):
        import ctypes
        k = self._np(out_keys)
        ctypes.memmove(k.ctypes.data, int(src_keys_ptr), k.nbytes)
        if out_values is not None:
            v = self._np(out_values)
            ctypes.memmove(v.ctypes.data, int(src_values_ptr), v.nbytes)
        self.sort(out_keys, out_values, descending)

    def peer_alloc(self, nbytes):
        import ctypes
        from multiprocessing import shared_memory
        if getattr(self, "fail_peer_alloc", False):
            raise RuntimeError("no peer memory here")
        shm = shared_memory.SharedMemory(create=True, size=nbytes)
        self._shm = getattr(self, "_shm", {})
        addr = ctypes.addressof(ctypes.c_char.from_buffer(shm.buf))
        self._shm[addr] = shm
        handle = np.zeros(64, dtype=np.uint8)
        name = shm.name.encode()
        handle[: len(name)] = np.frombuffer(name, dtype=np.uint8)
        return addr, handle

    def peer_open(self, handle):
        import ctypes
        from multiprocessing import shared_memory
        name = bytes(handle[handle != 0]).decode()
        shm = shared_memory.SharedMemory(name=name)
        addr = ctypes.addressof(ctypes.c_char.from_buffer(shm.buf))
        self._shm[addr] = shm
        return addr

    def peer_close(self, ptr):
        self._shm.pop(ptr)  # (the mapping stays until process exit: ctypes still holds an export of the buffer)

    def peer_free(self, ptr):
        shm = self._shm.pop(ptr)
        try:
            shm.unlink()
        except FileNotFoundError:
            pass

    def gather_bits(self, keys, positions):
        k = self._np(keys)
        return k.view({1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[k.dtype.itemsize])[positions]

    def reduce_to(self, x, op, result_dtype):
        r = oracle.reduce(self._np(x), op, self._T[result_dtype])
        return torch.from_numpy(np.array([r], dtype=self._T[result_dtype]))

    def scan(self, x, out, mode, init, op):
        xs = self._np(x)
        o = self._np(out)
        if mode == 1:
            o[:] = oracle.scan(xs, op, True, init, out_dtype=o.dtype)
        elif mode == 0:
            o[:] = oracle.scan(xs, op, False, 0, out_dtype=o.dtype)
        else:  # inclusive seeded with a carry == exclusive scan of x with init, shifted by one, plus the last element
            ex = oracle.scan(xs, op, True, init, out_dtype=o.dtype)
            o[:-1] = ex[1:]
            o[-1:] = oracle.scan(np.concatenate([ex[-1:], xs[-1:].astype(o.dtype)]), op, False, 0)[-1:]

    def empty(self, n, like):
        return torch.empty((n,) + tuple(like.shape[1:]), dtype=like.dtype)


class OracleLocalOpsDeviceCarry(OracleLocalOps):
    """The same with the stand-in for bcb_scan_with_carry: the carry is folded from the all-gathered 16-byte records
    ({partial at byte 0, "block not empty" at byte 8}) in rank order, as fold_carry_kernel does on the device."""

    def scan_with_carry(self, x, out, exclusive, init, op, records, rank):
        o = self._np(out)
        rec = records.numpy().reshape(-1, 16)
        carry = o.dtype.type(init if (exclusive and init is not None) else 0) if exclusive else None
        fn = cbd._NP_OPS[op]
        with np.errstate(over="ignore"):
            for r in range(rank):
                if rec[r, 8]:
                    v = rec[r, :o.dtype.itemsize].copy().view(o.dtype)[0]
                    carry = v if carry is None else o.dtype.type(fn(carry, v))
        if carry is None:
            self.scan(x, out, 0, None, op)
        else:
            self.scan(x, out, 1 if exclusive else 2, carry, op)


def _scan_cuts(world, n):
    if world == 2:
        return [0, 3000, n]
    return [0, 3000, 3000] + [n * r // world for r in range(3, world)] + [n]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ops = OracleLocalOps()
        ctx = cbd.Context(local_ops=ops, samples_per_rank=64)
        rng = np.random.default_rng(42)
        out = {}
        # --- sort: keys only (float with duplicates, descending) and by key (int keys, 8-byte payload rows) ---
        n_total = 20_000
        allk = rng.integers(-50, 50, size=n_total).astype(np.float32)
        allk[::97] = -0.0
        lo, hi = rank * n_total // world, (rank + 1) * n_total // world
        shard = torch.from_numpy(allk[lo:hi].copy())
        out["sort_f32_desc"] = ctx.sort(shard, None, descending=True).numpy().copy()
        keys = rng.integers(-1000, 1000, size=n_total).astype(np.int32)
        vals = np.stack([np.arange(n_total, dtype=np.int32), -np.arange(n_total, dtype=np.int32)], axis=1)
        k, v = ctx.sort(torch.from_numpy(keys[lo:hi].copy()), torch.from_numpy(vals[lo:hi].copy()))
        out["pairs_k"], out["pairs_v"] = k.numpy().copy(), v.numpy().copy()
        out["stats"] = dict(ctx.last_stats)
        # a second, larger call on the same Context: the receive buffers grow (collective re-allocation)
        big = rng.integers(0, 2**32, size=60_000, dtype=np.uint64).astype(np.uint32)
        blo, bhi = rank * 60_000 // world, (rank + 1) * 60_000 // world
        out["big"] = ctx.sort(torch.from_numpy(big[blo:bhi].copy())).numpy().copy()
        out["stats_big"] = dict(ctx.last_stats)
        # uneven blocks, rank 0 holds nothing: still the digit exchange
        ucuts = [0, 0] + [60_000 * r // (2 * world) for r in range(2, world)] + [60_000]
        out["big_uneven"] = ctx.sort(torch.from_numpy(big[ucuts[rank]:ucuts[rank + 1]].copy())).numpy().copy()
        out["stats_uneven"] = dict(ctx.last_stats)
        # the plans behind the digit exchange: partition pass into the peers' buffers + local sort (same bytes)
        ctx.use_digit_exchange = False
        out["big_ps"] = ctx.sort(torch.from_numpy(big[blo:bhi].copy())).numpy().copy()
        out["stats_big_ps"] = dict(ctx.last_stats)
        ctx.use_digit_exchange = True
        ops.no_digit_exchange = True    # the kernels refuse the shape: every rank moves on to the next plan together
        k3, v3 = ctx.sort(torch.from_numpy(keys[lo:hi].copy()), torch.from_numpy(vals[lo:hi].copy()))
        out["pairs_k3"], out["pairs_v3"] = k3.numpy().copy(), v3.numpy().copy()
        out["stats3"] = dict(ctx.last_stats)
        ops.no_digit_exchange = False
        ctx.use_peer_memory = False     # NCCL-style all-to-all plan: same bytes
        k1, v1 = ctx.sort(torch.from_numpy(keys[lo:hi].copy()), torch.from_numpy(vals[lo:hi].copy()))
        out["pairs_k1"], out["pairs_v1"] = k1.numpy().copy(), v1.numpy().copy()
        out["stats1"] = dict(ctx.last_stats)
        ops.force_sort_and_cut = True   # the fallback plan must give the same bytes
        k2, v2 = ctx.sort(torch.from_numpy(keys[lo:hi].copy()), torch.from_numpy(vals[lo:hi].copy()))
        out["pairs_k2"], out["pairs_v2"] = k2.numpy().copy(), v2.numpy().copy()
        out["stats2"] = dict(ctx.last_stats)
        ops.force_sort_and_cut = False
        # --- scans and reductions on unequal blocks (rank 0 gets 1/3) ---
        x = rng.integers(-2**31, 2**31 - 1, size=9001).astype(np.int32)
        cuts = _scan_cuts(world, x.size)   # unequal blocks; with 3 ranks the middle one is EMPTY
        mine = torch.from_numpy(x[cuts[rank]:cuts[rank + 1]].copy())
        o = torch.empty_like(mine)
        ctx.exclusive_scan(mine, o, 11)
        out["excl"] = o.numpy().copy()
        ctx.inclusive_scan(mine, o, "plus")
        out["incl"] = o.numpy().copy()
        ctx.inclusive_scan(mine, o, "max")
        out["incl_max"] = o.numpy().copy()
        # the same scans with the carry folded from the gathered records (the path the CUDA ops take: no host round trip)
        ctxd = cbd.Context(local_ops=OracleLocalOpsDeviceCarry(), samples_per_rank=64)
        ctxd.exclusive_scan(mine, o, 11)
        out["excl_d"] = o.numpy().copy()
        ctxd.inclusive_scan(mine, o, "plus")
        out["incl_d"] = o.numpy().copy()
        ctxd.inclusive_scan(mine, o, "max")
        out["incl_max_d"] = o.numpy().copy()
        out["sum"] = ctx.reduce(mine)
        out["min"] = ctx.reduce(mine, "min")
        out["acc"] = ctx.accumulate(mine, 5)
        # globally empty range, on a fresh Context (no receive buffers yet)
        ctx0 = cbd.Context(local_ops=ops, samples_per_rank=64)
        out["empty"] = ctx0.sort(torch.empty(0, dtype=torch.int32)).numel()
        ctx0.peer.release()
        # peer mapping unavailable on one rank only: every rank must fall back together
        ctx2 = cbd.Context(local_ops=ops, samples_per_rank=64)
        ops.fail_peer_alloc = (rank == 1)
        out["fb"] = ctx2.sort(torch.from_numpy(allk[lo:hi].copy()), None, descending=True).numpy().copy()
        out["stats_fb"] = dict(ctx2.last_stats)
        ops.fail_peer_alloc = False
        ctx.peer.release()
        ctx2.peer.release()
        results[rank] = out
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 3])
def test_gloo_ranks_match_single_device_oracle(world):
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    res = [results[r] for r in range(world)]
    rng = np.random.default_rng(42)
    n_total = 20_000
    allk = rng.integers(-50, 50, size=n_total).astype(np.float32)
    allk[::97] = -0.0
    got = np.concatenate([r["sort_f32_desc"] for r in res])
    assert got.tobytes() == oracle.radix_sort(allk, True).tobytes()
    keys = rng.integers(-1000, 1000, size=n_total).astype(np.int32)
    vals = np.stack([np.arange(n_total, dtype=np.int32), -np.arange(n_total, dtype=np.int32)], axis=1)
    ek, ev = oracle.radix_sort(keys, False, vals)
    assert np.concatenate([r["pairs_k"] for r in res]).tobytes() == ek.tobytes()
    assert np.concatenate([r["pairs_v"] for r in res]).tobytes() == ev.tobytes()
    assert np.concatenate([r["pairs_k2"] for r in res]).tobytes() == ek.tobytes()
    assert np.concatenate([r["pairs_v2"] for r in res]).tobytes() == ev.tobytes()
    assert np.concatenate([r["pairs_k1"] for r in res]).tobytes() == ek.tobytes()
    assert np.concatenate([r["pairs_v1"] for r in res]).tobytes() == ev.tobytes()
    big = rng.integers(0, 2**32, size=60_000, dtype=np.uint64).astype(np.uint32)
    assert np.concatenate([r["big"] for r in res]).tobytes() == oracle.radix_sort(big, False).tobytes()
    assert np.concatenate([r["fb"] for r in res]).tobytes() == oracle.radix_sort(allk, True).tobytes()
    assert [r["stats_fb"]["plan"] for r in res] == ["partition"] * world
    assert all(r["empty"] == 0 for r in res)
    assert res[0]["stats1"]["plan"] == "partition"
    assert np.concatenate([r["big_ps"] for r in res]).tobytes() == oracle.radix_sort(big, False).tobytes()
    assert np.concatenate([r["big_uneven"] for r in res]).tobytes() == oracle.radix_sort(big, False).tobytes()
    assert all(r["stats_uneven"]["plan"] == "digit-exchange" for r in res)
    assert np.concatenate([r["pairs_k3"] for r in res]).tobytes() == ek.tobytes()
    assert np.concatenate([r["pairs_v3"] for r in res]).tobytes() == ev.tobytes()
    # uniform 32-bit keys: the exchange is the sort's last radix pass; with that plan switched off they are dealt by the
    # top-digit histogram to the partition pass (one exchange, no sampling).  The small ints of the pair sort occupy two
    # digit values (negative / non-negative): an even deal for 2 ranks, too skewed for 3 (regular samples + count pass)
    assert all(r["stats_big"]["plan"] == "digit-exchange" for r in res), res[0]["stats_big"]
    assert all(r["stats_big_ps"]["plan"] == "peer-scatter" and r["stats_big_ps"]["splitters"] == "top-digit histogram" for r in res)
    want = ("digit-exchange", "top-digit histogram") if world == 2 else ("peer-scatter", "regular samples")
    assert all((r["stats"]["plan"], r["stats"]["splitters"]) == want for r in res), res[0]["stats"]
    want3 = "top-digit histogram" if world == 2 else "regular samples"
    assert all(r["stats3"]["plan"] == "peer-scatter" and r["stats3"]["splitters"] == want3 for r in res), res[0]["stats3"]
    assert res[0]["stats2"]["plan"] == "sort-and-cut"
    assert res[0]["stats"]["imbalance"] < 1.6
    x = rng.integers(-2**31, 2**31 - 1, size=9001).astype(np.int32)
    np.testing.assert_array_equal(np.concatenate([r["excl"] for r in res]), oracle.scan(x, "plus", True, 11))
    np.testing.assert_array_equal(np.concatenate([r["incl"] for r in res]), oracle.scan(x, "plus", False, 0))
    np.testing.assert_array_equal(np.concatenate([r["incl_max"] for r in res]), oracle.scan(x, "max", False, 0))
    np.testing.assert_array_equal(np.concatenate([r["excl_d"] for r in res]), oracle.scan(x, "plus", True, 11))
    np.testing.assert_array_equal(np.concatenate([r["incl_d"] for r in res]), oracle.scan(x, "plus", False, 0))
    np.testing.assert_array_equal(np.concatenate([r["incl_max_d"] for r in res]), oracle.scan(x, "max", False, 0))
    for r in res:
        assert r["sum"] == oracle.reduce(x, "plus")
        assert r["min"] == oracle.reduce(x, "min")
        assert r["acc"] == oracle.accumulate(x, np.int32(5), "plus")
