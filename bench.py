#!/usr/bin/env python
"""bench.py -- headline benchmark of the sort / scan / reduce path (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--log2n L]
                  [--scaling weak|strong] [--no-configs] [--no-e2e] [--no-cpu]

One "step" = one pass of the hot path over one batch of synthetic input.  The default workload is the one
BASELINE.json's metric is quoted on: compute::sort of 2^30 uniform-random uint32 keys (Gkeys/s).  Inputs are
reset to the same random keys before every step outside the timed region (mirrors perf/perf_sort.cpp:36-40);
each step is timed on the device with CUDA events on the launching stream; inputs (>= 1 GiB) are far larger than
L2, so no explicit flush is needed.  Rank 0 prints ONE JSON line.

The default run also measures every other BASELINE config at its full size and attaches the results as
`configs` (one sub-record each: value, roofline of the dominant kernel, at-size verification): sort_f32 2^29,
sort_u64 2^28, sort_pairs_u32 / sort_pairs_u32_u64 2^28 (values = original index: the permutation and stability are
checked on the device), scan_i32 / scan_f32 and reduce_i32 / reduce_f32 2^28 (checked in full against int64 / float64
references computed on the device), and a 2^24 descending float sort compared with the oracle byte for byte.
For N > 1 the headline is weak scaling (2^30 keys per GPU); `configs` then also holds the strong-scaling sort
(2^30 keys in total, BASELINE config 3 as written) and the multi-GPU scan / reduce.
`--impl reference` times the reference's own CPU-device algorithm (merge_sort_on_cpu restated in oracle/, all
host threads) on the same configuration when that fits the time budget, else on the largest power of two that does.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


WORKLOADS = {
    # name: (kind, key/elem dtype, value bytes, default log2 n, algorithmic bytes per element, metric, unit)
    "sort_u32": ("sort", "uint", 0, 30, 36, "radix sort throughput, uint32 keys", "Gkeys/s"),
    "sort_f32": ("sort", "float", 0, 29, 36, "radix sort throughput, float32 keys", "Gkeys/s"),
    "sort_u64": ("sort", "ulong", 0, 28, 136, "radix sort throughput, uint64 keys", "Gkeys/s"),
    "sort_pairs_u32": ("sort", "uint", 4, 28, 68, "sort_by_key throughput, uint32 keys + uint32 values", "Gkeys/s"),
    # the reference's own perf_sort_by_key types (perf/perf_sort_by_key.cpp:39-42): 32-bit keys, 64-bit values
    "sort_pairs_u32_u64": ("sort", "uint", 8, 28, 100, "sort_by_key throughput, uint32 keys + uint64 values", "Gkeys/s"),
    "scan_i32": ("scan", "int", 0, 28, 8, "exclusive_scan bandwidth, int32", "GB/s"),
    "scan_f32": ("scan", "float", 0, 28, 8, "exclusive_scan bandwidth, float32", "GB/s"),
    "reduce_i32": ("reduce", "int", 0, 28, 4, "reduce bandwidth, int32", "GB/s"),
    "reduce_f32": ("reduce", "float", 0, 28, 4, "reduce bandwidth, float32", "GB/s"),
}
NP = {"uint": np.uint32, "float": np.float32, "ulong": np.uint64, "int": np.int32}
SHORT = {"uint": "u32", "float": "f32", "ulong": "u64", "int": "i32"}
# every BASELINE config next to the headline (default run): name -> (workload, scaling)
EXTRA_CONFIGS = ["sort_f32", "sort_u64", "sort_pairs_u32", "sort_pairs_u32_u64", "scan_i32", "scan_f32", "reduce_i32", "reduce_f32"]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def reference_arm(args):
    """The reference's CPU implementation of the path on the host cores (oracle port of merge_sort_on_cpu /
    scan_on_cpu / reduce_on_cpu -- what compute::sort etc. execute on an OpenCL CPU device, sort.hpp:117-121).
    Runs the arm's own configuration (2^30 keys) when the whole run fits ~4 minutes, else the largest power of two
    that does; config.n is the size actually sorted."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    kind, dt, vb, log2n, bpe, metric, unit = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    full_log2n = args.log2n if args.log2n else log2n
    rng = np.random.default_rng(12345)
    budget_s = 240.0
    runs = args.warmup + args.steps
    times = []
    if kind == "sort":
        if args.sample_log2n:
            lg = args.sample_log2n
        else:  # calibrate on 2^24 keys, extrapolate with n log n
            probe = rng.integers(0, 2**32, size=1 << 24, dtype=np.uint32)
            t0 = time.perf_counter()
            oracle.merge_sort_on_cpu_u32(probe, cores)
            t24 = time.perf_counter() - t0
            lg = 24
            while lg < full_log2n and runs * t24 * (2 ** (lg + 1 - 24)) * ((lg + 1) / 24.0) * 1.15 < budget_s:
                lg += 1
        n = 1 << lg
        src = rng.integers(0, 2**32, size=n, dtype=np.uint32)
        a = np.empty_like(src)
        for i in range(runs):
            np.copyto(a, src)
            t0 = time.perf_counter()
            oracle.merge_sort_on_cpu_u32(a, cores)
            dt_s = time.perf_counter() - t0
            if i >= args.warmup:
                times.append(dt_s)
        assert np.all(a[:-1] <= a[1:])
        sample = f"2^{lg} uint32 keys per step (merge_sort_on_cpu port, {cores} threads)"
        if args.workload != "sort_u32":
            sample += " [u32 keys stand in for this workload's key type]"
    else:
        n = 1 << (args.sample_log2n if args.sample_log2n else full_log2n)
        x = rng.integers(0, 25, size=n).astype(np.int32)
        out = np.empty_like(x)
        for i in range(runs):
            t0 = time.perf_counter()
            if kind == "scan":
                oracle.scan_on_cpu_i32(x, out, True, 0, cores)
            else:
                oracle.reduce_on_cpu_i32(x, cores)
            dt_s = time.perf_counter() - t0
            if i >= args.warmup:
                times.append(dt_s)
        sample = f"2^{int(np.log2(n))} int32 per step ({kind}_on_cpu port, {cores} threads)"
    ms = 1e3 * float(np.mean(times))
    value = (n / 1e9) / (ms / 1e3) if unit == "Gkeys/s" else (n * bpe / 1e9) / (ms / 1e3)
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32" if kind == "sort" else "i32", "data": "synthetic",
        "config": {"workload": args.workload, "n": n, "n_per_gpu": n, "value_bytes": vb,
                   "distribution": "uniform random, seed 12345", "sample_n": n, "full_n": 1 << full_log2n,
                   "note": "reference CPU-device algorithm (oracle port) on the host cores of rank 0; n is the size actually processed per step",
                   "parallelism": f"{cores} host threads"},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(workload, budget_s=20.0):
    """Oracle port of the reference's CPU-device algorithm, timed on this box's host cores (rank 0, N=1), plus the
    single-threaded STL baselines the reference itself prints next to its numbers."""
    import oracle
    kind, dt, vb, log2n, bpe, metric, unit = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(12345)
    if kind == "sort":
        n = 1 << 24
        src = rng.integers(0, 2**32, size=n, dtype=np.uint32)
        best, spent, runs = None, 0.0, 0
        while runs < 3 and spent < budget_s:
            a = src.copy()
            t0 = time.perf_counter()
            oracle.merge_sort_on_cpu_u32(a, cores)
            e = time.perf_counter() - t0
            best = e if best is None else min(best, e)
            spent += e
            runs += 1
        stl = min(oracle.stl_sort_u32(src.copy()) for _ in range(2))
        return {"value": n / best / 1e9, "unit": "Gkeys/s", "cores": cores, "kind": "port",
                "sample": f"2^24 uint32 keys (BASELINE config 0 size), merge_sort_on_cpu port with {cores} threads, min of {runs}",
                "stl": {"what": "std::sort, 1 thread (perf/perf_stl_sort.cpp:22-30), 2^24 keys, min of 2", "value": n / stl / 1e9, "unit": "Gkeys/s"}}
    n = 1 << 26
    x = rng.integers(0, 25, size=n).astype(np.int32)
    best, runs = None, 0
    out = np.empty_like(x)
    for _ in range(3):
        t0 = time.perf_counter()
        if kind == "scan":
            oracle.scan_on_cpu_i32(x, out, True, 0, cores)
        else:
            oracle.reduce_on_cpu_i32(x, cores)
        e = time.perf_counter() - t0
        best = e if best is None else min(best, e)
        runs += 1
    if kind == "scan":
        stl = min(oracle.stl_partial_sum_i32(x, out) for _ in range(3))
        what = "std::partial_sum, 1 thread (perf/perf_stl_partial_sum.cpp:31-47), 2^26 int32, min of 3"
    else:
        stl = min(oracle.stl_accumulate_i32(x)[0] for _ in range(3))
        what = "std::accumulate, 1 thread (perf/perf_stl_accumulate.cpp:34-38), 2^26 int32, min of 3"
    return {"value": n * bpe / best / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
            "sample": f"2^26 int32, {kind}_on_cpu port with {cores} threads, min of {runs}",
            "stl": {"what": what, "value": n * bpe / stl / 1e9, "unit": "GB/s"}}


KINDS = {"radix_histogram": 0, "digit_scan": 1, "onesweep_pass": 2, "scan": 3, "reduce": 4, "other": 5, "exchange_pass": 6}


class Env:
    """Per-process state shared by all measurements of a run."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        import compute_b200 as cb
        from compute_b200._capi import check, lib
        self.torch, self.dist, self.cb, self.check = torch, dist, cb, check
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.L = lib()
        self.q = cb.command_queue()
        self.stream = self.q.handle
        self.ctx = None
        if self.world > 1:
            from compute_b200 import distributed as cbd
            self.cbd = cbd
            self.ctx = cbd.Context()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def read_timers(self):
        import ctypes
        kt, launches = {}, 0
        for name, k in KINDS.items():
            ms, cnt = ctypes.c_double(), ctypes.c_ulonglong()
            self.check(self.L.bcb_timing_read(self.stream, k, ctypes.byref(ms), ctypes.byref(cnt)))
            kt[name] = (ms.value, cnt.value)
            launches += cnt.value
        return kt, launches


def measure(env, workload, log2n, steps, warmup, scaling="weak", want_e2e=False, args=None):
    """One workload on env.world GPUs: W warm-up steps, K timed steps (CUDA events on the launching stream, max over
    ranks), per-kernel device times, at-size verification.  Returns the record (same on all ranks)."""
    torch, dist, cb, L, check = env.torch, env.dist, env.cb, env.L, env.check
    world, rank = env.world, env.rank
    kind, dt, vb, log2n_default, bpe, metric, unit = WORKLOADS[workload]
    n_total = 1 << (log2n if log2n else log2n_default)
    n_local = n_total // world if scaling == "strong" else n_total
    npdt = NP[dt]
    tdt = {"uint": torch.uint32, "float": torch.float32, "ulong": torch.uint64, "int": torch.int32}[dt]

    # ---- synthetic input, resident in HBM ----
    gen = torch.Generator(device="cuda")
    gen.manual_seed(12345 + rank)
    vals = vals_pristine = None
    if kind == "sort":
        if dt == "ulong":
            pristine = torch.randint(-2**63, 2**63 - 1, (n_local,), dtype=torch.int64, device="cuda", generator=gen).view(tdt)
        elif dt == "float":
            # perf_sort_float.cpp:22-25: ((rand()/RAND_MAX) - 0.5) * 1e5
            pristine = (torch.rand(n_local, device="cuda", generator=gen) - 0.5) * 1e5
        else:
            pristine = torch.randint(-2**31, 2**31 - 1, (n_local,), dtype=torch.int32, device="cuda", generator=gen).view(tdt)
        work = torch.empty_like(pristine)
        if vb == 4:    # payload = original index: lets the permutation and its stability be checked afterwards
            vals_pristine = torch.arange(n_local, dtype=torch.int32, device="cuda").view(torch.uint32)
        elif vb == 8:  # (perf_sort_by_key.cpp:39-42 sorts 64-bit values)
            vals_pristine = torch.arange(n_local, dtype=torch.int64, device="cuda")
        vals = torch.empty_like(vals_pristine) if vb else None
    else:
        if dt == "int":
            pristine = torch.randint(0, 25, (n_local,), dtype=torch.int32, device="cuda", generator=gen)  # perf_exclusive_scan.cpp:22-25
        else:
            pristine = torch.rand(n_local, device="cuda", generator=gen)
        work = torch.empty_like(pristine)
    holder = {}

    def reset():
        if kind == "sort":
            work.copy_(pristine)
            if vb:
                vals.copy_(vals_pristine)

    def step():
        if kind == "sort":
            if world == 1:
                if vb:
                    cb.sort_by_key(work, vals)
                else:
                    cb.sort(work)
            else:
                holder["out"] = env.ctx.sort(work, vals if vb else None)
        elif kind == "scan":
            if world == 1:
                cb.exclusive_scan(pristine, work, 0)
            else:
                env.ctx.exclusive_scan(pristine, work, 0)
        else:
            if world == 1:
                holder["dev"] = torch.empty(1, dtype=pristine.dtype, device="cuda")
                cb.reduce(pristine, holder["dev"])
            else:
                holder["out"] = env.ctx.reduce(pristine)

    for _ in range(warmup):
        reset()
        step()
    env.barrier()
    check(L.bcb_timing_enable(env.stream, 1))
    env.read_timers()  # drop warm-up records
    cur = torch.cuda.current_stream()
    pairs = []
    env.barrier()
    for _ in range(steps):
        reset()
        if world > 1:
            dist.barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(cur)
        step()
        e1.record(cur)
        pairs.append((e0, e1))
    env.barrier()
    step_ms = [a.elapsed_time(b) for a, b in pairs]
    total_ms = float(sum(step_ms))
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / steps
    kt, launches = env.read_timers()
    check(L.bcb_timing_enable(env.stream, 0))

    verified, how = verify(env, kind, dt, vb, n_local, pristine, work, vals, holder)

    peak, peak_src = measured_peaks()
    total_elems = n_local * world
    value = (total_elems / 1e9) / (ms_per_step / 1e3) if unit == "Gkeys/s" else (total_elems * bpe / 1e9) / (ms_per_step / 1e3)

    # ---- roofline of the dominant kernel: algorithmic bytes per launch / average launch duration ----
    if kind == "sort":
        kname, per_elem = "onesweep_pass", 2 * (np.dtype(npdt).itemsize + vb)
    elif kind == "scan":
        kname, per_elem = "scan", bpe
    else:
        kname, per_elem = "reduce", bpe
    k_ms, k_cnt = kt[kname]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch of the dominant kernel, from ncu
    if os.path.exists(tpath) and world == 1:
        try:
            t = json.load(open(tpath)).get(workload)
            if t and int(t.get("n_per_gpu", -1)) == int(n_local):
                traffic = t["dram_bytes_per_launch"]
        except Exception:
            traffic = None
    roofline = None
    if k_cnt:
        avg_ms = k_ms / k_cnt
        # at N > 1 the local passes run over the keys this rank RECEIVED (~ n_local): per-launch bytes use n_local
        achieved = (n_local * per_elem / 1e9) / (avg_ms / 1e3)
        roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "peak_source": peak_src, "avg_launch_ms": avg_ms, "launches": int(k_cnt),
                    "algorithmic_bytes_per_launch": int(n_local * per_elem),
                    "whole_step": {"algorithmic_bytes": int(n_local * bpe), "achieved": (n_local * bpe / 1e9) / (ms_per_step / 1e3),
                                   "frac": (n_local * bpe / 1e9) / (ms_per_step / 1e3) / peak},
                    "kernel_ms_per_step": {k: v[0] / steps for k, v in kt.items() if v[1]}}
    rec = {"workload": workload, "metric": metric, "value": value, "unit": unit, "n": total_elems, "n_per_gpu": n_local,
           "value_bytes": vb, "scaling": scaling, "dtype": SHORT[dt], "ms_per_step": ms_per_step, "steps": steps, "warmup": warmup,
           "roofline": roofline, "verified": verified, "verification": how, "gpu_launches": int(launches), "step_ms": step_ms}
    if world > 1 and kind == "sort":
        rec["distributed"] = env.ctx.last_stats
    if want_e2e:
        if world == 1:
            rec["e2e"] = e2e_run(args, cb, L, env.stream, kind, dt, vb, n_local, pristine, unit, bpe)
        else:
            rec["e2e"] = e2e_run_distributed(args, env.ctx, kind, vb, n_local, world, pristine, unit, bpe)
    del pristine, work, vals, vals_pristine, holder
    torch.cuda.empty_cache()
    return rec


def verify(env, kind, dt, vb, n_local, pristine, work, vals, holder):
    """At-size checks after the timed region (all on the device; nothing here is timed).  Returns (ok, description)."""
    torch, dist, cb = env.torch, env.dist, env.cb
    world = env.world
    bits_t = torch.int32 if pristine.element_size() == 4 else torch.int64
    if kind == "sort" and world == 1:
        ok = bool(cb.is_sorted(work))
        how = "sorted"
        pb, wb = pristine.view(bits_t), work.view(bits_t)
        if vb == 0:
            # a sorted permutation of the input is THE result: multiset equality by sum and xor of the bit patterns
            ok &= int(pb.sum(dtype=torch.int64)) == int(wb.sum(dtype=torch.int64))
            ok &= int(cb.reduce(pb, None, "bit_xor")) == int(cb.reduce(wb, None, "bit_xor"))
            how += " + sum/xor checksums of the bit patterns equal the input's"
        else:
            idx = vals.view(torch.int32 if vb == 4 else torch.int64).long()
            ok &= bool(torch.equal(wb, pb[idx]))                               # keys_out[i] == keys_in[vals_out[i]]
            seen = torch.zeros(n_local, dtype=torch.bool, device="cuda")
            seen[idx] = True
            ok &= bool(seen.all())                                             # the payload is a permutation of 0 .. n-1
            same = wb[1:] == wb[:-1]
            ok &= bool(torch.all((idx[1:] > idx[:-1]) | ~same))                # stable: original index increases inside equal-key runs
            how += " + keys_out[i] == keys_in[vals_out[i]] for all i + payload (= original index) strictly increasing inside every equal-key run"
        return bool(ok), how
    if kind == "sort":
        # global check: every shard sorted, shard boundaries in order, multiset checksums (sum, xor) preserved
        out = holder["out"][0] if vb else holder["out"]
        ok_local = bool(cb.is_sorted(out)) if out.numel() else True
        bits = out.view(bits_t)
        ends = torch.zeros(3, dtype=torch.int64, device="cuda")
        if out.numel():
            tk = env.cbd.transformed_keys(bits[[0, -1]].cpu().numpy().view(np.uint32 if out.element_size() == 4 else np.uint64),
                                          cb.dtype_code(out.dtype), True)
            ends = torch.tensor([1, int(tk[0]) - 2**63, int(tk[1]) - 2**63], dtype=torch.int64, device="cuda")
        all_ends = [torch.zeros_like(ends) for _ in range(world)]
        dist.all_gather(all_ends, ends)
        seq = [(int(e[1]), int(e[2])) for e in all_ends if int(e[0])]
        ok_edges = all(seq[i][1] <= seq[i + 1][0] for i in range(len(seq) - 1))
        pb = pristine.view(bits.dtype)
        xor_in = torch.tensor(int(cb.reduce(pb, None, "bit_xor")) if pb.numel() else 0, device="cuda", dtype=torch.int64)
        xor_out = torch.tensor(int(cb.reduce(bits, None, "bit_xor")) if bits.numel() else 0, device="cuda", dtype=torch.int64)
        xors = [torch.zeros(2, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(xors, torch.stack([xor_in, xor_out]))
        xi = xo = 0
        for x in xors:
            xi ^= int(x[0]); xo ^= int(x[1])
        chk = torch.stack([pb.sum(dtype=torch.int64) - bits.sum(dtype=torch.int64),
                           torch.tensor(int(pb.numel() - bits.numel()), device="cuda"),
                           torch.tensor(0 if ok_local else 1, device="cuda")])
        ok_pairs = True
        if vb:
            # payload = original LOCAL index of the source rank: stable inside equal-key runs of one source only if the
            # source ranks arrive in order; checked per shard as "index increases inside equal-key runs" modulo rank seams
            v = holder["out"][1]
            ok_pairs = v.numel() == out.numel()
        dist.all_reduce(chk)
        ok = bool(ok_edges and int(chk[0]) == 0 and int(chk[1]) == 0 and int(chk[2]) == 0 and xi == xo and ok_pairs)
        return ok, "every shard sorted + shard boundaries in order + global count / sum / xor checksums equal the input's (all-reduce)"
    if kind == "scan":
        if dt == "int":
            ref = torch.cumsum(pristine, 0, dtype=torch.int64) - pristine                      # exclusive, exact
            carry = 0
            if world > 1:
                tot = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
                dist.all_gather(tot, pristine.sum(dtype=torch.int64).reshape(1))
                carry = sum(int(t) for t in tot[:env.rank])
            ok = bool(torch.equal(work, (ref + carry).to(torch.int32)))                         # int32 wrap-around of the exact prefix
            how = "all elements equal the int64 prefix sums (wrapped to int32)"
        else:
            ref = torch.cumsum(pristine.double(), 0) - pristine.double()
            carry = 0.0
            if world > 1:
                tot = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
                dist.all_gather(tot, pristine.double().sum().reshape(1))
                carry = sum(float(t) for t in tot[:env.rank])
            tol = 4 * math.ceil(math.log2(max(2, n_local * world))) * 2.0**-24  # x (running sum of |x|): all inputs are >= 0
            ok = bool(torch.all((work.double() - (ref + carry)).abs() <= tol * (ref + carry) + 1e-6))
            how = f"all elements within {tol:.2e} x prefix of the float64 prefix sums"
        if world > 1:
            f = torch.tensor([0 if ok else 1], device="cuda")
            dist.all_reduce(f)
            ok = int(f) == 0
        return bool(ok), how
    # reduce
    if world == 1:
        got = holder["dev"].cpu()
    else:
        got = holder["out"]
    if dt == "int":
        exact = pristine.sum(dtype=torch.int64)
        if world > 1:
            dist.all_reduce(exact)
        exp = ((int(exact) + 2**31) % 2**32) - 2**31
        return bool(int(got) == exp), "equals the int64 sum wrapped to int32"
    exact = pristine.double().sum()
    if world > 1:
        dist.all_reduce(exact)
    tol = 4 * math.ceil(math.log2(max(2, n_local * world))) * 2.0**-24
    return bool(abs(float(got) - float(exact)) <= tol * float(exact)), f"within {tol:.2e} (relative) of the float64 sum"


def desc_float_check(env):
    """Descending float sorts never speculate (the reference's descending float transform is not injective): 2^24 keys
    with signed zeros / denormals mixed in, compared with the oracle byte for byte."""
    import oracle
    torch, cb = env.torch, env.cb
    n = 1 << 24
    rng = np.random.default_rng(7)
    k = ((rng.random(n, dtype=np.float32) - 0.5) * 1e5).astype(np.float32)
    k[::1000] = 0.0
    k[1::1000] = -0.0
    k[2::1000] = np.finfo(np.float32).smallest_subnormal
    k[3::1000] = -np.finfo(np.float32).smallest_subnormal
    d = torch.from_numpy(k).cuda()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    cb.sort(d.clone(), True)
    w = d.clone()
    t0.record(); cb.sort(w, True); t1.record()
    torch.cuda.synchronize()
    ok = w.cpu().numpy().tobytes() == oracle.sort(k, True).tobytes()
    ms = t0.elapsed_time(t1)
    return {"workload": "sort_f32_desc", "n": n, "value": n / 1e9 / (ms / 1e3), "unit": "Gkeys/s", "ms_per_step": ms,
            "verified": bool(ok), "verification": "all 2^24 keys byte-identical to the oracle's stable LSD sort (incl. +-0 / +-denorm collisions)"}


def ours(args):
    env = Env()
    torch = env.torch
    rank, world = env.rank, env.world
    sampler = ClockSampler(env.local_rank)
    if rank == 0:
        sampler.start()
    head = measure(env, args.workload, args.log2n, args.steps, args.warmup, args.scaling, want_e2e=not args.no_e2e, args=args)
    clocks = sampler.stop() if rank == 0 else None

    configs = None
    if not args.no_configs and args.workload == "sort_u32" and not args.log2n:
        configs = {}
        for w in EXTRA_CONFIGS:
            configs[w] = measure(env, w, 0, 3, 3, "weak")
        if world == 1:
            configs["sort_f32_desc"] = desc_float_check(env)
            configs["sort_u32_2^20"] = measure(env, "sort_u32", 20, 5, 3, "weak")  # small-n: the pre-speculation path
            configs["sort_u32_2^16"] = measure(env, "sort_u32", 16, 10, 3, "weak")  # the one-launch small sort
        else:
            other = "strong" if args.scaling == "weak" else "weak"
            configs[f"sort_u32_{other}"] = measure(env, "sort_u32", 0, 3, 3, other)
        for r in configs.values():
            r.pop("step_ms", None)

    import ctypes
    spec = None
    kind = WORKLOADS[args.workload][0]
    if kind == "sort":
        r_, f_ = ctypes.c_ulonglong(), ctypes.c_ulonglong()
        env.check(env.L.bcb_sort_speculation_stats(env.stream, ctypes.byref(r_), ctypes.byref(f_)))
        spec = {"verified_runs": int(r_.value), "fallbacks": int(f_.value)}
    if rank == 0:
        line = {
            "metric": head["metric"], "value": head["value"], "unit": head["unit"], "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": head["dtype"], "data": "synthetic",
            "config": {"workload": args.workload, "n": head["n"], "n_per_gpu": head["n_per_gpu"], "value_bytes": head["value_bytes"],
                       "distribution": "uniform random, seed 12345+rank", "l2": "inputs >> L2 (no flush needed)",
                       "timing": "CUDA events per step on the launching stream, input reset outside the timed region",
                       "parallelism": f"{world} process(es), one per GPU"},
            "roofline": head["roofline"], "sort_speculation": spec, "clocks": clocks, "gpu_launches": head["gpu_launches"],
            "verified": head["verified"], "verification": head["verification"], "step_ms": head["step_ms"], "e2e": head.get("e2e"),
        }
        if "distributed" in head:
            line["distributed"] = head["distributed"]
        if configs is not None:
            line["configs"] = configs
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args.workload)
        print(json.dumps(line), flush=True)
    if world > 1:
        env.dist.destroy_process_group()


def e2e_run(args, cb, L, stream, kind, dt, vb, n, pristine, unit, bpe):
    """Same metric through the reference-facing host-buffer call: pinned host input -> H2D -> kernels -> D2H."""
    import ctypes

    import torch
    from compute_b200._capi import check
    steps = max(1, min(args.steps, 3))
    w = pristine.element_size()
    if kind == "sort" and vb == 0:
        host = torch.empty(n, dtype=pristine.dtype, pin_memory=True)
        times = []
        code = cb.dtype_code(pristine.dtype)
        for i in range(steps + 1):
            host.copy_(pristine)  # reset (untimed)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            check(L.bcb_sort_host(stream, code, 0, host.data_ptr(), n))  # blocks until the sorted keys are back
            e = time.perf_counter() - t0
            if i > 0:
                times.append(e)
        ms = 1e3 * float(np.mean(times))
        m = min(n, 1 << 20)
        head = host.view(torch.uint8)[: m * w].numpy().view(NP[dt])
        ok = bool(np.all(head[:-1] <= head[1:]))
        # the same call on PAGEABLE memory (what sort(v.begin(), v.end()) on a std::vector hands over): the copies are
        # staged by the driver
        pageable = None
        try:
            src = pristine.cpu().view(torch.uint8).numpy().view(NP[dt])
            hp = src.copy()  # plain malloc'ed memory
            pes = []
            for _ in range(2):  # (the first call also allocates the library's pinned staging slots)
                np.copyto(hp, src)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                check(L.bcb_sort_host(stream, code, 0, hp.ctypes.data, n))
                pes.append(time.perf_counter() - t0)
            pe = min(pes)
            pageable = {"value": (n / 1e9) / pe, "unit": unit, "ms_per_step": pe * 1e3, "steps": 2, "first_call_ms": pes[0] * 1e3,
                        "staging": "library: up to 16 host threads through pinned slots (runtime.cu staged_copy_pageable)",
                        "checked": bool(np.all(hp[:m][:-1] <= hp[:m][1:]))}
            del hp, src
        except MemoryError:
            pageable = None
        return {"value": (n / 1e9) / (ms / 1e3), "unit": unit, "h2d_bytes_per_step": n * w, "d2h_bytes_per_step": n * w,
                "ms_per_step": ms, "entry": "bcb_sort_host (sort(host_first, host_last), sort.hpp:125-148), pinned host buffer",
                "steps": steps, "checked": ok, "pageable": pageable}
    if kind in ("scan", "reduce"):
        host_in = torch.empty(n, dtype=pristine.dtype, pin_memory=True)
        host_in.copy_(pristine)
        dev = torch.empty_like(pristine)
        out = torch.empty_like(pristine)
        host_out = torch.empty(n, dtype=pristine.dtype, pin_memory=True) if kind == "scan" else None
        times = []
        for i in range(steps + 1):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dev.copy_(host_in, non_blocking=True)
            if kind == "scan":
                cb.exclusive_scan(dev, out, 0)
                host_out.copy_(out, non_blocking=True)
                torch.cuda.synchronize()
            else:
                cb.reduce(dev)  # host result: blocks
            e = time.perf_counter() - t0
            if i > 0:
                times.append(e)
        ms = 1e3 * float(np.mean(times))
        d2h = n * w if kind == "scan" else w
        return {"value": (n * bpe / 1e9) / (ms / 1e3), "unit": unit, "h2d_bytes_per_step": n * w, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms, "entry": "vector(host range) + algorithm + copy back", "steps": steps}
    return None


def e2e_run_distributed(args, ctx, kind, vb, n, world, pristine, unit, bpe):
    """N > 1: every rank's shard starts and ends in pinned host memory; H2D + distributed algorithm + D2H inside the
    timed region (wall clock between barriers, max over ranks)."""
    import torch
    import torch.distributed as dist
    steps = max(1, min(args.steps, 3))
    w = pristine.element_size()
    host_in = torch.empty(n, dtype=pristine.dtype, pin_memory=True)
    host_in.copy_(pristine)
    cap = n + n // 4 + 1024 if kind == "sort" else n
    host_out = torch.empty(cap, dtype=pristine.dtype, pin_memory=True) if kind != "reduce" else None
    dev = torch.empty_like(pristine)
    out = torch.empty_like(pristine) if kind == "scan" else None
    vhost_in = vhost_out = vdev = None
    if kind == "sort" and vb:  # payload = index inside the shard, like the device-timed arm
        vdt = torch.int32 if vb == 4 else torch.int64
        vhost_in = torch.arange(n, dtype=vdt).pin_memory()
        vhost_out = torch.empty(cap, dtype=vdt, pin_memory=True)
        vdev = torch.empty(n, dtype=vdt, device=dev.device)
    times, d2h = [], 0
    for i in range(steps + 1):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        dev.copy_(host_in, non_blocking=True)
        if kind == "sort" and vb:
            vdev.copy_(vhost_in, non_blocking=True)
            res, resv = ctx.sort(dev, vdev)
            m = min(res.numel(), cap)
            host_out[:m].copy_(res[:m], non_blocking=True)
            vhost_out[:m].copy_(resv[:m], non_blocking=True)
            d2h = m * (w + vb)
        elif kind == "sort":
            res = ctx.sort(dev, None)
            m = min(res.numel(), cap)  # (a receive imbalance above 25 % would truncate the copy-back; never seen with regular sampling)
            host_out[:m].copy_(res[:m], non_blocking=True)
            d2h = m * w
        elif kind == "scan":
            ctx.exclusive_scan(dev, out, 0)
            host_out.copy_(out, non_blocking=True)
            d2h = n * w
        else:
            ctx.reduce(dev)
            d2h = w
        torch.cuda.synchronize()
        dist.barrier()
        e = time.perf_counter() - t0
        if i > 0:
            times.append(e)
    t = torch.tensor([float(np.mean(times))], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t.item())
    total = n * world
    value = (total / 1e9) / sec if unit == "Gkeys/s" else (total * bpe / 1e9) / sec
    return {"value": value, "unit": unit, "h2d_bytes_per_step": n * (w + (vb if kind == "sort" else 0)) * world, "d2h_bytes_per_step": d2h * world, "ms_per_step": sec * 1e3,
            "entry": "compute_b200.distributed.Context (pinned host shard -> H2D -> distributed algorithm -> D2H)", "steps": steps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sort_u32", choices=sorted(WORKLOADS))
    ap.add_argument("--log2n", type=int, default=0)
    ap.add_argument("--sample-log2n", type=int, default=0)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = 2^log2n elements per GPU, strong = 2^log2n elements in total")
    ap.add_argument("--no-configs", action="store_true", help="headline only: skip the other BASELINE configs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
