#!/usr/bin/env python
"""bench.py -- headline benchmark of the sort / scan / reduce path (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--log2n L]

One "step" = one pass of the hot path over one batch of synthetic input.  The default workload is the one
BASELINE.json's metric is quoted on: compute::sort of 2^30 uniform-random uint32 keys (Gkeys/s).  Inputs are
reset to the same random keys before every step outside the timed region (mirrors perf/perf_sort.cpp:36-40);
each step is timed on the device with CUDA events on the launching stream; inputs (4 GiB) are far larger than
L2, so no explicit flush is needed.  Rank 0 prints ONE JSON line.

Other workloads (parity-tested configs, selectable for profiling): sort_u64, sort_f32, sort_pairs_u32,
scan_i32, scan_f32, reduce_i32, reduce_f32.  `--impl reference` times the reference's own CPU-device algorithm
(merge_sort_on_cpu restated in oracle/, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
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
    Bounded sample per step so the run ends within minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    kind, dt, vb, log2n, bpe, metric, unit = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    n = 1 << (args.sample_log2n if args.sample_log2n else (24 if kind == "sort" else 26))
    rng = np.random.default_rng(12345)
    times = []
    if kind == "sort":
        src = rng.integers(0, 2**32, size=n, dtype=np.uint32)
        for i in range(args.warmup + args.steps):
            a = src.copy()
            t0 = time.perf_counter()
            oracle.merge_sort_on_cpu_u32(a, cores)
            dt_s = time.perf_counter() - t0
            if i >= args.warmup:
                times.append(dt_s)
        assert np.all(a[:-1] <= a[1:])
        sample = f"2^{int(np.log2(n))} uint32 keys per step (merge_sort_on_cpu port, {cores} threads)"
        if args.workload != "sort_u32":
            sample += " [u32 keys stand in for this workload's key type]"
    elif kind == "scan":
        x = rng.integers(0, 25, size=n).astype(np.int32)
        out = np.empty_like(x)
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            oracle.scan_on_cpu_i32(x, out, True, 0, cores)
            dt_s = time.perf_counter() - t0
            if i >= args.warmup:
                times.append(dt_s)
        sample = f"2^{int(np.log2(n))} int32 per step (scan_on_cpu port, {cores} threads)"
    else:
        x = rng.integers(0, 25, size=n).astype(np.int32)
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            oracle.reduce_on_cpu_i32(x, cores)
            dt_s = time.perf_counter() - t0
            if i >= args.warmup:
                times.append(dt_s)
        sample = f"2^{int(np.log2(n))} int32 per step (reduce_on_cpu port, {cores} threads)"
    ms = 1e3 * float(np.mean(times))
    value = (n / 1e9) / (ms / 1e3) if unit == "Gkeys/s" else (n * bpe / 1e9) / (ms / 1e3)
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32" if kind == "sort" else "i32", "data": "synthetic",
        "config": {"workload": args.workload, "n": (1 << (args.log2n if args.log2n else log2n)) * max(1, args.gpus),
                   "n_per_gpu": 1 << (args.log2n if args.log2n else log2n), "value_bytes": vb,
                   "distribution": "uniform random, seed 12345", "sample_n": n,
                   "note": "reference CPU-device algorithm (oracle port) on the host cores of rank 0, bounded sample per step",
                   "parallelism": f"{cores} host threads"},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(workload, budget_s=20.0):
    """Oracle port of the reference's CPU-device algorithm, timed on this box's host cores (rank 0, N=1)."""
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
        return {"value": n / best / 1e9, "unit": "Gkeys/s", "cores": cores, "kind": "port",
                "sample": f"2^24 uint32 keys (BASELINE config 0 size), merge_sort_on_cpu port with {cores} threads, min of {runs}"}
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
    return {"value": n * bpe / best / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
            "sample": f"2^26 int32, {kind}_on_cpu port with {cores} threads, min of {runs}"}


def ours(args):
    import torch
    import torch.distributed as dist

    import compute_b200 as cb
    from compute_b200._capi import check, lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    kind, dt, vb, log2n_default, bpe, metric, unit = WORKLOADS[args.workload]
    log2n = args.log2n if args.log2n else log2n_default
    n_total = 1 << log2n
    L = lib()
    q = cb.command_queue()
    stream = q.handle
    npdt = NP[dt]
    tdt = {"uint": torch.uint32, "float": torch.float32, "ulong": torch.uint64, "int": torch.int32}[dt]

    # ---- synthetic input, resident in HBM ----
    gen = torch.Generator(device="cuda")
    gen.manual_seed(12345 + rank)
    if kind == "sort":
        scaling = "weak"  # every rank holds one block of 2^log2n keys of the global range (N x 2^30 keys in total)
        n_local = n_total
        if dt == "ulong":
            pristine = torch.randint(-2**63, 2**63 - 1, (n_local,), dtype=torch.int64, device="cuda", generator=gen).view(tdt)
        elif dt == "float":
            # perf_sort_float.cpp:22-25: ((rand()/RAND_MAX) - 0.5) * 1e5
            pristine = (torch.rand(n_local, device="cuda", generator=gen) - 0.5) * 1e5
        else:
            pristine = torch.randint(-2**31, 2**31 - 1, (n_local,), dtype=torch.int32, device="cuda", generator=gen).view(tdt)
        work = torch.empty_like(pristine)
        vals_pristine = None
        if vb == 4:
            vals_pristine = torch.arange(n_local, dtype=torch.int32, device="cuda").view(torch.uint32)
        elif vb == 8:
            vals_pristine = torch.arange(n_local, dtype=torch.int64, device="cuda")
        vals = torch.empty_like(vals_pristine) if vb else None
    else:
        scaling = "weak"  # every rank scans / reduces its own 2^28 block; carries are P scalars
        n_local = n_total
        if dt == "int":
            pristine = torch.randint(0, 25, (n_local,), dtype=torch.int32, device="cuda", generator=gen)  # perf_exclusive_scan.cpp:22-25
        else:
            pristine = torch.rand(n_local, device="cuda", generator=gen)
        work = torch.empty_like(pristine)

    if world > 1:
        from compute_b200 import distributed as cbd
        ctx = cbd.Context()

    def reset():
        if kind == "sort":
            work.copy_(pristine)
            if vb:
                vals.copy_(vals_pristine)

    result_holder = {}

    def step():
        if kind == "sort":
            if world == 1:
                if vb:
                    cb.sort_by_key(work, vals)
                else:
                    cb.sort(work)
            else:
                result_holder["out"] = ctx.sort(work, vals if vb else None)
        elif kind == "scan":
            if world == 1:
                cb.exclusive_scan(pristine, work, 0)
            else:
                ctx.exclusive_scan(pristine, work, 0)
        else:
            if world == 1:
                rd = torch.empty(1, dtype=pristine.dtype, device="cuda")
                cb.reduce(pristine, rd)
            else:
                result_holder["out"] = ctx.reduce(pristine)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        reset()
        step()
    barrier()
    check(L.bcb_timing_enable(stream, 1))
    for k in range(6):
        L.bcb_timing_read(stream, k, None, None)  # drop warm-up records
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    cur = torch.cuda.current_stream()
    pairs = []
    barrier()
    for _ in range(args.steps):
        reset()
        if world > 1:
            dist.barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(cur)
        step()
        e1.record(cur)
        pairs.append((e0, e1))
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [a.elapsed_time(b) for a, b in pairs]
    total_ms = float(sum(step_ms))
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps

    # per-kernel device time of the timed region (event pairs recorded by the launchers on this stream)
    kinds = {"radix_histogram": 0, "digit_scan": 1, "onesweep_pass": 2, "scan": 3, "reduce": 4, "other": 5}
    kt = {}
    launches = 0
    import ctypes
    for name, k in kinds.items():
        ms = ctypes.c_double()
        cnt = ctypes.c_ulonglong()
        check(L.bcb_timing_read(stream, k, ctypes.byref(ms), ctypes.byref(cnt)))
        kt[name] = (ms.value, cnt.value)
        launches += cnt.value
    check(L.bcb_timing_enable(stream, 0))

    # ---- verification (outside the timed region): sortedness + checksum, or oracle on a prefix ----
    verified = None
    if kind == "sort" and world == 1:
        ok_sorted = cb.is_sorted(work)
        if dt in ("uint",):
            s0 = cb.reduce(pristine.view(torch.uint32), None, "plus")
            s1 = cb.reduce(work.view(torch.uint32), None, "plus")
            x0 = cb.reduce(pristine.view(torch.uint32), None, "bit_xor")
            x1 = cb.reduce(work.view(torch.uint32), None, "bit_xor")
            verified = bool(ok_sorted and s0 == s1 and x0 == x1)
        else:
            verified = bool(ok_sorted)
    elif kind == "sort":
        # global check: every shard sorted, shard boundaries in order, multiset checksums (sum, xor) preserved
        out = result_holder["out"][0] if vb else result_holder["out"]
        ok_local = bool(cb.is_sorted(out)) if out.numel() else True
        bits = out.view(torch.int32 if out.element_size() == 4 else torch.int64)
        ends = torch.zeros(3, dtype=torch.int64, device="cuda")
        if out.numel():
            tk = cbd.transformed_keys(bits[[0, -1]].cpu().numpy().view(np.uint32 if out.element_size() == 4 else np.uint64),
                                      cb.dtype_code(out.dtype), True)
            ends = torch.tensor([1, int(tk[0]) - 2**63, int(tk[1]) - 2**63], dtype=torch.int64, device="cuda")
        all_ends = [torch.zeros_like(ends) for _ in range(world)]
        dist.all_gather(all_ends, ends)
        seq = [(int(e[1]), int(e[2])) for e in all_ends if int(e[0])]
        ok_edges = all(seq[i][1] <= seq[i + 1][0] for i in range(len(seq) - 1))
        pb = pristine.view(bits.dtype)
        chk = torch.stack([pb.sum(dtype=torch.int64) - bits.sum(dtype=torch.int64),
                           torch.tensor(int(pb.numel() - bits.numel()), device="cuda"),
                           torch.tensor(0 if ok_local else 1, device="cuda")])
        dist.all_reduce(chk)
        verified = bool(ok_edges and int(chk[0]) == 0 and int(chk[1]) == 0 and int(chk[2]) == 0)
    elif kind == "scan" and world == 1:
        m = 1 << 20
        import oracle
        exp = oracle.scan(pristine[:m].cpu().numpy(), "plus", True, 0)
        got = work[:m].cpu().numpy()
        verified = bool(np.array_equal(got, exp)) if dt == "int" else bool(np.allclose(got, exp, rtol=1e-4))

    peak, peak_src = measured_peaks()
    if unit == "Gkeys/s":
        value = (n_local * world / 1e9) / (ms_per_step / 1e3)
    else:
        value = (n_local * world * bpe / 1e9) / (ms_per_step / 1e3)

    # ---- roofline of the dominant kernel ----
    if kind == "sort":
        kname, per_elem = "onesweep_pass", 2 * (np.dtype(npdt).itemsize + vb)
    elif kind == "scan":
        kname, per_elem = "scan", bpe
    else:
        kname, per_elem = "reduce", bpe
    k_ms, k_cnt = kt[kname]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch of the dominant kernel, from ncu
    if os.path.exists(tpath):
        try:
            t = json.load(open(tpath)).get(args.workload)
            if t and int(t.get("n_per_gpu", -1)) == int(n_local):
                traffic = t["dram_bytes_per_launch"]
        except Exception:
            traffic = None
    roofline = None
    if k_cnt:
        avg_ms = k_ms / k_cnt
        achieved = (n_local * per_elem / 1e9) / (avg_ms / 1e3)
        roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "peak_source": peak_src, "avg_launch_ms": avg_ms, "launches": int(k_cnt),
                    "algorithmic_bytes_per_launch": int(n_local * per_elem),
                    "whole_step": {"algorithmic_bytes": int(n_local * bpe), "achieved": (n_local * bpe / 1e9) / (ms_per_step / 1e3),
                                   "frac": (n_local * bpe / 1e9) / (ms_per_step / 1e3) / peak},
                    "kernel_ms_per_step": {k: v[0] / args.steps for k, v in kt.items() if v[1]}}

    # ---- e2e: the same metric through the host-buffer entry point (H2D + sort + D2H inside the timed region) ----
    e2e = None
    if world == 1 and not args.no_e2e:
        e2e = e2e_run(args, cb, L, stream, kind, dt, vb, n_local, pristine, unit, bpe)
    elif world > 1 and not args.no_e2e:
        e2e = e2e_run_distributed(args, ctx, kind, vb, n_local, world, pristine, unit, bpe)

    spec = None
    if kind == "sort":
        r_, f_ = ctypes.c_ulonglong(), ctypes.c_ulonglong()
        check(L.bcb_sort_speculation_stats(stream, ctypes.byref(r_), ctypes.byref(f_)))
        spec = {"verified_runs": int(r_.value), "fallbacks": int(f_.value)}
    if rank == 0:
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": {"uint": "u32", "float": "f32", "ulong": "u64", "int": "i32"}[dt], "data": "synthetic",
            "config": {"workload": args.workload, "n": n_local * world, "n_per_gpu": n_local, "value_bytes": vb,
                       "distribution": "uniform random, seed 12345+rank", "l2": "inputs >> L2 (no flush needed)",
                       "timing": "CUDA events per step on the launching stream, input reset outside the timed region",
                       "parallelism": f"{world} process(es), one per GPU"},
            "roofline": roofline, "sort_speculation": spec, "clocks": clocks, "gpu_launches": int(launches), "verified": verified,
            "step_ms": step_ms, "e2e": e2e,
        }
        if world > 1 and kind == "sort":
            line["distributed"] = ctx.last_stats
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args.workload)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
        return {"value": (n / 1e9) / (ms / 1e3), "unit": unit, "h2d_bytes_per_step": n * w, "d2h_bytes_per_step": n * w,
                "ms_per_step": ms, "entry": "bcb_sort_host (sort(host_first, host_last), sort.hpp:125-148)", "steps": steps,
                "checked": ok}
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
    if kind == "sort" and vb:
        return None
    times, d2h = [], 0
    for i in range(steps + 1):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        dev.copy_(host_in, non_blocking=True)
        if kind == "sort":
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
    return {"value": value, "unit": unit, "h2d_bytes_per_step": n * w * world, "d2h_bytes_per_step": d2h * world, "ms_per_step": sec * 1e3,
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
