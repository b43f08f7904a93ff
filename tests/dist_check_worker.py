#!/usr/bin/env python
"""Multi-GPU parity check (worker of tests/test_gpu_distributed.py), run under torchrun (one rank per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/dist_check_worker.py
Every distributed result must equal the single-GPU result bit for bit (rank 0 recomputes on one GPU and, for small
sizes, against the CPU oracle)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import compute_b200 as cb
from compute_b200 import distributed as cbd


def gather_var(t, world, rank):
    """gather variable-length 1-D/2-D tensors to rank 0 (bitwise)."""
    n = torch.tensor([t.shape[0]], device="cuda", dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    row = t.element_size() * (t.numel() // t.shape[0] if t.shape[0] else 1)
    mx = max(sizes)
    buf = torch.zeros(mx * row, dtype=torch.uint8, device="cuda")
    buf[: t.shape[0] * row] = t.contiguous().view(torch.uint8).reshape(-1)
    outs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    if rank != 0:
        return None
    return torch.cat([o[: s * row] for o, s in zip(outs, sizes)])


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ctx = cbd.Context()
    ok = True
    # (the first case gives every rank >= 2^23 keys: the local sorts take the warp-specialised pass kernel)
    for n_total, dt, vb, desc in (((1 << 23) * world + 12345, torch.uint32, 0, False), (1 << 22, torch.uint32, 0, False), (1 << 21, torch.float32, 0, True), (1 << 21, torch.int32, 8, False),
                                  (100_003, torch.uint64, 4, True), (1 << 20, torch.int16, 0, False)):
        g = torch.Generator(device="cuda"); g.manual_seed(7)
        if dt == torch.float32:
            full = (torch.rand(n_total, device="cuda", generator=g) - 0.5) * 100
            full[::1000] = -0.0
        elif dt == torch.uint64:
            full = torch.randint(-2**63, 2**63 - 1, (n_total,), device="cuda", generator=g, dtype=torch.int64).view(dt)
        elif dt == torch.int16:
            full = torch.randint(-2**15, 2**15 - 1, (n_total,), device="cuda", generator=g, dtype=torch.int16)
        else:
            full = torch.randint(-1000 if vb else -2**31, 1000 if vb else 2**31 - 1, (n_total,), device="cuda", generator=g, dtype=torch.int32).view(dt)
        vals_full = None
        if vb:
            vals_full = torch.arange(n_total * (vb // 4), device="cuda", dtype=torch.int32).reshape(n_total, vb // 4).contiguous()
        lo, hi = rank * n_total // world, (rank + 1) * n_total // world
        if n_total == 1 << 22:  # uneven blocks: rank 0 holds NOTHING, the last rank the rest
            cuts = [0, 0] + [n_total * r // (2 * world) for r in range(2, world)] + [n_total]
            lo, hi = cuts[rank], cuts[rank + 1]
        shard = full[lo:hi].clone()
        vshard = vals_full[lo:hi].clone() if vb else None
        res = ctx.sort(shard, vshard, descending=desc)
        out_k = res[0] if vb else res
        gk = gather_var(out_k, world, rank)
        gv = gather_var(res[1], world, rank) if vb else None
        if rank == 0:
            ref_k = full.clone(); ref_v = vals_full.clone() if vb else None
            if vb: cb.stable_sort_by_key(ref_k, ref_v, desc)
            else: cb.stable_sort(ref_k, desc)
            torch.cuda.synchronize()
            same = torch.equal(gk, ref_k.view(torch.uint8).reshape(-1)) and (not vb or torch.equal(gv, ref_v.view(torch.uint8).reshape(-1)))
            print(f"sort n={n_total} {dt} vb={vb} desc={desc}: {'OK' if same else 'MISMATCH'} stats={ctx.last_stats}", flush=True)
            ok &= bool(same)
    # scans / reductions: rank r holds block r
    for dt in (torch.int32, torch.float32, torch.int64):
        n_total = 3_000_001
        g = torch.Generator(device="cuda"); g.manual_seed(11)
        full = torch.randint(-2**31, 2**31 - 1, (n_total,), device="cuda", generator=g, dtype=torch.int32).to(dt) if dt != torch.float32 \
            else torch.rand(n_total, device="cuda", generator=g)
        lo, hi = rank * n_total // world, (rank + 1) * n_total // world
        mine = full[lo:hi].clone(); out = torch.empty_like(mine)
        for mode in ("excl", "incl"):
            if mode == "excl": ctx.exclusive_scan(mine, out, 5)
            else: ctx.inclusive_scan(mine, out)
            go = gather_var(out, world, rank)
            if rank == 0:
                ref = torch.empty_like(full)
                if mode == "excl": cb.exclusive_scan(full, ref, 5)
                else: cb.inclusive_scan(full, ref)
                torch.cuda.synchronize()
                got = go.view(dt)
                same = torch.equal(got, ref) if dt != torch.float32 else bool(torch.allclose(got, ref, rtol=1e-5, atol=1e-2))
                print(f"{mode} scan {dt}: {'OK' if same else 'MISMATCH'}", flush=True)
                ok &= bool(same)
        s = ctx.reduce(mine); s1 = cb.reduce(full) if rank == 0 else None
        if rank == 0:
            same = (s == s1) if dt != torch.float32 else abs(float(s) - float(s1)) <= 1e-5 * abs(float(s1))
            print(f"reduce {dt}: {'OK' if same else 'MISMATCH'} {s} {s1}", flush=True)
            ok &= bool(same)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_CHECK", "PASS" if ok else "FAIL", flush=True)
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
