#!/usr/bin/env python
"""NVLink store bandwidth into a peer's HBM, one process per GPU (torchrun): copy engines vs SM-issued stores of 4 / 8 / 16
bytes per thread, every rank writing to rank (r+1) % P at the same time.  Prints GB/s per GPU."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

import compute_b200 as cb
from compute_b200 import distributed as cbd
from compute_b200._capi import check, lib


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ctx = cbd.Context()
    nbytes = 1 << 31
    assert ctx.peer.ensure(nbytes), "peer memory unavailable"
    src = torch.randint(0, 255, (nbytes,), dtype=torch.uint8, device="cuda")
    L = lib()
    q = cb.command_queue()
    cur = torch.cuda.current_stream()
    results = {}
    for target, tname in (((rank + 1) % world, "peer"), (rank, "local")):
        dst = ctx.peer.peers[target]
        # (the SM-issued 4/8/16-byte store rows of round 1 -- 696-706 GB/s -- came from a diagnostic copy kernel that no
        # longer ships in the product library)
        for name, fn in (("copy-engine", lambda: check(L.bcb_memcpy_d2d(q.handle, dst, src.data_ptr(), nbytes))),):
            for _ in range(2):
                fn()
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(cur)
            for _ in range(5):
                fn()
            e1.record(cur)
            torch.cuda.synchronize(); dist.barrier()
            results[f"{tname}/{name}"] = 5 * nbytes / 1e9 / (e0.elapsed_time(e1) / 1e3)
    if rank == 0:
        for k, v in results.items():
            print(f"{k:24s} {v:8.1f} GB/s", flush=True)
    ctx.peer.release()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
