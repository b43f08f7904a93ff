#!/usr/bin/env python
"""Emulates the 8-GPU exchange pass on 2 GPUs: every rank partitions its shard with 7 splitters and stores 7 of the 8
buckets into the OTHER rank's HBM (1 bucket stays local), i.e. the per-GPU NVLink egress/ingress of an 8-rank
all-to-all at a quarter of the GPU cost.  Prints the scatter time per rank.  torchrun, 2 ranks."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import compute_b200 as cb
from compute_b200 import distributed as cbd
from compute_b200._capi import check, lib


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    remote = int(sys.argv[2]) if len(sys.argv) > 2 else 7   # how many of the 8 buckets go to the peer
    n = 1 << log2n
    ctx = cbd.Context()
    assert ctx.peer.ensure(2 * n * 4 + 4096), "peer memory unavailable"
    L = lib(); q = cb.command_queue(); cur = torch.cuda.current_stream()
    g = torch.Generator(device="cuda"); g.manual_seed(5 + rank)
    keys = torch.randint(-2**31, 2**31 - 1, (n,), dtype=torch.int32, device="cuda", generator=g).view(torch.uint32)
    ns = 7
    sp = np.array([(j + 1) * (1 << 32) // (ns + 1) for j in range(ns)], dtype=np.uint64)
    counts = np.zeros(ns + 1, dtype=np.uint64)
    check(L.bcb_partition_counts(q.handle, 5, 1, keys.data_ptr(), n, sp.ctypes.data, ns, counts.ctypes.data))
    offs = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
    peer = ctx.peer.peers[(rank + 1) % world]
    local = ctx.peer.local
    # buckets 0..remote-1 -> first half of the peer's buffer, the rest -> second half of my own buffer
    dst = [(peer if b < remote else local + n * 4) + int(offs[b]) * 4 for b in range(ns + 1)]
    pk = (ctypes.c_void_p * (ns + 1))(*dst)
    ts = []
    for it in range(5):
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cur)
        check(L.bcb_partition_scatter(q.handle, 5, 1, keys.data_ptr(), None, 0, n, sp.ctypes.data, ns, pk, None))
        e1.record(cur)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    dist.barrier()
    rbytes = float(counts[:remote].sum()) * 4
    print(f"rank {rank}: scatter n=2^{log2n} remote_buckets={remote}: {min(ts[1:]):.3f} ms  (remote {rbytes / 1e9:.2f} GB -> "
          f"{rbytes / 1e6 / min(ts[1:]):.0f} GB/s)", flush=True)
    ctx.peer.release()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
