#!/usr/bin/env python
"""Single-GPU cost of the splitter partition pass (all destinations local): counts + scatter for 1, 3 and 7 splitters."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import compute_b200 as cb
from compute_b200._capi import check, lib

n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 30
L = lib()
q = cb.command_queue()
g = torch.Generator(device="cuda"); g.manual_seed(1)
keys = torch.randint(-2**31, 2**31 - 1, (n,), dtype=torch.int32, device="cuda", generator=g).view(torch.uint32)
out = torch.empty(n + 64, dtype=torch.uint32, device="cuda")
cur = torch.cuda.current_stream()
for ns in (1, 3, 7):
    sp = np.array([(j + 1) * (1 << 32) // (ns + 1) for j in range(ns)], dtype=np.uint64)
    counts = np.zeros(ns + 1, dtype=np.uint64)
    check(L.bcb_partition_counts(q.handle, 5, 1, keys.data_ptr(), n, sp.ctypes.data, ns, counts.ctypes.data))
    offs = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
    pk = (ctypes.c_void_p * (ns + 1))(*[out.data_ptr() + int(o) * 4 for o in offs])
    for what in ("counts", "scatter"):
        ts = []
        for it in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(cur)
            if what == "counts":
                check(L.bcb_partition_counts(q.handle, 5, 1, keys.data_ptr(), n, sp.ctypes.data, ns, counts.ctypes.data))
            else:
                check(L.bcb_partition_scatter(q.handle, 5, 1, keys.data_ptr(), None, 0, n, sp.ctypes.data, ns, pk, None))
            e1.record(cur)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"nsplit={ns} {what:8s} {min(ts[1:]):7.3f} ms  ({n * (4 if what == 'counts' else 8) / 1e6 / min(ts[1:]):7.1f} GB/s)", flush=True)
