"""The digit-exchange plan of the multi-GPU sort with all ranks emulated on ONE GPU (the placement of
compute_b200.distributed.digit_exchange_plan, every "peer" buffer local): times bcb_radix_exchange_scatter and
bcb_radix_sort_segments with CUDA events and checks the concatenated result.  Also the target of the ncu captures of the
segment kernels (scripts/gpu_digit_exchange_evidence.sh).  usage: digit_exchange_emulation.py [world] [log2 keys per rank]"""
import ctypes, os, sys
import numpy as np
import torch
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import compute_b200 as cb
from compute_b200 import distributed as cbd
from compute_b200.core import dtype_code

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 28)
lib, q, code = cb.lib(), cb.command_queue(), dtype_code(torch.uint32)
g = torch.Generator(device="cuda"); g.manual_seed(3)
shards = [torch.randint(-2**31, 2**31 - 1, (n,), device="cuda", generator=g, dtype=torch.int32).view(torch.uint32) for _ in range(world)]
allh = []
for s in shards:
    c = np.zeros(256, dtype=np.uint64)
    cb._capi.check(lib.bcb_radix_top_histogram(q.handle, code, 1, s.data_ptr(), n, c.ctypes.data))
    allh.append(c.astype(np.int64))
owner, first, seg_begin, seg_len, recv, span, imb = cbd.digit_exchange_plan(np.stack(allh), world)
recv_bufs = [torch.empty(int(span[d]), dtype=torch.int32, device="cuda") for d in range(world)]
outs = [torch.empty(int(recv[d]), dtype=torch.int32, device="cuda") for d in range(world)]
pk = np.array([recv_bufs[int(owner[x])].data_ptr() for x in range(256)], dtype=np.uint64)

def ev():
    return torch.cuda.Event(enable_timing=True)

for it in range(3):
    t = [ev() for _ in range(3)]
    t[0].record()
    for r in range(world):
        df = np.ascontiguousarray(first[r], dtype=np.uint64)
        cb._capi.check(lib.bcb_radix_exchange_scatter(q.handle, code, 1, shards[r].data_ptr(), None, 0, n, pk.ctypes.data, None, df.ctypes.data))
    t[1].record()
    for d in range(world):
        mine = owner == d
        sb, sl = np.ascontiguousarray(seg_begin[mine], dtype=np.uint64), np.ascontiguousarray(seg_len[mine], dtype=np.uint64)
        cb._capi.check(lib.bcb_radix_sort_segments(q.handle, code, 1, recv_bufs[d].data_ptr(), None, 0, outs[d].data_ptr(), None,
                                                   sb.ctypes.data, sl.ctypes.data, sb.size))
    t[2].record()
    torch.cuda.synchronize()
    print(f"world {world}, 2^{int(np.log2(n))} keys per rank: exchange pass {t[0].elapsed_time(t[1]) / world:.3f} ms per rank (local destinations), "
          f"segment sort {t[1].elapsed_time(t[2]) / world:.3f} ms per rank ({int(mine.sum())} segments)", flush=True)
full = torch.cat(outs).view(torch.uint32)
a = full.view(torch.int32).to(torch.int64) & 0xffffffff
ok = bool((a[1:] >= a[:-1]).all()) and full.numel() == world * n
ref = torch.cat(shards).view(torch.int32).to(torch.int64) & 0xffffffff
ok = ok and int(a.sum()) == int(ref.sum())
print("sorted + checksum:", "OK" if ok else "MISMATCH")
