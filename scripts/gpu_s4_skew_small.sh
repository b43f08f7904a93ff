#!/bin/bash
# how the r01 pass kernels (sorts below 2^27 keys, 64-bit keys) take constant digits: uniform against small-integer keys
python - <<'PY'
import torch
import compute_b200 as cb
def t_sort(fn, src, vals=None):
    work = torch.empty_like(src); wv = torch.empty_like(vals) if vals is not None else None
    ts = []
    for it in range(6):
        work.copy_(src)
        if vals is not None: wv.copy_(vals)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(work, wv); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts[2:])
for log2n in (22, 25, 26):
    n = 1 << log2n
    vals = torch.arange(n, device="cuda", dtype=torch.int32).view(torch.uint32)
    for name, gen in (("uniform u32", lambda: torch.randint(-2**31, 2**31-1, (n,), device="cuda", dtype=torch.int32).view(torch.uint32)),
                      ("u32 < 2^16", lambda: torch.randint(0, 65536, (n,), device="cuda", dtype=torch.int32).view(torch.uint32)),
                      ("uniform u64", lambda: torch.randint(-2**63, 2**63-1, (n,), device="cuda", dtype=torch.int64).view(torch.uint64)),
                      ("u64 < 2^16", lambda: torch.randint(0, 65536, (n,), device="cuda", dtype=torch.int64).view(torch.uint64))):
        src = gen()
        tk = t_sort(lambda w, v: cb.sort(w), src)
        tp = t_sort(lambda w, v: cb.sort_by_key(w, v), src, vals)
        print(f"2^{log2n} {name}: keys {tk:.3f} ms = {n / tk / 1e6:.1f} Gkeys/s; with u32 payload {tp:.3f} ms = {n / tp / 1e6:.1f} Gkeys/s", flush=True)
PY
