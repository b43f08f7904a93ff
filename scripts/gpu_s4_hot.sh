#!/bin/bash
# hot-digit ballot ranking: parity, then A/B (BCB_SORT_HOT) on keys-only and key-value sorts of small integers
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hot_digit or speculative or 2_26 or pairs" > gpurun_out/s4_pytest_hot.log 2>&1; tail -5 gpurun_out/s4_pytest_hot.log
for hot in 1 0; do
BCB_SORT_HOT=$hot python - <<'PY'
import os, torch
import compute_b200 as cb
n = 1 << 28
hot = os.environ["BCB_SORT_HOT"]
def t_sort(fn, src, vals=None):
    work = torch.empty_like(src); wv = torch.empty_like(vals) if vals is not None else None
    ts = []
    for it in range(5):
        work.copy_(src)
        if vals is not None: wv.copy_(vals)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(work, wv); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts[2:])
vals = torch.arange(n, device="cuda", dtype=torch.int32).view(torch.uint32)
for name, gen in (("uniform u32", lambda: torch.randint(-2**31, 2**31-1, (n,), device="cuda", dtype=torch.int32).view(torch.uint32)),
                  ("u32 < 2^16", lambda: torch.randint(0, 65536, (n,), device="cuda", dtype=torch.int32).view(torch.uint32)),
                  ("u32 < 1000", lambda: torch.randint(0, 1000, (n,), device="cuda", dtype=torch.int32).view(torch.uint32))):
    src = gen()
    tk = t_sort(lambda w, v: cb.sort(w), src)
    tp = t_sort(lambda w, v: cb.sort_by_key(w, v), src, vals)
    print(f"hot={hot} {name}: 2^28 keys {tk:.3f} ms = {n / tk / 1e6:.1f} Gkeys/s; with u32 payload {tp:.3f} ms = {n / tp / 1e6:.1f} Gkeys/s", flush=True)
PY
done
