#!/bin/bash
# constant-digit passes as streaming copies: parity, then small-integer keys at several sizes (both kernel families)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hot_digit or speculative or 2_26 or keys_bit_exact or pairs_bit_exact or golden" > gpurun_out/s4_pytest_const.log 2>&1; tail -4 gpurun_out/s4_pytest_const.log
for hot in 1 0; do
BCB_SORT_HOT=$hot python - <<'PY'
import os, torch
import compute_b200 as cb
hot = os.environ["BCB_SORT_HOT"]
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
    return min(ts[2:]), work
for log2n in (22, 26, 28):
    n = 1 << log2n
    vals = torch.arange(n, device="cuda", dtype=torch.int32).view(torch.uint32)
    for name, gen in (("uniform u32", lambda: torch.randint(-2**31, 2**31-1, (n,), device="cuda", dtype=torch.int32).view(torch.uint32)),
                      ("u32 < 2^16", lambda: torch.randint(0, 65536, (n,), device="cuda", dtype=torch.int32).view(torch.uint32)),
                      ("i32 in +-1000", lambda: torch.randint(-1000, 1000, (n,), device="cuda", dtype=torch.int32)),
                      ("u64 < 2^16", lambda: torch.randint(0, 65536, (n,), device="cuda", dtype=torch.int64).view(torch.uint64))):
        src = gen()
        tk, out = t_sort(lambda w, v: cb.sort(w), src)
        ref = torch.sort(src.view(torch.int64) if src.dtype == torch.uint64 else src.view(torch.int32) if "i32" in name else src.view(torch.int32).to(torch.int64) & 0xffffffff).values
        got = out.view(torch.int64) if out.dtype == torch.uint64 else out.view(torch.int32) if "i32" in name else out.view(torch.int32).to(torch.int64) & 0xffffffff
        ok = bool(torch.equal(ref, got))
        tp, _ = t_sort(lambda w, v: cb.sort_by_key(w, v), src, vals)
        print(f"hot={hot} 2^{log2n} {name}: keys {tk:.3f} ms = {n / tk / 1e6:.1f} Gkeys/s {'OK' if ok else 'MISMATCH'}; with u32 payload {tp:.3f} ms = {n / tp / 1e6:.1f} Gkeys/s", flush=True)
PY
done
