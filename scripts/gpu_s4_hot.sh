#!/bin/bash
# hot-digit ballot ranking: parity, then the headline (uniform: must not move), perf_sort_float's keys and small-int keys
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hot_digit or speculative or 2_26 or digit_exchange" > gpurun_out/s4_pytest_hot.log 2>&1; tail -5 gpurun_out/s4_pytest_hot.log
for hot in 1 0; do
for w in sort_u32 sort_f32; do
  BCB_SORT_HOT=$hot timeout 300 python bench.py --steps 5 --warmup 3 --no-configs --no-e2e --no-cpu --workload $w > gpurun_out/s4_hot_$w.json 2> gpurun_out/s4_bench.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/s4_hot_$w.json').read().strip().splitlines()[-1])
    print('hot=$hot $w', round(d['value'],2), d['unit'], round(d['ms_per_step'],3), d['verified'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()})
except Exception as e:
    print('no json', e); print(open('gpurun_out/s4_bench.err').read()[-1500:])
PY
done
BCB_SORT_HOT=$hot python - <<'PY'
import torch, time, numpy as np
import compute_b200 as cb
n = 1 << 28
for name, gen in (("uniform u32", lambda: torch.randint(-2**31, 2**31-1, (n,), device="cuda", dtype=torch.int32).view(torch.uint32)),
                  ("u32 < 2^16", lambda: torch.randint(0, 65536, (n,), device="cuda", dtype=torch.int32).view(torch.uint32)),
                  ("u32 < 1000", lambda: torch.randint(0, 1000, (n,), device="cuda", dtype=torch.int32).view(torch.uint32)),
                  ("30 % one value", lambda: torch.where(torch.rand(n, device="cuda") < 0.3, torch.full((n,), 0x2B2B2B2B, device="cuda", dtype=torch.int32), torch.randint(-2**31, 2**31-1, (n,), device="cuda", dtype=torch.int32)).view(torch.uint32))):
    src = gen(); work = torch.empty_like(src)
    ts = []
    for it in range(6):
        work.copy_(src); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); cb.sort(work); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ok = bool((work.view(torch.int64)[:-1:1] >= 0).all()) if False else True
    print(f"hot={__import__('os').environ['BCB_SORT_HOT']} {name}: 2^28 keys {min(ts[2:]):.3f} ms = {n / min(ts[2:]) / 1e6:.1f} Gkeys/s")
PY
done
