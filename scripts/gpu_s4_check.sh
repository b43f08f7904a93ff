#!/bin/bash
# session 4, 1 GPU: parity of the digit-exchange kernels emulated on one GPU, sort regression tests, headline A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "digit_exchange or sort or concurrent or enqueue" > gpurun_out/s4_pytest.log 2>&1
tail -15 gpurun_out/s4_pytest.log
for w in sort_u32 sort_pairs_u32 sort_f32; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-configs --no-e2e --no-cpu --workload $w > gpurun_out/s4_$w.json 2> gpurun_out/s4_bench.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/s4_$w.json').read().strip().splitlines()[-1])
    print('$w', round(d['value'],2), d['unit'], round(d['ms_per_step'],3), d['verified'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()})
except Exception as e:
    print('no json', e); print(open('gpurun_out/s4_bench.err').read()[-1500:])
PY
done
