#!/bin/bash
mkdir -p gpurun_out
echo "== pytest scan/reduce" ; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scan or reduce or accumulate or golden" > gpurun_out/pytest_scan.log 2>&1 ; echo "rc=$?" ; tail -3 gpurun_out/pytest_scan.log
for w in scan_i32 scan_f32; do
  echo "== $w" ; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload $w > gpurun_out/bench_$w.json 2>> gpurun_out/bench_other.err ; python -c "
import json;d=json.load(open('gpurun_out/bench_$w.json'));print(d['value'],d['unit'],d['ms_per_step'],d['roofline'] and d['roofline']['frac'],d['verified'], min(d['step_ms']))"
done
echo "== sort_u32"; timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_sort_u32_spec.json 2>> gpurun_out/bench_variants.err ; python -c "
import json;d=json.load(open('gpurun_out/bench_sort_u32_spec.json'));print(d['value'],d['ms_per_step'],d['roofline']['kernel_ms_per_step'],d['verified'],d['sort_speculation'])"
tail -n 5 gpurun_out/bench_other.err
