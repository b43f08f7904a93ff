#!/bin/bash
# First GPU pass: smoke, parity tests, headline bench + variants.  Everything under `timeout`.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -3 gpurun_out/smoke.log
echo "== pytest" ; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -15 gpurun_out/pytest_gpu.log
echo "== bench headline" ; timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_sort_u32.json 2> gpurun_out/bench_sort_u32.err ; echo "rc=$?" ; cat gpurun_out/bench_sort_u32.json | cut -c1-1500
for v in 1 2 3 4; do
  echo "== variant $v" ; BCB_SORT_VARIANT=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_sort_u32_v$v.json 2>> gpurun_out/bench_variants.err ; python -c "
import json;d=json.load(open('gpurun_out/bench_sort_u32_v$v.json'));print(d['value'],d['ms_per_step'],d['roofline']['kernel_ms_per_step'],d['verified'])"
done
for w in scan_i32 scan_f32 reduce_i32 reduce_f32 sort_pairs_u32 sort_u64 sort_f32; do
  echo "== $w" ; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --workload $w > gpurun_out/bench_$w.json 2>> gpurun_out/bench_other.err ; python -c "
import json;d=json.load(open('gpurun_out/bench_$w.json'));print(d['value'],d['unit'],d['ms_per_step'],d['roofline'] and d['roofline']['frac'],d['verified'],d['e2e'] and d['e2e']['value'])"
done
tail -5 gpurun_out/*.err
