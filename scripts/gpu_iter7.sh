#!/bin/bash
mkdir -p gpurun_out
echo "== pytest" ; timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -12 gpurun_out/pytest_gpu.log
echo "== sort_u32 default (speculative)"; timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_sort_u32_spec.json 2>> gpurun_out/bench_variants.err ; python -c "
import json;d=json.load(open('gpurun_out/bench_sort_u32_spec.json'));print(d['value'],d['ms_per_step'],d['roofline']['kernel_ms_per_step'],d['verified'])"
echo "== sort_u32 deterministic only"; BCB_SORT_SPECULATIVE=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_sort_u32_det.json 2>> gpurun_out/bench_variants.err ; python -c "
import json;d=json.load(open('gpurun_out/bench_sort_u32_det.json'));print(d['value'],d['ms_per_step'],d['roofline']['kernel_ms_per_step'],d['verified'])"
for w in sort_u64 sort_f32 sort_pairs_u32; do
  echo "== $w" ; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --workload $w > gpurun_out/bench_$w.json 2>> gpurun_out/bench_other.err ; python -c "
import json;d=json.load(open('gpurun_out/bench_$w.json'));print(d['value'],d['unit'],d['ms_per_step'],d['roofline'] and d['roofline']['frac'],d['verified'])"
done
tail -n 5 gpurun_out/*.err
