#!/bin/bash
mkdir -p gpurun_out
echo "== pytest" ; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -4 gpurun_out/pytest_gpu.log
for w in scan_i32 scan_f32; do
  echo "== $w (tma)" ; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload $w > gpurun_out/bench_$w.json 2>> gpurun_out/bench_other.err ; python -c "
import json;d=json.load(open('gpurun_out/bench_$w.json'));print(d['value'],d['unit'],d['ms_per_step'],d['roofline'] and d['roofline']['frac'],d['verified'])"
done
for v in 0; do
  echo "== variant $v" ; BCB_SORT_VARIANT=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_sort_u32_v$v.json 2>> gpurun_out/bench_variants.err ; python -c "
import json;d=json.load(open('gpurun_out/bench_sort_u32_v$v.json'));print(d['value'],d['ms_per_step'],d['roofline']['kernel_ms_per_step'],d['verified'])"
done
for v in 0; do
  echo "== ordered variant $v" ; BCB_SORT_RANK=ordered BCB_SORT_VARIANT=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_sort_u32_ord_v$v.json 2>> gpurun_out/bench_variants.err ; python -c "
import json;d=json.load(open('gpurun_out/bench_sort_u32_ord_v$v.json'));print(d['value'],d['ms_per_step'],d['roofline']['kernel_ms_per_step'],d['verified'])"
done
echo "== ordered-atomics parity (u32 keys)"; BCB_SORT_RANK=ordered timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "radix_sort_keys_bit_exact and uint or unaligned or repeat_calls or host_range" > gpurun_out/pytest_ordered.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_ordered.log
for w in sort_pairs_u32 sort_u64; do
  echo "== $w" ; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --workload $w > gpurun_out/bench_$w.json 2>> gpurun_out/bench_other.err ; python -c "
import json;d=json.load(open('gpurun_out/bench_$w.json'));print(d['value'],d['unit'],d['ms_per_step'],d['roofline'] and d['roofline']['frac'],d['verified'])"
done
echo "== ncu scan" ; timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_tma -s 1 -c 1 -f -o gpurun_out/prof_scan3 python bench.py --workload scan_i32 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_scan.log 2>&1 ; tail -2 gpurun_out/ncu_scan.log
tail -n 5 gpurun_out/*.err
