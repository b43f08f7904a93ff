#!/bin/bash
mkdir -p gpurun_out
echo "== pytest scan/reduce" ; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scan or reduce or accumulate or golden" > gpurun_out/pytest_scan.log 2>&1 ; echo "rc=$?" ; tail -3 gpurun_out/pytest_scan.log
for w in scan_i32 scan_f32; do
  echo "== $w" ; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload $w > gpurun_out/bench_$w.json 2>> gpurun_out/bench_other.err ; python -c "
import json;d=json.load(open('gpurun_out/bench_$w.json'));print(d['value'],d['unit'],d['ms_per_step'],d['roofline'] and d['roofline']['frac'],d['verified'])"
done
for v in 0 1 2 3 4 5 6; do
  echo "== reduce_i32 variant $v" ; BCB_REDUCE_VARIANT=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --workload reduce_i32 2>> gpurun_out/bench_other.err | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['value'],d['unit'],d['ms_per_step'],d['roofline'] and d['roofline']['frac'], min(d['step_ms']))"
done
echo "== ncu scan" ; timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_tma -s 1 -c 1 -f -o gpurun_out/prof_scan4 python bench.py --workload scan_i32 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_scan.log 2>&1 ; tail -2 gpurun_out/ncu_scan.log
tail -n 5 gpurun_out/bench_other.err
