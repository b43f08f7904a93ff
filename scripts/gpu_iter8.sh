#!/bin/bash
bash scripts/gpu_iter7.sh
echo "== dram traffic per launch (scan, reduce)"
for w in scan_i32 reduce_i32; do
  case $w in scan_i32) k=scan_tma;; reduce_i32) k=reduce_kernel;; esac
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:$k -s 1 -c 1 --csv --log-file gpurun_out/traffic_$w.csv python bench.py --workload $w --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/traffic_$w.log 2>&1
  tail -4 gpurun_out/traffic_$w.csv | cut -c1-300
done
for w in scan_i32 scan_f32 reduce_i32 reduce_f32; do
  echo "== $w" ; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload $w > gpurun_out/bench_$w.json 2>> gpurun_out/bench_other.err ; python -c "
import json;d=json.load(open('gpurun_out/bench_$w.json'));print(d['value'],d['unit'],d['ms_per_step'],d['roofline'] and d['roofline']['frac'],d['verified'])"
done
