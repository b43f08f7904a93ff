#!/bin/bash
# Round-end evidence on one GPU: default bench (both arms), other workloads, launch list, DRAM traffic, full ncu captures.
mkdir -p gpurun_out
echo "== default bench"; timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.json; echo
echo "== reference arm"; timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 700 gpurun_out/bench_reference.json; echo
for w in sort_f32 sort_u64 sort_pairs_u32 scan_i32 scan_f32 reduce_i32 reduce_f32; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --workload $w > gpurun_out/bench_$w.json 2>> gpurun_out/bench_other.err
  python -c "
import json;d=json.load(open('gpurun_out/bench_$w.json'));print('$w',round(d['value'],2),d['unit'],round(d['ms_per_step'],3),d['roofline'] and round(d['roofline']['frac'],3),d['verified'],d['e2e'] and round(d['e2e']['value'],2))"
done
echo "== deterministic-only sort"; BCB_SORT_SPECULATIVE=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_sort_u32_det.json 2>> gpurun_out/bench_other.err; python -c "
import json;d=json.load(open('gpurun_out/bench_sort_u32_det.json'));print(d['value'],d['ms_per_step'],d['roofline']['kernel_ms_per_step'])"
echo "== launch list (default bench, 2 steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_sort_u32.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/launches_sort_u32.log 2>&1; tail -1 gpurun_out/launches_sort_u32.log | cut -c1-200
echo "== dram traffic per launch"
for w in sort_u32 scan_i32 reduce_i32; do
  case $w in sort_u32) k=onesweep_pass; skip=4;; scan_i32) k=scan_tma; skip=1;; reduce_i32) k=reduce_kernel; skip=1;; esac
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:$k -s $skip -c 1 --csv --log-file gpurun_out/traffic_$w.csv python bench.py --workload $w --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/traffic_$w.log 2>&1
  tail -3 gpurun_out/traffic_$w.csv | cut -c1-60,200-330
done
