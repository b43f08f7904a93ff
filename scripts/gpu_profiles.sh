#!/bin/bash
# Evidence for profiles/: launch list of the default bench command, DRAM traffic of the dominant kernels at full size.
mkdir -p gpurun_out
echo "== launch list (default bench, 2 steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_sort_u32.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/launches_sort_u32.log 2>&1; tail -1 gpurun_out/launches_sort_u32.log | cut -c1-200
echo "== dram traffic per launch"
for w in sort_u32 scan_i32 reduce_i32; do
  case $w in sort_u32) k=onesweep_pass;; scan_i32) k=scan_tma;; reduce_i32) k=reduce_kernel;; esac
  if [ $w = sort_u32 ]; then skip=4; else skip=1; fi
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:$k -s $skip -c 1 --csv --log-file gpurun_out/traffic_$w.csv python bench.py --workload $w --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/traffic_$w.log 2>&1
  tail -4 gpurun_out/traffic_$w.csv | cut -c1-300
done
echo "== full ncu of final kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:onesweep_pass -s 5 -c 1 -f -o gpurun_out/prof_sort_final python bench.py --log2n 28 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_sort_final.log 2>&1; tail -1 gpurun_out/ncu_sort_final.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_tma -s 1 -c 1 -f -o gpurun_out/prof_scan_final python bench.py --workload scan_i32 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_scan_final.log 2>&1; tail -1 gpurun_out/ncu_scan_final.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:reduce_kernel -s 1 -c 1 -f -o gpurun_out/prof_reduce_final python bench.py --workload reduce_i32 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_reduce_final.log 2>&1; tail -1 gpurun_out/ncu_reduce_final.log
