#!/bin/bash
# Round-2 evidence on one GPU: ncu launch list of the bench command, DRAM traffic per launch of the dominant kernels,
# full ncu captures of the pass / scan / reduce kernels condensed by scripts/ncu_summary.py.
mkdir -p gpurun_out
B="--no-e2e --no-cpu --no-configs"
echo "== launch list (default bench, 2 steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_sort_u32.csv python bench.py --steps 2 --warmup 1 $B > gpurun_out/launches_sort_u32.log 2>&1; tail -1 gpurun_out/launches_sort_u32.log | cut -c1-200
echo "== dram traffic per launch"
for w in sort_u32 scan_i32 reduce_i32; do
  case $w in sort_u32) k=onesweep_ws; skip=4;; scan_i32) k=scan_ws; skip=1;; reduce_i32) k=reduce_tma; skip=1;; esac
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:$k -s $skip -c 1 --csv --log-file gpurun_out/traffic_$w.csv python bench.py --workload $w --steps 1 --warmup 1 $B > gpurun_out/traffic_$w.log 2>&1
  tail -3 gpurun_out/traffic_$w.csv | cut -c1-60,200-330
done
python scripts/make_traffic_json.py > /dev/null && cp profiles/traffic.json gpurun_out/traffic.json
echo "== full captures"
cap() { # name kernel-regex skip workload log2n
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/prof_$1 python bench.py --workload $4 --log2n $5 --steps 1 --warmup 1 $B > gpurun_out/ncu_$1.log 2>&1
  python scripts/ncu_summary.py gpurun_out/prof_$1.ncu-rep gpurun_out/r02_$1.txt "$6" > /dev/null 2>&1; head -30 gpurun_out/r02_$1.txt
}
for what in "$@"; do
case $what in
ws)     cap sort_pass_ws onesweep_ws 4 sort_u32 28 "onesweep_ws<u32> (keys-only, speculative), 2^28 keys";;
wsdet)  cap sort_pass_ws_pairs onesweep_ws 4 sort_pairs_u32 28 "onesweep_ws<u32,4,DET> (u32 + u32 payload), 2^28 pairs";;
scan)   cap scan_tma scan_tma 1 scan_i32 28 "scan_tma_kernel<int>, 2^28";;
reduce) cap reduce_tma reduce_tma 1 reduce_i32 28 "reduce_tma_kernel<int>, 2^28";;
esac
done
rm -f gpurun_out/prof_reduce_tma.ncu-rep
ls -la gpurun_out/
