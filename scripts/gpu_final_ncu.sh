#!/bin/bash
# full ncu captures of the kernels around the pass kernel (reports are ~20 MB each: at most two per gpurun call)
# usage: gpu_final_ncu.sh hist|verify|split|scan|reduce ...
mkdir -p gpurun_out
for what in "$@"; do
case $what in
hist)   timeout 600 ncu --set full --clock-control none --import-source on -k regex:radix_histogram -s 1 -c 1 -f -o gpurun_out/prof_hist_columns python bench.py --log2n 28 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_hist.log 2>&1; tail -1 gpurun_out/ncu_hist.log;;
verify) timeout 600 ncu --set full --clock-control none --import-source on -k regex:verify_sorted -s 1 -c 1 -f -o gpurun_out/prof_verify python bench.py --log2n 28 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_verify.log 2>&1; tail -1 gpurun_out/ncu_verify.log;;
split)  timeout 600 ncu --set full --clock-control none --import-source on -k regex:onesweep_pass -s 8 -c 1 -f -o gpurun_out/prof_split_pass python scripts/partition_timing.py 28 > gpurun_out/ncu_split.log 2>&1; tail -1 gpurun_out/ncu_split.log;;
esac
done
ls -la gpurun_out/*.ncu-rep
