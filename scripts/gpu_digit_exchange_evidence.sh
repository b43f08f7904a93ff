#!/bin/bash
# ncu evidence for the kernels of the digit-exchange plan (all ranks emulated on one GPU, 2 x 2^28 keys)
mkdir -p gpurun_out
python scripts/digit_exchange_emulation.py 2 28 2>&1 | tail -4
cap() { # name kernel-regex skip title
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/prof_$1 python scripts/digit_exchange_emulation.py 2 28 > gpurun_out/ncu_$1.log 2>&1
  python scripts/ncu_summary.py gpurun_out/prof_$1.ncu-rep gpurun_out/r02_$1.txt "$4" > /dev/null 2>&1; head -32 gpurun_out/r02_$1.txt
}
cap exchange_pass_digit onesweep_ws 0 "onesweep_ws<u32> with a destination table: the pass over the top digit (exchange pass, local destinations), 2^28 keys"
cap segment_pass onesweep_ws 2 "onesweep_ws<u32> with a tile table: one LSD pass over all 128 segments of a rank (2^28 keys)"
cap segment_histogram segment_histogram 0 "segment_histogram<u32>: digits 0-2 of 128 segments, 2^28 keys"
rm -f gpurun_out/prof_segment_histogram.ncu-rep gpurun_out/prof_exchange_pass_digit.ncu-rep
