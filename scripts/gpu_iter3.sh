#!/bin/bash
mkdir -p gpurun_out
echo "== pytest" ; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -4 gpurun_out/pytest_gpu.log
for v in ${VARIANTS:-0 1 2 3 4 5 6 7 8}; do
  echo "== variant $v" ; BCB_SORT_VARIANT=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_sort_u32_v$v.json 2>> gpurun_out/bench_variants.err ; python -c "
import json;d=json.load(open('gpurun_out/bench_sort_u32_v$v.json'));print(d['value'],d['ms_per_step'],d['roofline']['kernel_ms_per_step'],d['verified'])"
done
for v in ${ORD_VARIANTS:-0 1 5}; do
  echo "== ordered variant $v" ; BCB_SORT_RANK=ordered BCB_SORT_VARIANT=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_sort_u32_ord_v$v.json 2>> gpurun_out/bench_variants.err ; python -c "
import json;d=json.load(open('gpurun_out/bench_sort_u32_ord_v$v.json'));print(d['value'],d['ms_per_step'],d['roofline']['kernel_ms_per_step'],d['verified'])"
done
for w in ${WORKLOADS:-sort_pairs_u32 sort_u64}; do
  echo "== $w" ; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --workload $w > gpurun_out/bench_$w.json 2>> gpurun_out/bench_other.err ; python -c "
import json;d=json.load(open('gpurun_out/bench_$w.json'));print(d['value'],d['unit'],d['ms_per_step'],d['roofline'] and d['roofline']['frac'],d['verified'])"
done
if [ -n "$NCU_VARIANT" ]; then
echo "== ncu sort" ; BCB_SORT_VARIANT=$NCU_VARIANT timeout 600 ncu --set full --clock-control none --import-source on -k regex:onesweep_pass -s 5 -c 1 -f -o gpurun_out/prof_sort3 python bench.py --log2n 28 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_sort.log 2>&1 ; tail -2 gpurun_out/ncu_sort.log
fi
tail -n 5 gpurun_out/*.err
