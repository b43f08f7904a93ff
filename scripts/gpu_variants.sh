#!/bin/bash
# sweep tile shapes of the speculative two-sweep pass: VARIANTS="0 1 2" WORKLOAD=sort_u32 bash scripts/gpu_variants.sh
for v in ${VARIANTS:-0 1 2 3 4 5 6}; do
  BCB_SORT_VARIANT=$v timeout 120 python bench.py --workload ${WORKLOAD:-sort_u32} --steps 2 --warmup 2 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read());k=d['roofline']['kernel_ms_per_step'];print('${WORKLOAD:-sort_u32} variant $v', round(d['value'],2), 'pass_ms', round(k['onesweep_pass'],3), d['verified'], d['sort_speculation'])"
done
