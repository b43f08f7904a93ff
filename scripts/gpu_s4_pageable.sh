#!/bin/bash
# staged pageable copies: parity, then the e2e leg of the bench (pinned and pageable) with the staging on and off
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sort_host" 2>&1 | tail -3
for st in 1 0; do
  BCB_STAGED_COPY=$st timeout 600 python bench.py --steps 3 --warmup 3 --no-configs --no-cpu > gpurun_out/s4_pageable_$st.json 2> gpurun_out/s4_bench.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/s4_pageable_$st.json').read().strip().splitlines()[-1])
    e=d['e2e']; print('staged=$st e2e pinned', round(e['value'],2), e['unit'], round(e['ms_per_step'],1), 'ms; pageable', e.get('pageable'))
except Exception as ex:
    print('no json', ex); print(open('gpurun_out/s4_bench.err').read()[-1500:])
PY
done
nproc; free -g | head -2
