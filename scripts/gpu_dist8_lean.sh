#!/bin/bash
# usage: gpu_dist8_lean.sh N -- lean multi-GPU evidence run: parity of the default plan, then the headline sort weak / strong
# with the digit-exchange plan and weak with the partition pass + local sort
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 tests/dist_check_worker.py > gpurun_out/dist_check_N${N}.log 2>&1
grep -E "MISMATCH|DIST_CHECK|rror|sort n=" gpurun_out/dist_check_N${N}.log | cut -c1-160 | tail -8
for cfg in "1 sort_u32 weak" "1 sort_u32 strong" "0 sort_u32 weak" "1 sort_pairs_u32 strong"; do
  set -- $cfg
  BCB_DIST_DIGIT_EXCHANGE=$1 timeout 300 $TR --master-port 29522 bench.py --gpus $N --steps 4 --warmup 3 --no-configs --no-e2e --no-cpu --workload $2 --scaling $3 > gpurun_out/lean_N${N}_dx$1_$2_$3.json 2> gpurun_out/lean.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/lean_N${N}_dx$1_$2_$3.json').read().strip().splitlines()[-1])
    print('dx=$1 $2 $3', round(d['value'],2), d['unit'], round(d['ms_per_step'],3), d['verified'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d['distributed']['plan'], d['distributed'].get('imbalance'))
except Exception as e:
    print('no json', e); print(open('gpurun_out/lean.err').read()[-1500:])
PY
done
