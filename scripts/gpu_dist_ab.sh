#!/bin/bash
# usage: gpu_dist_ab.sh N -- the multi-GPU sort on N GPUs: parity first (every plan), then A/B of the digit-exchange plan
# (exchange = the pass over the most significant digit) against the partition pass + local sort
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for dx in 1 0; do
  echo "== parity, BCB_DIST_DIGIT_EXCHANGE=$dx"
  BCB_DIST_DIGIT_EXCHANGE=$dx timeout 600 $TR --master-port 29521 tests/dist_check_worker.py > gpurun_out/dist_check_N${N}_dx$dx.log 2>&1
  grep -E "MISMATCH|DIST_CHECK|rror|sort n=" gpurun_out/dist_check_N${N}_dx$dx.log | cut -c1-220 | tail -12
done
for dx in 1 0; do
  for w in sort_u32 sort_pairs_u32 sort_u64 sort_f32; do
    for sc in weak strong; do
      BCB_DIST_DIGIT_EXCHANGE=$dx timeout 600 $TR --master-port 29522 bench.py --gpus $N --steps 5 --warmup 3 --no-configs --no-e2e --no-cpu --workload $w --scaling $sc > gpurun_out/ab_N${N}_dx${dx}_${w}_$sc.json 2> gpurun_out/ab.err
      python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab_N${N}_dx${dx}_${w}_$sc.json').read().strip().splitlines()[-1])
    print('dx=$dx $w $sc', round(d['value'],2), d['unit'], round(d['ms_per_step'],3), d['verified'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d['distributed']['plan'], d['distributed'].get('imbalance'))
except Exception as e:
    print('no json', e); print(open('gpurun_out/ab.err').read()[-1500:])
PY
    done
  done
done
