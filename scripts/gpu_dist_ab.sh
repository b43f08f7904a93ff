#!/bin/bash
# usage: gpu_dist_ab.sh N -- A/B of the multi-GPU sort's exchange on N GPUs: splitters from the top-digit histogram or
# from samples, LSU exchange kernel or the (opt-in) warp-specialised one; parity first
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for ws in 0 1; do
  echo "== parity, BCB_SPLIT_WS=$ws"
  BCB_SPLIT_WS=$ws BCB_SPLIT_WS_MIN_LOG2=20 timeout 600 $TR --master-port 29521 tests/dist_check_worker.py 2>&1 | grep -E "MISMATCH|DIST_CHECK|rror" | cut -c1-200 | tail -4
done
for cfg in "1 0" "0 0" "1 1"; do
  set -- $cfg
  for w in sort_u32 sort_pairs_u32 sort_u64; do
    BCB_DIST_HISTOGRAM=$1 BCB_SPLIT_WS=$2 timeout 600 $TR --master-port 29522 bench.py --gpus $N --steps 5 --warmup 3 --no-configs --no-e2e --workload $w > gpurun_out/ab_N${N}_h$1_ws$2_$w.json 2> gpurun_out/ab.err
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab_N${N}_h$1_ws$2_$w.json').read().strip().splitlines()[-1])
    print('histogram=$1 ws=$2 $w', round(d['value'],2), d['unit'], round(d['ms_per_step'],3), d['verified'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d['distributed']['splitters'])
except Exception as e:
    print('no json', e); print(open('gpurun_out/ab.err').read()[-1500:])
PY
  done
done
