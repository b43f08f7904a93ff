#!/bin/bash
# usage: gpu_dist.sh N   -- multi-GPU parity + bench on N GPUs of one box
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== dist_check N=$N (peer-memory plan)"
timeout 600 $TR --master-port 29511 tests/dist_check_worker.py > gpurun_out/dist_check_$N.log 2>&1; echo "rc=$?"; grep -E "OK|MISMATCH|DIST_CHECK|Error|error" gpurun_out/dist_check_$N.log | cut -c1-260 | tail -20
echo "== dist_check N=$N (NCCL all-to-all plan)"
BCB_DIST_PEER=0 timeout 600 $TR --master-port 29512 tests/dist_check_worker.py > gpurun_out/dist_check_nccl_$N.log 2>&1; echo "rc=$?"; grep -E "DIST_CHECK|Error|error" gpurun_out/dist_check_nccl_$N.log | cut -c1-200 | tail -5
for mode in peer nccl; do
  if [ $mode = nccl ]; then export BCB_DIST_PEER=0; else export BCB_DIST_PEER=1; fi
  echo "== bench sort_u32 N=$N plan=$mode"
  timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_sort_u32_N${N}_$mode.json 2> gpurun_out/bench_N${N}_$mode.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_sort_u32_N${N}_$mode.json').read().strip().splitlines()[-1])
    print(d['value'], d['unit'], d['ms_per_step'], d['verified'], d.get('distributed'), d['step_ms'])
except Exception as e:
    print('no json', e); print(open('gpurun_out/bench_N${N}_$mode.err').read()[-1500:])
PY
  echo "== phases (profiled, synchronised) plan=$mode"
  BCB_DIST_PROFILE=1 timeout 600 $TR --master-port 29514 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_prof_N${N}_$mode.json 2>> gpurun_out/bench_N${N}_$mode.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_prof_N${N}_$mode.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], d.get('distributed'))
except Exception as e:
    print('no json', e)
PY
done
unset BCB_DIST_PEER
for w in sort_pairs_u32 scan_i32 reduce_i32; do
  echo "== bench $w N=$N"
  timeout 600 $TR --master-port 29515 bench.py --gpus $N --steps 5 --warmup 3 --workload $w > gpurun_out/bench_${w}_N$N.json 2> gpurun_out/bench_${w}_N$N.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${w}_N$N.json').read().strip().splitlines()[-1])
    print(d['value'], d['unit'], d['ms_per_step'], d['verified'], d.get('distributed'))
except Exception as e:
    print('no json', e); print(open('gpurun_out/bench_${w}_N$N.err').read()[-1500:])
PY
done
