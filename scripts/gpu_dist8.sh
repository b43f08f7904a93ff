#!/bin/bash
# 8-GPU box: parity at N=8, sort bench at N=8 and N=4 (peer plan), phases at N=8
mkdir -p gpurun_out
for N in 8 4; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ $N = 8 ]; then
echo "== dist_check N=$N"
timeout 600 $TR --master-port 29511 tests/dist_check_worker.py > gpurun_out/dist_check_$N.log 2>&1; echo "rc=$?"; grep -E "OK|MISMATCH|DIST_CHECK|rror" gpurun_out/dist_check_$N.log | cut -c1-230 | tail -20
fi
echo "== bench sort_u32 N=$N"
timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_sort_u32_N${N}_peer.json 2> gpurun_out/bench_N${N}_peer.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_sort_u32_N${N}_peer.json').read().strip().splitlines()[-1])
    print(d['value'], d['unit'], d['ms_per_step'], d['verified'], d.get('distributed'), d['step_ms'], d.get('e2e'))
except Exception as e:
    print('no json', e); print(open('gpurun_out/bench_N${N}_peer.err').read()[-1500:])
PY
done
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== phases N=8"
BCB_DIST_PROFILE=1 timeout 600 $TR --master-port 29514 bench.py --gpus $N --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_prof_N${N}_peer.json 2>> gpurun_out/bench_N${N}_peer.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_prof_N8_peer.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], d.get('distributed'))
except Exception as e:
    print('no json', e)
PY
for w in scan_i32 reduce_i32; do
  echo "== bench $w N=$N"
  timeout 600 $TR --master-port 29515 bench.py --gpus $N --steps 5 --warmup 3 --workload $w --no-e2e > gpurun_out/bench_${w}_N$N.json 2> gpurun_out/bench_${w}_N$N.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${w}_N$N.json').read().strip().splitlines()[-1])
    print(d['value'], d['unit'], d['ms_per_step'])
except Exception as e:
    print('no json', e); print(open('gpurun_out/bench_${w}_N$N.err').read()[-1500:])
PY
done
