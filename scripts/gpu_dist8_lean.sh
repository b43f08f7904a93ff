#!/bin/bash
mkdir -p gpurun_out
for N in 8 4; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_sort_u32_N${N}_peer.json 2> gpurun_out/bench_N${N}_peer.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_sort_u32_N${N}_peer.json').read().strip().splitlines()[-1])
    print($N, d['value'], d['unit'], d['ms_per_step'], d['verified'], d.get('distributed'), d['step_ms'])
except Exception as e:
    print('no json', e); print(open('gpurun_out/bench_N${N}_peer.err').read()[-1500:])
PY
done
N=8
BCB_DIST_PROFILE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_prof_N${N}_peer.json 2>> gpurun_out/bench_N${N}_peer.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_prof_N8_peer.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d.get('distributed'))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 tests/dist_check_worker.py 2>&1 | grep -E "MISMATCH|DIST_CHECK|rror" | head -5
