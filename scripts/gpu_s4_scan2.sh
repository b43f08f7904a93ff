#!/bin/bash
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 tests/dist_check_worker.py 2>&1 | grep -E "MISMATCH|DIST_CHECK|rror|scan|reduce" | cut -c1-120 | tail -12
for w in scan_i32 reduce_i32 scan_f32; do
  timeout 300 $TR --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 --no-configs --no-e2e --no-cpu --workload $w 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['value'],1), d['unit'], round(d['ms_per_step'],3), d.get('verified'))"
done
