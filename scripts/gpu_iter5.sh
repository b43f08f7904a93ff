#!/bin/bash
mkdir -p gpurun_out
echo "== pytest" ; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_gpu.log
for w in scan_i32 scan_f32 reduce_i32 reduce_f32; do
  echo "== $w" ; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --workload $w > gpurun_out/bench_$w.json 2>> gpurun_out/bench_other.err ; python -c "
import json;d=json.load(open('gpurun_out/bench_$w.json'));print(d['value'],d['unit'],d['ms_per_step'],d['roofline'] and d['roofline']['frac'],d['verified'], min(d['step_ms']))"
done
echo "== dist check (2 GPUs)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py > gpurun_out/dist_check.log 2>&1; grep -E "OK|MISMATCH|DIST_CHECK|Error|error" gpurun_out/dist_check.log | tail -20
echo "== bench 2 GPUs"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_sort_2gpu.json 2> gpurun_out/bench_2gpu.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_sort_2gpu.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['step_ms'])"; tail -3 gpurun_out/bench_2gpu.err
tail -n 5 gpurun_out/bench_other.err
