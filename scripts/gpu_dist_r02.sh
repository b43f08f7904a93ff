#!/bin/bash
# usage: gpu_dist_r02.sh N -- multi-GPU parity (pytest, both plans), the default bench line at N GPUs (weak headline +
# strong-scaling sort + scan / reduce configs), and a phase profile of the sort
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu 2>&1 | tail -3
timeout 900 $TR --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_N${N}.json 2> gpurun_out/r02_bench_N${N}.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_N${N}.json').read().strip().splitlines()[-1])
    print('headline', round(d['value'],2), d['unit'], round(d['ms_per_step'],3), d['verified'], d.get('distributed'))
    print('kernels', d['roofline']['kernel_ms_per_step'], 'frac', round(d['roofline']['frac'],3))
    print('e2e', d.get('e2e'))
    for k,r in d['configs'].items():
        print(k, round(r['value'],2), r['unit'], round(r['ms_per_step'],3), r['scaling'], r['verified'], (r.get('roofline') or {}).get('kernel_ms_per_step'))
except Exception as e:
    print('no json', e); print(open('gpurun_out/r02_bench_N${N}.err').read()[-2500:])
PY
for sc in weak strong; do
BCB_DIST_PROFILE=1 timeout 600 $TR --master-port 29514 bench.py --gpus $N --steps 3 --warmup 3 --no-configs --no-e2e --scaling $sc > gpurun_out/r02_prof_N${N}_$sc.json 2>> gpurun_out/r02_bench_N${N}.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_prof_N${N}_$sc.json').read().strip().splitlines()[-1])
    print('$sc profiled', round(d['value'],2), round(d['ms_per_step'],3), d.get('distributed'))
except Exception as e:
    print('no json', e)
PY
done
