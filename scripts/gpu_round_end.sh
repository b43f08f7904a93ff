#!/bin/bash
# usage: gpu_round_end.sh N -- what the driver runs at round end: the GPU test suite (N = 1 only), the default bench line, the
# reference arm; for N > 1 under torchrun
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s4_pytest_full.log 2>&1; tail -4 gpurun_out/s4_pytest_full.log
  timeout 900 python bench.py > gpurun_out/s4_bench_default.json 2> gpurun_out/s4_bench_default.err; echo "bench rc=$?"
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N > gpurun_out/s4_bench_N$N.json 2> gpurun_out/s4_bench_N$N.err; echo "bench rc=$?"
fi
python - <<PY
import json
f = 'gpurun_out/s4_bench_default.json' if '$N' == '1' else 'gpurun_out/s4_bench_N$N.json'
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(d['metric'], round(d['value'],2), d['unit'], 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'e2e', d.get('e2e',{}).get('value'), 'verified', d.get('verified'))
    for k,v in d.get('configs',{}).items():
        print('  ', k, round(v.get('value',0),2), v.get('unit'), 'frac', round(v.get('roofline',{}).get('frac',0),3) if v.get('roofline') else None, v.get('verified'), v.get('scaling'), (v.get('distributed') or {}).get('plan'))
    print('cpu', d.get('cpu_baseline',{}).get('value'), 'launches', d.get('gpu_launches'), 'clocks', d.get('clocks'))
except Exception as e:
    print('no json', e); print(open(f.replace('.json','.err')).read()[-2000:])
PY
