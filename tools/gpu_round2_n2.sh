#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
echo skip pytest
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --scale-cube 256x128 --no-gauss > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -4 gpurun_out/bench_n2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_n2.json').read().splitlines()[-1])
print('n_gpus', d['n_gpus'], 'value', d['value'], 'e2e', d['e2e']['value'], 'gauss', d.get('gauss_loglike', {}).get('value'))
for k in ('cube_fit_config2', 'cube_fit', 'cube_fit_full'):
    if k in d: print(k, json.dumps(d[k])[:1600])
PY
