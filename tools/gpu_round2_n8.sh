#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | head -2 | tail -1; df -h /tmp /dev/shm | tail -2
( time timeout 1400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err ) 2>&1 | grep real; echo "bench rc=$?"; grep -i "error\|Traceback" gpurun_out/bench_n8.err | head -3
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_n8.json').read().splitlines()[-1])
print('n_gpus', d['n_gpus'], 'value', d['value'], 'e2e', d['e2e']['value'], 'gauss', d.get('gauss_loglike', {}).get('value'))
for k in ('cube_fit', 'cube_fit_full'):
    if k in d:
        c = d[k]
        print(k, {q: c[q] for q in ('value', 'seconds', 'api', 'likelihood_evals_per_pixel', 'rank_busy_fraction', 'rank_blocks', 'nbest_matches_truth', 'store')})
PY
