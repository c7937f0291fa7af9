#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
for bpg in 8; do
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$bpg bench.py --gpus $N --steps 3 --warmup 3 --no-gauss --scale-cube 256x128 --blocks-per-gpu $bpg > gpurun_out/bench_n${N}_b$bpg.json 2> gpurun_out/bench_n${N}_b$bpg.err ) 2>&1 | grep real; echo "bench rc=$?"; grep -i "error\|Traceback" gpurun_out/bench_n${N}_b$bpg.err | head -3
python - $N $bpg <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/bench_n{sys.argv[1]}_b{sys.argv[2]}.json').read().splitlines()[-1])
c = d['cube_fit']
print('n_gpus', d['n_gpus'], 'bpg', sys.argv[2], 'value', d['value'], '| cube', {k: c[k] for k in ('value', 'seconds', 'api', 'likelihood_evals_per_pixel', 'rank_busy_fraction', 'rank_blocks', 'nbest_matches_truth')})
PY
done
