#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_realdata.py tests/test_gpu_cube.py tests/test_gpu_sampler.py -x -q > gpurun_out/pytest_b.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_b.log
timeout 900 python bench.py --steps 3 --warmup 3 --cube-size 48 --scale-cube 48x32 --no-cpu > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_small.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_small.json').read().splitlines()[-1])
print('value', d['value'], 'frac', d['roofline']['frac'])
for k in ('cube_fit_config2', 'cube_fit', 'cube_fit_full'):
    if k in d: print(k, json.dumps(d[k])[:900])
PY
