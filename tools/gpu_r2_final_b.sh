#!/bin/bash
# round-2 closing run: the default bench line (all legs), the strict parity sweep of the block-owner kernel, the GPU
# test suite and smoke
mkdir -p gpurun_out
( time timeout 1200 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_full.err
timeout 400 python tools/parity_strict.py 65536 > gpurun_out/parity_strict.log 2>&1; echo "parity rc=$?"; cat gpurun_out/parity_strict.log
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep '3 components, 32 runs' gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_full.json').read().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'pageable', d['e2e']['pageable_host_buffers']['value'], 'frac', d['roofline']['frac'])
print('cpu', d.get('cpu_baseline'))
print('gauss', d.get('gauss_loglike'))
for k in ('cube_fit_config2', 'cube_fit'):
    if k in d: print(k, json.dumps(d[k])[:1500])
PY
