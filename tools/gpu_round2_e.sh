#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -3 gpurun_out/bench_full.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_full.json').read().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e'].get('pageable_host_buffers', {}).get('value'), 'frac', d['roofline']['frac'])
print('cpu', json.dumps(d.get('cpu_baseline'))[:600])
print('gauss', json.dumps(d.get('gauss_loglike'))[:300])
for k in ('cube_fit_config2', 'cube_fit', 'cube_fit_full'):
    if k in d: print(k, json.dumps(d[k])[:1500])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nf_gauss_kernel -s 4 -c 1 -f -o gpurun_out/r02_gauss_v2 python tools/bench_gauss.py 1048576 > gpurun_out/ncu_gauss2.log 2>&1; echo "gauss capture rc=$?"
