#!/bin/bash
# round-2 closing evidence: short bench line, launch list and full capture of the block-owner likelihood kernel at
# 2^20 evals, the GPU test suite, smoke
mkdir -p gpurun_out
bcmd="python bench.py --steps 20 --warmup 3 --no-cpu --cube-size 0 --scale-cube 0x0 --no-gauss"
timeout 200 $bcmd > gpurun_out/bench_prof.json 2> gpurun_out/bench_prof.err; echo "plain bench rc=$?"; tail -c 600 gpurun_out/bench_prof.err
pcmd="python bench.py --steps 2 --warmup 3 --no-cpu --cube-size 0 --scale-cube 0x0 --no-gauss"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $pcmd > gpurun_out/ncu_list.log 2>&1; echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:nf_nh3_kernel -s 3 -c 1 -f -o gpurun_out/r02_prof $pcmd > gpurun_out/ncu_full.log 2>&1; echo "full capture rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_prof.json').read().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'pageable', d['e2e']['pageable_host_buffers']['value'], 'frac', d['roofline']['frac'])
PY
