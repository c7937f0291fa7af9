#!/bin/bash
# configs[2] cube leg of bench.py with sub-blocks of the wave escalating concurrently (CubeFitter n_streams)
mkdir -p gpurun_out
( time timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu --no-gauss --scale-cube 0x0 --cube-streams ${1:-4} --cube-pps ${2:-1024} > gpurun_out/bench_streams.json 2> gpurun_out/bench_streams.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_streams.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_streams.json').read().splitlines()[-1])
print(json.dumps(d['cube_fit_config2'])[:1400])
PY
