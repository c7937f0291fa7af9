#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/sanitize_small.py 2>&1 | tail -2
for v in "--size 128" "--size 128 --streams 4 --pps 4096" "--size 128 --streams 2 --pps 8192" "--size 128 --nprop 64" "--size 128 --nprop 64 --streams 4 --pps 4096"; do
  timeout 600 python tools/bench_cube.py $v > gpurun_out/cube_f.json 2> gpurun_out/cube_f.err; echo "[$v] rc=$?"; python - <<'PY'
import json
d = json.loads(open('gpurun_out/cube_f.json').read().splitlines()[-1])
print({k: d[k] for k in ('pixels_per_s', 'evals_per_pixel', 'evals_per_s', 'nbest_agreement', 'seconds_by_ncomp', 'n_truncated')})
PY
done
