#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_cube.py --size 64 > gpurun_out/cube64_d.json 2> gpurun_out/cube64_d.err; echo "cube64 rc=$?"; cat gpurun_out/cube64_d.json
timeout 600 python tools/bench_cube.py --size 64 --mmodal > gpurun_out/cube64_dm.json 2> gpurun_out/cube64_dm.err; echo "cube64 mmodal rc=$?"; cat gpurun_out/cube64_dm.json
timeout 600 python tools/parity_strict.py 131072 2>&1 | tail -8
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
