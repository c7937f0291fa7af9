#!/bin/bash
# round-2 validation pass A: GPU tests, then cube fits (no store / store / old behaviour)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_cube.py --size 64 > gpurun_out/cube64_new.json 2> gpurun_out/cube64_new.err; echo "cube64 rc=$?"; cat gpurun_out/cube64_new.json; tail -3 gpurun_out/cube64_new.err
timeout 600 python tools/bench_cube.py --size 64 --store /tmp/nfstore > gpurun_out/cube64_store.json 2> gpurun_out/cube64_store.err; echo "cube64 store rc=$?"; cat gpurun_out/cube64_store.json; tail -3 gpurun_out/cube64_store.err
