#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/repro_concurrent.py 128 64 8 8 4 > gpurun_out/repro_a.log 2>&1; echo "8 threads rc=$?"; tail -2 gpurun_out/repro_a.log | cut -c1-300
timeout 600 python tools/repro_concurrent.py 128 64 16 8 4 > gpurun_out/repro_b.log 2>&1; echo "8 threads 16 blocks rc=$?"; tail -2 gpurun_out/repro_b.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
