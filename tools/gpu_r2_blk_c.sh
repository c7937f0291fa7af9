#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/ab_kernel.py blk 3 1 2 4 2>&1 | tail -6
timeout 300 python tools/parity_strict.py 8192 2>&1 | tail -9
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_n2hp.py tests/test_gpu_device_abi.py tests/test_postprocess.py -x -q -m gpu 2>&1 | tail -5
