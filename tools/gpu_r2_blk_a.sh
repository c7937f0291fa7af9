#!/bin/bash
# first contact of the block-owner kernel: A/B timing against the shipped kernel, lnL difference, strict parity, tests
mkdir -p gpurun_out
timeout 200 python tools/ab_kernel.py blk 3 1 2 4 2>&1 | tail -6
NF_NH3_KERNEL=v8 timeout 200 python tools/ab_kernel.py v8 3 1 2 4 2>&1 | tail -6
python tools/ab_kernel.py --diff blk v8
timeout 300 python tools/parity_strict.py 32768 2>&1 | tail -9
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_n2hp.py tests/test_gpu_device_abi.py tests/test_postprocess.py -x -q -m gpu 2>&1 | tail -8
