#!/bin/bash
# A/B: record prefetch in the main loop of the likelihood kernel (BLK_PREFETCH) against the shipped build
mkdir -p gpurun_out
timeout 120 python tools/ab_kernel.py base 3 1 2 4 2>&1 | tail -4
NESTFIT_B200_LIB=$PWD/nestfit_b200/_variants/lib_pf.so timeout 120 python tools/ab_kernel.py pf 3 1 2 4 2>&1 | tail -4
python tools/ab_kernel.py --diff base pf
