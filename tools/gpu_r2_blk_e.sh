#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/ab_kernel.py fin 3 1 2 4 2>&1 | tail -4
timeout 300 python tools/parity_strict.py 16384 2>&1 | tail -8
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
