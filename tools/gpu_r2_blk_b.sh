#!/bin/bash
# full ncu capture of the block-owner kernel on a quarter-size batch
mkdir -p gpurun_out
export NF_BENCH_NPIX=256

timeout 400 ncu --set full --clock-control none --import-source on -k regex:nf_nh3_kernel -s 4 -c 1 -f -o gpurun_out/r02_blk3 python tools/ab_kernel.py blkq 3 > gpurun_out/ncu_blk.log 2>&1; echo "capture rc=$?"
ls -la gpurun_out/*.ncu-rep
