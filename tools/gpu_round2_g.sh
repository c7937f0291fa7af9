#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/repro_concurrent.py 128 64 8 8 4 > gpurun_out/repro_a.log 2>&1; echo "8 threads (memset fix) rc=$?"; tail -2 gpurun_out/repro_a.log | cut -c1-300
NF_SYNC_ALLOC=1 timeout 600 python tools/repro_concurrent.py 128 64 8 8 4 > gpurun_out/repro_b.log 2>&1; echo "8 threads sync alloc rc=$?"; tail -2 gpurun_out/repro_b.log | cut -c1-300
timeout 600 python tools/repro_concurrent.py 128 64 8 8 4 nosink > gpurun_out/repro_c.log 2>&1; echo "8 threads no sink rc=$?"; tail -2 gpurun_out/repro_c.log | cut -c1-300
