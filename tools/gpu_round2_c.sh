#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sampler.py tests/test_gpu_cube.py -x -q > gpurun_out/pytest_c.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_c.log
for v in "" "--single-ellipsoid"; do
  timeout 600 python tools/bench_cube.py --size 64 $v > gpurun_out/cube64_m$v.json 2> gpurun_out/cube64_m$v.err; echo "cube64 $v rc=$?"; cat gpurun_out/cube64_m$v.json
done
NF_NS_PROFILE=1 timeout 300 python tools/ns_profile3.py 48 2 100000 2>&1 | grep ns-prof | tail -2
NF_NS_PROFILE=1 timeout 300 python tools/ns_profile3.py 48 3 100000 2>&1 | grep ns-prof | tail -2
