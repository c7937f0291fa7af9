#!/bin/bash
# ncu evidence for the sampler kernels and the Gaussian-model kernel (one gpurun call).
mkdir -p gpurun_out
cmd="python tools/ns_profile3.py 24 3 300"
timeout 300 $cmd > gpurun_out/ns3_plain.log 2>&1; echo "plain rc=$?"; tail -1 gpurun_out/ns3_plain.log
NF_NS_PROFILE=1 timeout 300 python tools/ns_profile3.py 48 3 100000 > gpurun_out/ns3_split.log 2>&1; echo "split rc=$?"; grep ns-prof gpurun_out/ns3_split.log | tail -3
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_ns_launches.csv \
    $cmd > gpurun_out/ncu_ns_list.log 2>&1; echo "ns launch list rc=$?"
for k in ns_bounds_kernel ns_propose_kernel ns_update_kernel ns_compact_kernel; do
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 200 -c 1 -f -o gpurun_out/r02_$k \
        $cmd > gpurun_out/ncu_$k.log 2>&1; echo "$k capture rc=$?"
done
timeout 200 python tools/bench_gauss.py 1048576 > gpurun_out/gauss_plain.log 2>&1; echo "gauss rc=$?"; tail -1 gpurun_out/gauss_plain.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nf_gauss_kernel -s 4 -c 1 -f -o gpurun_out/r02_gauss \
    python tools/bench_gauss.py 1048576 > gpurun_out/ncu_gauss.log 2>&1; echo "gauss capture rc=$?"
ls -la gpurun_out/*.ncu-rep
