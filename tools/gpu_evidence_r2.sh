#!/bin/bash
# evidence pass: sampler-kernel ncu captures + launch list of one ncomp = 3 wave, compute-sanitizer with two streams,
# refreshed launch list + full capture of the likelihood kernel
mkdir -p gpurun_out
cmd="python tools/ns_profile3.py 32 3 1000000"
timeout 300 $cmd > gpurun_out/ns3_plain.log 2>&1; echo "plain rc=$?"; tail -1 gpurun_out/ns3_plain.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 1200 --csv --log-file gpurun_out/r02_ns_wave_launches.csv \
    $cmd > gpurun_out/ncu_ns_list.log 2>&1; echo "ns launch list rc=$?"
for k in ns_bounds_single_kernel ns_propose_kernel ns_update_kernel ns_compact_kernel; do
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 1000 -c 1 -f -o gpurun_out/r02_$k \
        $cmd > gpurun_out/ncu_$k.log 2>&1; echo "$k capture rc=$?"
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r02_sanitizer_racecheck.log
bcmd="python bench.py --steps 2 --warmup 3 --no-cpu --cube-size 0 --scale-cube 0x0 --no-gauss"
timeout 200 $bcmd > gpurun_out/bench_prof.json 2> gpurun_out/bench_prof.err; echo "plain bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $bcmd > gpurun_out/ncu_list.log 2>&1; echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:nf_nh3_kernel -s 3 -c 1 -f -o gpurun_out/r02_prof $bcmd > gpurun_out/ncu_full.log 2>&1; echo "full capture rc=$?"
ls -la gpurun_out/*.ncu-rep
