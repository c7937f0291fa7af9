#!/bin/bash
# One gpurun call = one validation pass; every step runs under its own timeout and logs to gpurun_out/.
#   gpurun --timeout 900 -- 'bash tools/gpu_validate.sh quick'     tests + kernel-only bench line
#   gpurun --timeout 900 -- 'bash tools/gpu_validate.sh full'      tests + smoke + default bench + reference arm
#   gpurun --timeout 900 -- 'bash tools/gpu_validate.sh profile'   ncu launch list + one full capture of the likelihood kernel
#                                                                  (then: python tools/refresh_profiles.py gpurun_out/prof.ncu-rep 1048576)
mode=${1:-quick}
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json, sys
lines = open(sys.argv[1]).read().splitlines()
d = json.loads(lines[-1])
print(len(lines), "line(s):", d.get("value"), "evals/s", "| e2e", d.get("e2e", {}).get("value"), "| roofline", d.get("roofline", {}).get("frac"),
      "| gauss", d.get("gauss_loglike", {}).get("value"), "| cube", d.get("cube_fit", {}).get("value"),
      "| cube c2", d.get("cube_fit_config2", {}).get("value"),
      "| cpu", d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("parity_ok"))
PY
}
case "$mode" in
quick|full)
    timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
    if [ "$mode" = quick ]; then
        timeout 200 python bench.py --no-cpu --cube-size 0 --scale-cube 0x0 --steps 5 --warmup 3 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
        summ gpurun_out/bench_quick.json
    else
        timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
        timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"
        summ gpurun_out/bench_full.json
        timeout 200 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/bench_ref.err | cut -c1-300
    fi
    ;;
profile)
    cmd="python bench.py --steps 2 --warmup 3 --no-cpu --cube-size 0 --scale-cube 0x0 --no-gauss"
    timeout 200 $cmd > gpurun_out/bench_prof.json 2> gpurun_out/bench_prof.err; echo "plain run rc=$?"
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
        $cmd > gpurun_out/ncu_list.log 2>&1; echo "launch list rc=$?"
    # launches 1-3 of the kernel build the synthetic problem (predict); the 4th is the first full 2^20-vector step
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:nf_nh3_kernel -s 3 -c 1 -f -o gpurun_out/prof \
        $cmd > gpurun_out/ncu_full.log 2>&1; echo "full capture rc=$?"; ls -la gpurun_out/prof.ncu-rep
    ;;
*) echo "usage: $0 quick|full|profile"; exit 2 ;;
esac
