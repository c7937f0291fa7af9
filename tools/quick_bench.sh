#!/bin/bash
# usage: tools/quick_bench.sh lib1.so lib2.so ...   (runs the short bench for each library build)
for lib in "$@"; do
  NESTFIT_B200_LIB=$PWD/$lib python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', 'evals/s %.4g'%d['value'], 'ms %.3f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'frac %.3f'%d['roofline']['frac'])"
done
