#!/bin/bash
# A/B of kernel build variants (nestfit_b200/_variants/lib_<X>.so)
mkdir -p gpurun_out
for v in "$@"; do
  NESTFIT_B200_LIB=$PWD/nestfit_b200/_variants/lib_$v.so timeout 120 python tools/ab_kernel.py $v 3 1 4 2>&1 | tail -3
done
first=$1; shift
for v in "$@"; do echo "diff $first $v"; python tools/ab_kernel.py --diff $first $v; done
