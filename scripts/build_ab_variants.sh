#!/bin/bash
# usage: build_ab_variants.sh "name:flags" ...  -> variants/libfa_<name>.so (product builds, no trace)
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $ROOT/variants
for cfg in "$@"; do
  n=${cfg%%:*}; f=${cfg#*:}
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC $f \
    -o $ROOT/variants/libfa_$n.so $ROOT/flashattention_kernel_project_b200/csrc/api.cu > /tmp/ab_$n.log 2>&1 &
done
wait
grep -l "error" /tmp/ab_*.log 2>/dev/null | head -3
ls $ROOT/variants | wc -l
