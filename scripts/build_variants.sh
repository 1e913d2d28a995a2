#!/bin/bash
# Developer tool: builds tuning / ablation / trace variants of the library next to the product .so.
# usage: scripts/build_variants.sh name1:"-DFLAG ..." name2:"..."   -> libfa_b200_<name>.so
set -e
cd "$(dirname "$0")/../flashattention_kernel_project_b200/csrc"
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC $flags -o ../libfa_b200_${name}.so api.cu &
done
wait
