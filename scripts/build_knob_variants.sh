#!/bin/bash
# usage: build_knob_variants.sh "SPEC SPLIT [extra flags]" ...   -> variants/libfa_trace_k<SPEC>_g<SPLIT>.so (FA_TRACE builds)
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $ROOT/variants
for cfg in "$@"; do
  set -- $cfg
  k=$1; g=$2; shift 2
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -DFA_TRACE \
    -DFA_SPEC_KEYS=$k -DFA_SPLIT_GROUPS=$g "$@" -o $ROOT/variants/libfa_trace_k${k}_g${g}$(echo "$@" | tr -d ' -=').so \
    $ROOT/flashattention_kernel_project_b200/csrc/api.cu > /tmp/build_k${k}_g${g}.log 2>&1 &
done
wait
grep -l error /tmp/build_k*.log 2>/dev/null | head
ls $ROOT/variants | wc -l
