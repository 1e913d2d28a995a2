#!/bin/bash
# usage: build_knob_variants.sh "SPEC SPLIT [extra flags]" ...   -> variants/libfa_trace_k<SPEC>_g<SPLIT>.so (FA_TRACE builds)
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $ROOT/variants
rm -f /tmp/build_knob_*.log
idx=0
for cfg in "$@"; do
  set -- $cfg
  k=$1; g=$2; shift 2
  tag=$(echo "$@" | sed 's/-DFA_//g; s/[ =]/_/g')
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -DFA_TRACE \
    -DFA_SPEC_KEYS=$k -DFA_SPLIT_GROUPS=$g "$@" -o $ROOT/variants/libfa_trace_k${k}_g${g}_${tag}.so \
    $ROOT/flashattention_kernel_project_b200/csrc/api.cu > /tmp/build_knob_$idx.log 2>&1 &
  idx=$((idx + 1))
done
wait
grep -l error /tmp/build_knob_*.log 2>/dev/null | head
ls $ROOT/variants | wc -l
