// Include-wrapper around the UNMODIFIED reference file
//   GEMM/FlashAttention Forward Fused/flashattn_forward_fused_5_4_2.cu
// exposing flashattn_cpu_ref (:224-271; fp32 inputs, double accumulation, scale = 1/sqrt(D)).
#define main ref_fa542_unused_main
#include "GEMM/FlashAttention Forward Fused/flashattn_forward_fused_5_4_2.cu"
#undef main

extern "C" void ref_fa542_cpu_ref(const float* q, const float* k, const float* v, float* o, int BH, int N, int D) {
  const size_t n = (size_t)BH * N * D;
  std::vector<float> Q(q, q + n), K(k, k + n), V(v, v + n), O(n, 0.0f);
  flashattn_cpu_ref(Q, K, V, O, BH, N, D);
  memcpy(o, O.data(), n * 4);
}
