// Include-wrapper around the UNMODIFIED reference file
//   GEMM/QK Tensor Core/qk_wmma_5_1_2.cu
// exposing qk_cpu_ref (:101-122) and the kernel QK_wmma_kernel (:32-98).
#define main ref_qk_wmma_unused_main
#include "GEMM/QK Tensor Core/qk_wmma_5_1_2.cu"
#undef main
#include <cstring>

extern "C" {

void ref_qk_cpu(const float* q, const float* k, float* s, int bh, int n, int d, float scale) {
  std::vector<float> Q(q, q + (size_t)bh * n * d), K(k, k + (size_t)bh * n * d), S((size_t)bh * n * n, 0.0f);
  qk_cpu_ref(Q, K, S, bh, n, d, scale);
  memcpy(s, S.data(), S.size() * 4);
}

// device pointers; n, d multiples of 16 (:30); one warp per 16x16 score tile as in the harness
int ref_qk_launch(const void* dQ, const void* dK, float* dS, int bh, int n, int d, float scale, void* stream) {
  dim3 block(32, 1, 1), grid(n / 16, n / 16, bh);
  QK_wmma_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const __half*)dQ, (const __half*)dK, dS, bh, n, d, scale);
  return (int)cudaGetLastError();
}

}  // extern "C"
