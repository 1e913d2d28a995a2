// Include-wrapper around the UNMODIFIED reference file
//   GEMM/gemm_fused_epilogue/gemm_fused_epilogue.cu
// exposing its cpu_gemm_fused_ref (:116-136) and its kernel gemm_tiled_fused_kernel (:44-111), the fp32 sibling
// of the tensor-core fused GEMM.
#define main ref_gemm_fp32_fused_unused_main
#include "GEMM/gemm_fused_epilogue/gemm_fused_epilogue.cu"
#undef main
#include <cstring>

extern "C" {

void ref_gemm_fp32_fused_cpu(const float* a, const float* b, const float* bias, const float* residual, float* c, int m,
                             int n, int k) {
  std::vector<float> A(a, a + (size_t)m * k), B(b, b + (size_t)k * n), Bi(bias, bias + n),
      R(residual, residual + (size_t)m * n), C((size_t)m * n, 0.0f);
  cpu_gemm_fused_ref(A, B, Bi, R, C, m, n, k);
  memcpy(c, C.data(), C.size() * 4);
}

// device pointers; any m, n, k (the kernel guards its edges, :58-71,:88)
int ref_gemm_fp32_fused_launch(const float* dA, const float* dB, const float* dBias, const float* dRes, float* dC,
                               int m, int n, int k, void* stream) {
  dim3 block(TILE, TILE), grid((n + TILE - 1) / TILE, (m + TILE - 1) / TILE);
  gemm_tiled_fused_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(dA, dB, dBias, dRes, dC, m, n, k);
  return (int)cudaGetLastError();
}

}  // extern "C"
