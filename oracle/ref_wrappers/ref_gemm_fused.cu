// Include-wrapper around the UNMODIFIED reference file
//   GEMM/tensorcore_gemm_fused_epilogue/tensorcore_gemm_fused_epilogue.cu
// exposing cpu_gemm_fused_ref (:136-155) and the kernel wmma_gemm_fused_kernel (:53-130).
#define main ref_gemm_fused_unused_main
#include "GEMM/tensorcore_gemm_fused_epilogue/tensorcore_gemm_fused_epilogue.cu"
#undef main

extern "C" {

void ref_gemm_fused_cpu(const float* a, const float* b, const float* bias, const float* residual, float* c, int m,
                        int n, int k) {
  std::vector<float> A(a, a + (size_t)m * k), B(b, b + (size_t)k * n), Bi(bias, bias + n),
      R(residual, residual + (size_t)m * n), C((size_t)m * n, 0.0f);
  cpu_gemm_fused_ref(A, B, Bi, R, C, m, n, k);
  memcpy(c, C.data(), C.size() * 4);
}

// The reference's fixture: mt19937(42), uniform(-1,1), draw order A, B, bias, residual (:178-184).
void ref_gemm_fused_fill(unsigned seed, float* a, size_t na, float* b, size_t nb, float* bias, size_t nbias,
                         float* res, size_t nres) {
  std::mt19937 rng(seed);
  std::uniform_real_distribution<float> dist(-1.0f, 1.0f);
  for (size_t i = 0; i < na; ++i) a[i] = dist(rng);
  for (size_t i = 0; i < nb; ++i) b[i] = dist(rng);
  for (size_t i = 0; i < nbias; ++i) bias[i] = dist(rng);
  for (size_t i = 0; i < nres; ++i) res[i] = dist(rng);
}

// device pointers; m, n, k multiples of 16 (:35)
int ref_gemm_fused_launch(const void* dA, const void* dB, const float* dBias, const float* dRes, float* dC, int m,
                          int n, int k, void* stream) {
  dim3 block(32, 1, 1), grid(n / 16, m / 16, 1);
  wmma_gemm_fused_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const half*)dA, (const half*)dB, dBias, dRes, dC, m,
                                                                    n, k);
  return (int)cudaGetLastError();
}

}  // extern "C"
