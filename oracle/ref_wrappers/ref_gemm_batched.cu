// Include-wrapper around the UNMODIFIED reference file
//   GEMM/gemm_tensorcore_batched/gemm_4_9_tensorcore_batched.cu
// exposing cpu_batched_gemm_ref (:94-121) and the kernel wmma_batched_gemm_kernel (:42-92).
// The file names its problem size with global constexprs M, N, K, BATCH; the functions take their own.
#define main ref_gemm_batched_unused_main
#include "GEMM/gemm_tensorcore_batched/gemm_4_9_tensorcore_batched.cu"
#undef main
#include <cstring>

extern "C" {

void ref_gemm_batched_cpu(const float* a, const float* b, float* c, int batch, int m, int n, int k) {
  std::vector<float> A(a, a + (size_t)batch * m * k), B(b, b + (size_t)batch * k * n), C((size_t)batch * m * n, 0.0f);
  cpu_batched_gemm_ref(A, B, C, batch, m, n, k);
  memcpy(c, C.data(), C.size() * 4);
}

// device pointers; m, n, k multiples of 16 (:36); grid as in the harness: one 32-thread block per 16x16 tile and batch
int ref_gemm_batched_launch(const void* dA, const void* dB, float* dC, int batch, int m, int n, int k, void* stream) {
  dim3 block(32, 1, 1), grid(n / 16, m / 16, batch);
  wmma_batched_gemm_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const half*)dA, (const half*)dB, dC, batch, m, n, k);
  return (int)cudaGetLastError();
}

}  // extern "C"
