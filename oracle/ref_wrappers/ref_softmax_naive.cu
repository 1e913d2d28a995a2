// Include-wrapper around the UNMODIFIED reference file
//   Softmax micro kernel/softmax_naive_baseline/softmax_naive_baseline.cu
// exposing a launcher for softmax_naive_baseline_kernel (:56-130): one 128-thread block per row,
// dynamic smem (N + 128) floats (:172-198).
#define main ref_softmax_naive_unused_main
#include "Softmax micro kernel/softmax_naive_baseline/softmax_naive_baseline.cu"
#undef main

extern "C" int ref_softmax_naive_launch(const float* dX, float* dP, int rows, int N, void* stream) {
  const int threads = BLOCK_THREADS;
  size_t smem = ((size_t)N + threads) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(softmax_naive_baseline_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return (int)e;
  softmax_naive_baseline_kernel<<<rows, threads, smem, (cudaStream_t)stream>>>(dX, dP, rows, N);
  return (int)cudaGetLastError();
}
