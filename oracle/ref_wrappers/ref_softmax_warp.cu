// Include-wrapper around the UNMODIFIED reference file
//   Softmax micro kernel/softmax_warp_shfl_baseline/softmax_warp_shfl_baseline.cu
// exposing a launcher for softmax_warp_shfl_kernel (:69-115): one warp per row, N <= 32, four rows per block
// (launch shape :147-152).
#define main ref_softmax_warp_unused_main
#include "Softmax micro kernel/softmax_warp_shfl_baseline/softmax_warp_shfl_baseline.cu"
#undef main

extern "C" int ref_softmax_warp_launch(const float* dX, float* dP, int rows, int N, void* stream) {
  if (N > WARP_SIZE) return -1;
  const int grid = (rows + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
  softmax_warp_shfl_kernel<<<grid, BLOCK_THREADS, 0, (cudaStream_t)stream>>>(dX, dP, rows, N);
  return (int)cudaGetLastError();
}
