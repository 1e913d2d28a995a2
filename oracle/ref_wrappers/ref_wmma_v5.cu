// Include-wrapper around the UNMODIFIED reference file
//   FlashAttention/flashattn_forward_memory_bound/flashattn_forward_wmma_v5_cp_async.cu
// exposing a launcher for flashattn_forward_wmma_v5_cp_async_kernel (:99-459).  At D = 128 the
// kernel needs ~78 KB of dynamic shared memory; the reference never opts in
// (cudaFuncSetAttribute), the launcher does so without touching reference code.
#define main ref_wmma_v5_unused_main
#include "FlashAttention/flashattn_forward_memory_bound/flashattn_forward_wmma_v5_cp_async.cu"
#undef main

extern "C" int ref_wmma_v5_launch(const void* dQ, const void* dK, const void* dV, float* dO, int BH, int N, int D,
                                  float scale, void* stream) {
  // Q + 2 K + 2 V tiles (half), S (float), 3 row-stat vectors, W (half), O and PV (float) + slack
  size_t smem = (size_t)D * (BLOCK_M * 2 + 4 * TILE_N * 2 + 2 * BLOCK_M * 4) + (size_t)BLOCK_M * TILE_N * 6 +
                3 * BLOCK_M * 4 + 256;
  cudaError_t e = cudaFuncSetAttribute(flashattn_forward_wmma_v5_cp_async_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((N + BLOCK_M - 1) / BLOCK_M, BH, 1), block(BLOCK_THREADS, 1, 1);
  flashattn_forward_wmma_v5_cp_async_kernel<<<grid, block, smem, (cudaStream_t)stream>>>(
      (const half*)dQ, (const half*)dK, (const half*)dV, dO, BH, N, D, scale);
  return (int)cudaGetLastError();
}
