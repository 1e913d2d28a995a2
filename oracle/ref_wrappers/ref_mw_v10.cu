// Include-wrapper around the UNMODIFIED reference file
//   flashattn_warp_spc/flashattn_streaming_16x16_mw_v10.cu
// exposing a launcher for flashattn_streaming_16x16_kernel_mw_v10 (the best true-WMMA streaming kernels):
// Q [B,16,16], K_T [B,L,16] (row per key), V [B,L,16] half, O [B,16,16] float, L a multiple of 16, 64 threads per batch
// element (launch shape as in the harness).  Its host oracle is the streaming family's flashattn_streaming_cpu_ref,
// restated in oracle/fa_oracle.c and pinned through ref_streaming16.
#define main ref_mw_v10_unused_main
#include "flashattn_warp_spc/flashattn_streaming_16x16_mw_v10.cu"
#undef main

extern "C" int ref_mw_v10_launch(const void* dQ, const void* dKT, const void* dV, float* dO, int num_batches, int seq_len,
                                float scale, void* stream) {
  dim3 block(64, 1, 1), grid(num_batches, 1, 1);
  flashattn_streaming_16x16_kernel_mw_v10<<<grid, block, 0, (cudaStream_t)stream>>>((const __half*)dQ, (const __half*)dKT,
                                                                                  (const __half*)dV, dO, num_batches,
                                                                                  seq_len, scale);
  return (int)cudaGetLastError();
}
