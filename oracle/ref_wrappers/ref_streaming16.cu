// Include-wrapper around the UNMODIFIED reference file
//   Streaming_FlashAttention_Forward_Kernel/flashattn_streaming_16x16.cu
// exposing flashattn_streaming_cpu_ref (:280-348) and the kernel flashattn_streaming_16x16_kernel (:91-276),
// whose tile-to-tile (m, l, y) merge (:228-261) is the rule fa_forward_partial carries between K/V shards.
// Shapes fixed by the file: M = Kdim = Dv = 16, seq_len a multiple of 16; K is [B, Kdim, L].
#define main ref_streaming16_unused_main
#include "Streaming_FlashAttention_Forward_Kernel/flashattn_streaming_16x16.cu"
#undef main
#include <cstring>

extern "C" {

void ref_streaming16_cpu(const void* q, const void* kt, const void* v, float* o, int num_batches, int seq_len, float scale) {
  const __half* hq = (const __half*)q;
  const __half* hk = (const __half*)kt;
  const __half* hv = (const __half*)v;
  std::vector<__half> Q(hq, hq + (size_t)num_batches * M * Kdim), K(hk, hk + (size_t)num_batches * Kdim * seq_len),
      V(hv, hv + (size_t)num_batches * seq_len * Dv);
  std::vector<float> O((size_t)num_batches * M * Dv, 0.0f);
  flashattn_streaming_cpu_ref(Q, K, V, O, num_batches, seq_len, scale);
  memcpy(o, O.data(), O.size() * 4);
}

// The file's fixture: mt19937(42), normal(0,1), draw order Q, K, V (:365-380).
void ref_streaming16_fill(unsigned seed, void* q, size_t nq, void* k, size_t nk, void* v, size_t nv) {
  std::mt19937 rng(seed);
  std::normal_distribution<float> dist(0.0f, 1.0f);
  __half* hq = (__half*)q;
  __half* hk = (__half*)k;
  __half* hv = (__half*)v;
  for (size_t i = 0; i < nq; ++i) hq[i] = __float2half(dist(rng));
  for (size_t i = 0; i < nk; ++i) hk[i] = __float2half(dist(rng));
  for (size_t i = 0; i < nv; ++i) hv[i] = __float2half(dist(rng));
}

// device pointers; one warp per batch element as in the harness (:398-405)
int ref_streaming16_launch(const void* dQ, const void* dKt, const void* dV, float* dO, int num_batches, int seq_len,
                           float scale, void* stream) {
  dim3 block(WARP_SIZE, 1, 1), grid(num_batches, 1, 1);
  flashattn_streaming_16x16_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const __half*)dQ, (const __half*)dKt,
                                                                             (const __half*)dV, dO, num_batches, seq_len,
                                                                             scale);
  return (int)cudaGetLastError();
}

}  // extern "C"
