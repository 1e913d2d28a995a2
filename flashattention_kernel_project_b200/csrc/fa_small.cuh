// Attention forward for the reference's D = 16 programs (decode-shaped: few queries per batch
// element, Nq != Nkv allowed): one warp per query row, one lane per key of a 32-key group, online
// softmax per lane merged across the warp at the end.
//
// Replaces, at their native shapes, the comparators BASELINE.json names:
//   flashattn_2warp_nosmemW_online_softmax<16,16,4,32>
//     (flashattn_warp_spc/flashattn_streaming_16x32_2warp_softmax_v12i.cu:60-240)  Q [B,M,16], K,V [B,L,16]
//   flashattn_streaming_16x16_kernel_mw_v7_5
//     (flashattn_warp_spc_v2/flashattn_streaming_16x16_mw_v7_5_spinless_cpasync2stage.cu:84-248)  one query per b
// and the D = 16 streaming family (..mw_v10.cu:104-281).  These shapes move ~17 MB for 134 MFLOP: they
// are latency/HBM bound, so this path uses no tensor cores: a lane reads its key's 32 B of K and of V
// with one 128-bit load pair (a warp covers 1 KB contiguous), keeps (m, l, acc[16]) in registers and
// the merge rule of the reference (Streaming_FlashAttention_Forward_Kernel/flashattn_streaming_16x16.cu:247-261)
// is applied once per lane and once across lanes.  fp32 math, fp32 output as the reference kernels write it.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>

namespace fa {

__device__ __forceinline__ void load_half16(const __half* p, float* out) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
  const uint4 b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    out[2 * i] = f.x;
    out[2 * i + 1] = f.y;
  }
}

// Q [BH, Nq, 16], K, V [BH, Nkv, 16] fp16; O fp32 [BH, Nq, 16].  blockDim = 256 (8 rows per CTA).
__global__ void __launch_bounds__(256) fa_fwd_d16_kernel(const __half* __restrict__ Q, const __half* __restrict__ K,
                                                         const __half* __restrict__ V, float* __restrict__ O,
                                                         long long rows_total, int nq, int nkv, float scale_log2) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows_total) return;
  const long long bh = row / nq;
  float q[16];
  load_half16(Q + row * 16, q);
#pragma unroll
  for (int i = 0; i < 16; ++i) q[i] *= scale_log2;  // exp2 domain: scale * log2(e) folded into q
  const __half* kb = K + bh * nkv * 16;
  const __half* vb = V + bh * nkv * 16;
  float m = -CUDART_INF_F, l = 0.f, acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int t = lane; t < nkv; t += 32) {
    float kk[16], vv[16];
    load_half16(kb + static_cast<long long>(t) * 16, kk);
    load_half16(vb + static_cast<long long>(t) * 16, vv);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s = fmaf(q[i], kk[i], s);
    const float m_new = fmaxf(m, s);
    const float a = exp2f(m - m_new);  // first key: exp2(-inf) = 0
    const float p = exp2f(s - m_new);
    l = l * a + p;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, p * vv[i]);
    m = m_new;
  }
  // merge the 32 per-lane states (lanes without any key hold m = -inf, l = 0)
  float m_all = m;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m_all = fmaxf(m_all, __shfl_xor_sync(0xffffffffu, m_all, o));
  const float w = (m == -CUDART_INF_F) ? 0.f : exp2f(m - m_all);
  l *= w;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  const float inv = 1.f / l;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float v = acc[i] * w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    acc[i] = v * inv;
  }
  if (lane < 4)  // 16 floats = 64 B: four lanes store 16 B each
    *reinterpret_cast<float4*>(O + row * 16 + lane * 4) =
        make_float4(acc[lane * 4], acc[lane * 4 + 1], acc[lane * 4 + 2], acc[lane * 4 + 3]);
}

}  // namespace fa
