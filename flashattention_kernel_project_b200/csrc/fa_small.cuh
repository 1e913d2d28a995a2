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

// two packed halfs -> two floats without taking the address of a register (which would force local memory)
__device__ __forceinline__ void half2_bits_to_floats(uint32_t w, float& lo, float& hi) {
  lo = __half2float(__ushort_as_half(static_cast<unsigned short>(w & 0xffffu)));
  hi = __half2float(__ushort_as_half(static_cast<unsigned short>(w >> 16)));
}
__device__ __forceinline__ void unpack_half16(const uint4 a, const uint4 b, float* out) {
  half2_bits_to_floats(a.x, out[0], out[1]);
  half2_bits_to_floats(a.y, out[2], out[3]);
  half2_bits_to_floats(a.z, out[4], out[5]);
  half2_bits_to_floats(a.w, out[6], out[7]);
  half2_bits_to_floats(b.x, out[8], out[9]);
  half2_bits_to_floats(b.y, out[10], out[11]);
  half2_bits_to_floats(b.z, out[12], out[13]);
  half2_bits_to_floats(b.w, out[14], out[15]);
}
__device__ __forceinline__ void load_half16(const __half* p, float* out) {
  unpack_half16(__ldg(reinterpret_cast<const uint4*>(p)), __ldg(reinterpret_cast<const uint4*>(p) + 1), out);
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
  if (lane < 4) {  // 16 floats = 64 B: four lanes store 16 B each (selects, not a dynamically indexed register array)
    float4 o4 = make_float4(acc[0], acc[1], acc[2], acc[3]);
    if (lane == 1) o4 = make_float4(acc[4], acc[5], acc[6], acc[7]);
    if (lane == 2) o4 = make_float4(acc[8], acc[9], acc[10], acc[11]);
    if (lane == 3) o4 = make_float4(acc[12], acc[13], acc[14], acc[15]);
    *reinterpret_cast<float4*>(O + row * 16 + lane * 4) = o4;
  }
}

// Same operator for Nq >= 4: a CTA of four warps takes 16 consecutive query rows of one (b,h), stages K / V in
// shared memory 128 keys at a time (one 32 B row of each per thread) and every warp keeps FOUR query rows in
// registers, so a key's K and V rows are converted once per four (row, key) pairs and the 64 partial outputs of a
// lane are reduced across the warp by recursive halving (62 shuffles for four rows instead of 80 per row).
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(128) fa_fwd_d16_rows4_kernel(const __half* __restrict__ Q, const __half* __restrict__ K,
                                                               const __half* __restrict__ V, float* __restrict__ O,
                                                               int q_blocks, int nq, int nkv, float scale_log2) {
  __shared__ uint4 sK[128][2];  // 128 keys x 16 halfs
  __shared__ uint4 sV[128][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long bh = blockIdx.x / q_blocks;
  const int row0 = (blockIdx.x % q_blocks) * 16 + warp * 4;  // first of this warp's four query rows
  const __half* kb = K + bh * nkv * 16;
  const __half* vb = V + bh * nkv * 16;
  float q[4][16];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int row = min(row0 + r, nq - 1);  // rows past nq repeat the last row and are not stored
    load_half16(Q + (bh * nq + row) * 16, q[r]);
#pragma unroll
    for (int i = 0; i < 16; ++i) q[r][i] *= scale_log2;  // exp2 domain
  }
  float m_run[4], l_run[4], acc[4][16];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    m_run[r] = -CUDART_INF_F;
    l_run[r] = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[r][i] = 0.f;
  }
  for (int k0 = 0; k0 < nkv; k0 += 128) {
    __syncthreads();  // the previous chunk has been consumed by every warp
    {
      const int key = k0 + threadIdx.x;
      if (key < nkv) {
        const uint4* kp = reinterpret_cast<const uint4*>(kb + static_cast<long long>(key) * 16);
        const uint4* vp = reinterpret_cast<const uint4*>(vb + static_cast<long long>(key) * 16);
        sK[threadIdx.x][0] = __ldg(kp);
        sK[threadIdx.x][1] = __ldg(kp + 1);
        sV[threadIdx.x][0] = __ldg(vp);
        sV[threadIdx.x][1] = __ldg(vp + 1);
      }
    }
    __syncthreads();
    float s[4][4];  // [row][key slot]: this lane's keys are k0 + lane + 32 * slot
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kl = lane + 32 * i;
      if (k0 + kl < nkv) {
        float kk[16];
        unpack_half16(sK[kl][0], sK[kl][1], kk);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          float d = 0.f;
#pragma unroll
          for (int c = 0; c < 16; ++c) d = fmaf(q[r][c], kk[c], d);
          s[r][i] = d;
        }
      } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) s[r][i] = -CUDART_INF_F;
      }
    }
    float pr[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float mx = fmaxf(fmaxf(s[r][0], s[r][1]), fmaxf(s[r][2], s[r][3]));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float m_new = fmaxf(m_run[r], mx);  // finite: every chunk holds at least one valid key
      const float alpha = ex2_fast(m_run[r] - m_new);  // first chunk: exp2(-inf) = 0
      m_run[r] = m_new;
      float ps = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        pr[r][i] = ex2_fast(s[r][i] - m_new);
        ps += pr[r][i];
      }
      l_run[r] = l_run[r] * alpha + ps;  // lane-local partial sum; reduced across the warp at the end
      if (alpha != 1.f) {
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[r][c] *= alpha;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kl = lane + 32 * i;
      if (k0 + kl < nkv) {
        float vv[16];
        unpack_half16(sV[kl][0], sV[kl][1], vv);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 16; ++c) acc[r][c] = fmaf(pr[r][i], vv[c], acc[r][c]);
      }
    }
  }
  // row sums across the warp, then the 64 partial outputs by recursive halving: after the step with lane offset
  // `o` a lane keeps the half selected by that bit of its lane id, so it ends with two adjacent columns of one row
#pragma unroll
  for (int r = 0; r < 4; ++r) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], o);
  }
  float v64[64];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 16; ++c) v64[r * 16 + c] = acc[r][c];
  int n = 32;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1, n >>= 1) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j < n) {
        const float send = upper ? v64[j] : v64[j + n];
        const float keep = upper ? v64[j + n] : v64[j];
        v64[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
  }
  const int idx = ((lane >> 4) & 1) * 32 + ((lane >> 3) & 1) * 16 + ((lane >> 2) & 1) * 8 + ((lane >> 1) & 1) * 4 + (lane & 1) * 2;
  const int r = idx >> 4, c = idx & 15;
  if (row0 + r < nq) {
    float lr = l_run[0];  // select without dynamic register indexing
    lr = (r == 1) ? l_run[1] : lr;
    lr = (r == 2) ? l_run[2] : lr;
    lr = (r == 3) ? l_run[3] : lr;
    const float inv = 1.f / lr;
    *reinterpret_cast<float2*>(O + (bh * nq + row0 + r) * 16 + c) = make_float2(v64[0] * inv, v64[1] * inv);
  }
}

}  // namespace fa
