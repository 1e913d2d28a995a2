// Row softmax, fp32 in / fp32 (or fp16) out, one HBM read and one HBM write per element.
//
// Replaces softmax_naive_baseline_kernel (Softmax micro kernel/softmax_naive_baseline/
// softmax_naive_baseline.cu:56-130): one 128-thread block per row, row staged in shared memory,
// three passes and 14 tree-reduction __syncthreads.  Also covers the shapes of its siblings
// softmax_warp_shfl_kernel (N <= 32, ../softmax_warp_shfl_baseline.cu:69-115) and
// softmax_block_hybrid_kernel (N <= 512, ../softmax_block_hybrid.cu:74-169).
// Same arithmetic: p = expf(x - rowmax) / sum.
//
// This path is HBM-bound (8 bytes per element, ~45 lane-cycles of ALU budget per element on a
// B200), so it uses no tensor cores and no shared-memory staging of the row: each row lives in
// the registers of a thread group (one warp for N <= 1024, one 256-thread CTA for N <= 8192),
// loaded with 128-bit streaming loads, reduced with shuffles (+ one smem hop for a CTA), and
// written back with 128-bit streaming stores.  Longer or unaligned rows take a three-pass
// kernel whose 2nd/3rd passes hit L2.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>

namespace fa {

// Output element: float (softmax_rows) or __half (softmax_rows_f16out, the P operand of a following P*V GEMM).
__device__ __forceinline__ void store_probs4(float* p, int idx4, float4 v) { __stcs(reinterpret_cast<float4*>(p) + idx4, v); }
__device__ __forceinline__ void store_probs4(__half* p, int idx4, float4 v) {
  const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<const uint32_t*>(&lo);
  u.y = *reinterpret_cast<const uint32_t*>(&hi);
  __stcs(reinterpret_cast<uint2*>(p) + idx4, u);
}
__device__ __forceinline__ void store_prob(float* p, int i, float v) { p[i] = v; }
__device__ __forceinline__ void store_prob(__half* p, int i, float v) { p[i] = __float2half_rn(v); }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// TPR threads cooperate on one row; each caches V4 float4 (element (i*TPR + t)*4, i < V4).
// Requires N % 4 == 0, 16-byte aligned X and P, N <= TPR*V4*4.  blockDim.x = 256.
template <int TPR, int V4, typename TOut>
__global__ void __launch_bounds__(256) softmax_rows_reg_kernel(const float* X, TOut* P, int rows, int N) {
  static_assert(TPR == 32 || TPR == 256, "thread group is a warp or the CTA");
  constexpr int kGroupsPerCta = 256 / TPR;
  __shared__ float red[2][8];
  const int t = threadIdx.x % TPR;
  const int g = threadIdx.x / TPR;
  const int n4 = N >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * kGroupsPerCta;
  const float4 kNegInf4 = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
  long long row = static_cast<long long>(blockIdx.x) * kGroupsPerCta + g;
  // The next row of this thread group is already on its way from HBM while the current one is reduced and written:
  // the two reductions (shuffles, and one shared-memory hop for a CTA-wide row) would otherwise leave the load queue empty.
  float4 nxt[V4];
  if (row < rows) {
    const float4* xin = reinterpret_cast<const float4*>(X + row * N);
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      const int idx = i * TPR + t;
      nxt[i] = (idx < n4) ? __ldcs(xin + idx) : kNegInf4;
    }
  }
  for (; row < rows; row += stride) {
    TOut* pout = P + row * N;  // X == P allowed: a row is fully read before any of it is written (rows are disjoint)
    float4 v[V4];
#pragma unroll
    for (int i = 0; i < V4; ++i) v[i] = nxt[i];
    if (row + stride < rows) {
      const float4* xin = reinterpret_cast<const float4*>(X + (row + stride) * N);
#pragma unroll
      for (int i = 0; i < V4; ++i) {
        const int idx = i * TPR + t;
        nxt[i] = (idx < n4) ? __ldcs(xin + idx) : kNegInf4;
      }
    }
    float m = -CUDART_INF_F;
#pragma unroll
    for (int i = 0; i < V4; ++i) m = fmaxf(m, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
    m = warp_max(m);
    if constexpr (TPR == 256) {
      if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = m;
      __syncthreads();
      m = red[0][0];
#pragma unroll
      for (int w = 1; w < 8; ++w) m = fmaxf(m, red[0][w]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      v[i].x = expf(v[i].x - m);
      v[i].y = expf(v[i].y - m);
      v[i].z = expf(v[i].z - m);
      v[i].w = expf(v[i].w - m);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    s = warp_sum(s);
    if constexpr (TPR == 256) {
      if ((threadIdx.x & 31) == 0) red[1][threadIdx.x >> 5] = s;
      __syncthreads();
      s = red[1][0];
#pragma unroll
      for (int w = 1; w < 8; ++w) s += red[1][w];
    }
    const float inv = 1.f / s;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      const int idx = i * TPR + t;
      if (idx < n4) store_probs4(pout, idx, make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv));
    }
    if constexpr (TPR == 256) __syncthreads();  // red[] is reused by the next row
  }
}

// Any N, any alignment: one CTA per row, three passes (max, sum, write).
template <typename TOut>
__global__ void __launch_bounds__(256) softmax_rows_3pass_kernel(const float* X, TOut* P, int rows, int N) {
  __shared__ float red[8];
  __shared__ float bcast;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const float* xin = X + row * N;  // no __restrict__: in-place (X == P) is allowed
    TOut* pout = P + row * N;
    float m = -CUDART_INF_F;
    for (int i = threadIdx.x; i < N; i += 256) m = fmaxf(m, xin[i]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      float r = red[0];
      for (int w = 1; w < 8; ++w) r = fmaxf(r, red[w]);
      bcast = r;
    }
    __syncthreads();
    m = bcast;
    float s = 0.f;
    for (int i = threadIdx.x; i < N; i += 256) s += expf(xin[i] - m);
    s = warp_sum(s);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float r = red[0];
      for (int w = 1; w < 8; ++w) r += red[w];
      bcast = r;
    }
    __syncthreads();
    const float inv = 1.f / bcast;
    for (int i = threadIdx.x; i < N; i += 256) store_prob(pout, i, expf(xin[i] - m) * inv);
    __syncthreads();
  }
}

}  // namespace fa
