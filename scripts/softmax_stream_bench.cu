// Developer microbenchmark: the softmax warps' per-tile instruction stream of fa_fwd_kernel (row max over 128 scores,
// speculative head, exp2 on MUFU / FMA pipes per FA_POLY_MASK, packed row sums, fp16 packing) on REGISTER data, without
// TMEM, barriers or the tensor pipe: 1 or 2 warps per sub-partition, cycles per 128-key tile.  It answers how much of the
// kernel's ~3200-cycle period per tile pair is the instruction mix itself (what ptxas makes of it) and how much is
// the coupling to the MMA / TMEM side.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I<csrc> ...
#include <cstdio>
#include "sm100_ptx.cuh"
using namespace fa;

#ifndef POLY_MASK
#define POLY_MASK 0x11
#endif
#ifndef SPEC
#define SPEC 16
#endif
#ifndef WITH_MAX
#define WITH_MAX 1
#endif
#ifndef WITH_SUM
#define WITH_SUM 1
#endif

constexpr int kTiles = 256;

__device__ __forceinline__ float tile_stream(float* s, float c, float& m_used, float& l, uint32_t* sink) {
  const uint64_t c2 = pack_f32x2(c, c);
  float neg_mc = -m_used * c;
  uint64_t neg_mc2 = pack_f32x2(neg_mc, neg_mc);
  float pe[SPEC > 0 ? SPEC : 4];
#pragma unroll
  for (int k = 0; k < SPEC; k += 2) {
    float x0, x1;
    unpack_f32x2(ffma2(pack_f32x2(s[k], s[k + 1]), c2, neg_mc2), x0, x1);
    pe[k] = ex2_approx(x0);
    pe[k + 1] = ex2_approx(x1);
  }
  float m_new = m_used;
#if WITH_MAX == 2  // 2-input max chains
  float mx0 = s[0], mx1 = s[32], mx2 = s[64], mx3 = s[96];
#pragma unroll
  for (int k = 1; k < 32; ++k) {
    mx0 = fmaxf(mx0, s[k]); mx1 = fmaxf(mx1, s[32 + k]); mx2 = fmaxf(mx2, s[64 + k]); mx3 = fmaxf(mx3, s[96 + k]);
  }
  m_new = fmaxf(fmaxf(mx0, mx1), fmaxf(fmaxf(mx2, mx3), m_used));
#elif WITH_MAX == 3  // eight independent 3-input chains (more ILP)
  float mx[8];
#pragma unroll
  for (int c8 = 0; c8 < 8; ++c8) {
    mx[c8] = fmax3(s[c8 * 16], s[c8 * 16 + 1], s[c8 * 16 + 2]);
#pragma unroll
    for (int k = 3; k < 15; k += 2) mx[c8] = fmax3(mx[c8], s[c8 * 16 + k], s[c8 * 16 + k + 1]);
    mx[c8] = fmaxf(mx[c8], s[c8 * 16 + 15]);
  }
  m_new = fmaxf(fmax3(fmax3(mx[0], mx[1], mx[2]), fmax3(mx[3], mx[4], mx[5]), fmaxf(mx[6], mx[7])), m_used);
#elif WITH_MAX
  float mx0 = s[0], mx1 = s[32], mx2 = s[64], mx3 = s[96];
#pragma unroll
  for (int k = 1; k < 31; k += 2) {
    mx0 = fmax3(mx0, s[k], s[k + 1]);
    mx1 = fmax3(mx1, s[32 + k], s[32 + k + 1]);
    mx2 = fmax3(mx2, s[64 + k], s[64 + k + 1]);
    mx3 = fmax3(mx3, s[96 + k], s[96 + k + 1]);
  }
  mx0 = fmax3(mx0, s[31], s[63]);
  mx2 = fmax3(mx2, s[95], s[127]);
  m_new = fmaxf(fmax3(mx0, mx1, mx2), fmaxf(mx3, m_used));
#endif
  const bool need = (m_new - m_used) * c > 8.f;
  if (__any_sync(0xffffffffu, need)) {
    if (need) {
      l *= ex2_approx((m_used - m_new) * c);
      m_used = m_new;
      neg_mc = -m_used * c;
      neg_mc2 = pack_f32x2(neg_mc, neg_mc);
#pragma unroll
      for (int k = 0; k < SPEC; k += 2) {
        float x0, x1;
        unpack_f32x2(ffma2(pack_f32x2(s[k], s[k + 1]), c2, neg_mc2), x0, x1);
        pe[k] = ex2_approx(x0);
        pe[k + 1] = ex2_approx(x1);
      }
    }
  }
  uint64_t lsum_a = pack_f32x2(0.f, 0.f), lsum_b = pack_f32x2(0.f, 0.f);
  const float cy = c * (1.f / 128.f);
  const float by = (neg_mc + 120.f) * (1.f / 128.f);
  uint32_t acc = 0;
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    uint32_t pk[8];
#pragma unroll
    for (int k = 0; k < 16; k += 4) {
      const int e = g * 16 + k;
      float p0, p1, p2, p3;
      if (e < SPEC) {
        p0 = pe[e]; p1 = pe[e + 1]; p2 = pe[e + 2]; p3 = pe[e + 3];
      } else {
        float x0, x1, x2, x3;
        if ((POLY_MASK >> (k / 2)) & 1u) {
          exp2_poly_pair(pack_f32x2(ffma_sat(s[e + 0], cy, by), ffma_sat(s[e + 1], cy, by)), p0, p1);
        } else {
          unpack_f32x2(ffma2(pack_f32x2(s[e + 0], s[e + 1]), c2, neg_mc2), x0, x1);
          p0 = ex2_approx(x0); p1 = ex2_approx(x1);
        }
        if ((POLY_MASK >> (k / 2 + 1)) & 1u) {
          exp2_poly_pair(pack_f32x2(ffma_sat(s[e + 2], cy, by), ffma_sat(s[e + 3], cy, by)), p2, p3);
        } else {
          unpack_f32x2(ffma2(pack_f32x2(s[e + 2], s[e + 3]), c2, neg_mc2), x2, x3);
          p2 = ex2_approx(x2); p3 = ex2_approx(x3);
        }
      }
#if WITH_SUM
      lsum_a = fadd2(lsum_a, pack_f32x2(p0, p1));
      lsum_b = fadd2(lsum_b, pack_f32x2(p2, p3));
#endif
      pk[k / 2 + 0] = pack_half2(p0, p1);
      pk[k / 2 + 1] = pack_half2(p2, p3);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) acc ^= pk[q];  // stands in for the tcgen05.st of 8 packed columns
  }
  float a0, a1;
  unpack_f32x2(fadd2(lsum_a, lsum_b), a0, a1);
  l += a0 + a1;
  *sink ^= acc;
  return l;
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k(const float* in, float* out, long long* cycles) {
  float s0[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) s0[i] = in[(threadIdx.x * 128 + i) & 4095];
  float m_used = -INFINITY, l = 0.f;
  uint32_t sink = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int t = 0; t < kTiles; ++t) {
    float s[128];
    const float d = __uint_as_float((sink & 1u) | 0x3f800000u) - 1.f + t * 1e-3f;  // data-dependent: nothing hoists
#pragma unroll
    for (int i = 0; i < 128; ++i) s[i] = s0[i] + d;
    tile_stream(s, 0.1275f, m_used, l, &sink);
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned long long*>(cycles), static_cast<unsigned long long>(t1 - t0));
  out[blockIdx.x * blockDim.x + threadIdx.x] = l + m_used + __uint_as_float(sink);
}

template <int WARPS>
void run(const char* name) {
  float *in, *out;
  long long* cyc;
  cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
  float h[4096];
  for (int i = 0; i < 4096; ++i) h[i] = ((i * 2654435761u) >> 8 & 0xffff) / 65536.f * 8.f - 4.f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  long long c = 0;
  for (int rep = 0; rep < 2; ++rep) { cudaMemset(cyc, 0, 8); k<WARPS><<<148, WARPS * 32>>>(in, out, cyc); }
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  const double per_tile = static_cast<double>(c) / kTiles;
  printf("%-34s poly 0x%02x spec %2d max %d sum %d: %8.1f cycles per 128-key tile and warp = %5.2f per key; "
         "per sub-partition %5.2f cycles per key  %s\n", name, POLY_MASK, SPEC, WITH_MAX, WITH_SUM, per_tile, per_tile / 128,
         per_tile / 128 / (WARPS / 4), cudaGetErrorString(e));
}

int main() {
  run<4>("1 warp / sub-partition");
  run<8>("2 warps / sub-partition");
  run<16>("4 warps / sub-partition");
  return 0;
}
