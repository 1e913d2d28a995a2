// C = ReLU(A*B + bias[col]) + residual on sm_100a: persistent CTAs, TMA A/B ring ->
// tcgen05.mma 128x256x16 -> double-buffered fp32 accumulators in TMEM -> tcgen05.ld -> fused
// epilogue -> global.
//
// Replaces wmma_gemm_fused_kernel (GEMM/tensorcore_gemm_fused_epilogue/tensorcore_gemm_fused_epilogue.cu:53-130):
// one warp per 16x16 C tile, WMMA fragments loaded straight from global every k-step (:80-88),
// accumulator -> 1 KB smem -> 32 lanes apply bias/ReLU/residual (:94-128).  Same operand
// contract: A fp16 [M,K] row-major (K-major UMMA operand), B fp16 [K,N] row-major (an MN-major
// UMMA operand -- no host-side transposition), bias fp32 [N] / residual fp32 [M,N] nullable,
// epilogue order bias -> ReLU -> residual (:109-126).
//
// CTA = 256 threads: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 4-7 epilogue (warp & 3 selects the TMEM lane quadrant).  While the epilogue drains
// accumulator `a`, the mainloop of the next tile fills accumulator `a ^ 1`.
#pragma once
#include "sm100_ptx.cuh"

namespace fa {

struct GemmParams {
  int M, N, K;
  int m_blocks, n_blocks;  // ceil(M/128), ceil(N/256)
  const float* bias;       // [N] or null
  const float* residual;   // [M,N] or null
  float* C;                // [M,N]
  DebugChannel dbg;
};

struct GemmCfg {
  static constexpr int kBM = 128, kBN = 256, kBK = 64;
  static constexpr int kStages = 4;
  static constexpr int kABytes = kBM * kBK * 2;            // 16 KB: one 128B-swizzled box
  static constexpr int kBBoxBytes = 64 * 128;              // 64 k-rows x 128 B (64 n columns)
  static constexpr int kBBytes = (kBN / 64) * kBBoxBytes;  // 32 KB
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOffBar = kStages * kStageBytes;
  static constexpr int kNumBars = 2 * kStages + 4;
  static constexpr int kOffTmem = kOffBar + kNumBars * 8;
  static constexpr int kSmemBytes = kOffTmem + 16;
  static constexpr int kThreads = 256;
  static constexpr int kGroupN = 8;  // rasterisation: bands of 8 n-blocks, m fastest inside a band
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget exceeded");
};

enum : uint32_t { SITE_G_EMPTY = 0x21, SITE_G_FULL, SITE_G_TEMPTY, SITE_G_TFULL };

__device__ __forceinline__ void gemm_tile_coords(int tile, const GemmParams& p, int& mb, int& nb) {
  const int band_tiles = GemmCfg::kGroupN * p.m_blocks;
  const int band = tile / band_tiles;
  const int rem = tile - band * band_tiles;
  const int n_first = band * GemmCfg::kGroupN;
  const int band_w = min(GemmCfg::kGroupN, p.n_blocks - n_first);
  mb = rem / band_w;
  nb = n_first + rem % band_w;
}

__global__ void __launch_bounds__(256, 1)
gemm_fused_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                  const GemmParams p) {
  using Cfg = GemmCfg;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  uint64_t* const full = bars;                        // [kStages] TMA -> MMA
  uint64_t* const empty = bars + Cfg::kStages;        // [kStages] MMA -> TMA
  uint64_t* const tfull = empty + Cfg::kStages;       // [2] MMA -> epilogue
  uint64_t* const tempty = tfull + 2;                 // [2] epilogue -> MMA
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::kOffTmem);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const DebugChannel dbg = p.dbg;
  const int num_tiles = p.m_blocks * p.n_blocks;
  const int num_kb = (p.K + Cfg::kBK - 1) / Cfg::kBK;

  if (warp == 1 && lane == 0) {
    if ((smem_u32(smem) & 1023u) != 0u) wait_timeout_trap(dbg, 0xA12u, 0, smem_u32(smem));
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 128);
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&mapA);
    prefetch_tensormap(&mapB);
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------- TMA producer -----------------------------------
    // whole warp walks the schedule; the lane chosen by elect.sync issues (see fa_fwd.cuh)
    {
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int mb, nb;
        gemm_tile_coords(tile, p, mb, nb);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % Cfg::kStages;
          const int par = (it / Cfg::kStages) & 1;
          mbar_wait(&empty[s], par ^ 1, dbg, SITE_G_EMPTY, it);
          uint8_t* const a_dst = smem + s * Cfg::kStageBytes;
          uint8_t* const b_dst = a_dst + Cfg::kABytes;
          if (elect_one()) {
            mbar_arrive_expect_tx(&full[s], Cfg::kStageBytes);
            tma_load_2d(a_dst, &mapA, &full[s], kb * Cfg::kBK, mb * Cfg::kBM, kEvictNormal);
#pragma unroll
            for (int b = 0; b < Cfg::kBN / 64; ++b)
              tma_load_2d(b_dst + b * Cfg::kBBoxBytes, &mapB, &full[s], nb * Cfg::kBN + b * 64, kb * Cfg::kBK,
                          kEvictNormal);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------- MMA issuer -------------------------------------
    {
      constexpr uint32_t idesc = umma_idesc_f16(Cfg::kBM, Cfg::kBN, 0, 1);
      int it = 0, local = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
        const int acc = local & 1;
        const int acc_par = (local >> 1) & 1;
        mbar_wait(&tempty[acc], acc_par ^ 1, dbg, SITE_G_TEMPTY, local);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * Cfg::kBN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % Cfg::kStages;
          const int par = (it / Cfg::kStages) & 1;
          mbar_wait(&full[s], par, dbg, SITE_G_FULL, it);
          tc_fence_after();
          const uint32_t a0 = smem_u32(smem + s * Cfg::kStageBytes);
          const uint32_t b0 = a0 + Cfg::kABytes;
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < Cfg::kBK / 16; ++ks) {
              umma_ss(d, umma_desc_sw128(a0 + ks * 32, 16, 1024),
                      umma_desc_sw128(b0 + ks * 2048, Cfg::kBBoxBytes, 1024), idesc, (kb | ks) ? 1u : 0u);
            }
            umma_commit(&empty[s]);
            if (kb == num_kb - 1) umma_commit(&tfull[acc]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4) {
    // --------------------------------- epilogue --------------------------------------
    const int wq = warp & 3;
    const int row_in_tile = wq * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(wq * 32) << 16;
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      int mb, nb;
      gemm_tile_coords(tile, p, mb, nb);
      const int acc = local & 1;
      const int acc_par = (local >> 1) & 1;
      mbar_wait(&tfull[acc], acc_par, dbg, SITE_G_TFULL, local);
      tc_fence_after();
      const int grow = mb * Cfg::kBM + row_in_tile;
      const bool row_ok = grow < p.M;
      const size_t row_off = static_cast<size_t>(grow) * p.N;
#pragma unroll 1
      for (int cc = 0; cc < Cfg::kBN / 32; ++cc) {
        uint32_t r[32];
        tmem_ld_x32(tmem_base + lane_base + acc * Cfg::kBN + cc * 32, r);
        tmem_wait_ld();
        const int col0 = nb * Cfg::kBN + cc * 32;
        if (row_ok) {
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            const int col = col0 + k;
            if (col < p.N) {  // N % 8 == 0, so a 4-wide group is entirely inside or outside
              float4 v = make_float4(__uint_as_float(r[k]), __uint_as_float(r[k + 1]),
                                     __uint_as_float(r[k + 2]), __uint_as_float(r[k + 3]));
              if (p.bias != nullptr) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
              }
              v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
              if (p.residual != nullptr) {
                const float4 q = __ldcs(reinterpret_cast<const float4*>(p.residual + row_off + col));
                v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
              }
              __stcs(reinterpret_cast<float4*>(p.C + row_off + col), v);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace fa
