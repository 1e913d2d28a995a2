// C = act(alpha * A*op(B) + bias[col]) + residual on sm_100a: persistent CTA PAIRS (cluster of 2), TMA A/B ring ->
// tcgen05.mma.cta_group::2 256x256x16 -> double-buffered fp32 accumulators in TMEM -> tcgen05.ld ->
// epilogue combined in shared memory with TMA-fetched residual blocks -> TMA store.  Batched (third TMA coordinate),
// B either [K,N] (MN-major operand) or [N,K] (K-major operand, C = A*B^T).
//
// Replaces wmma_gemm_fused_kernel (GEMM/tensorcore_gemm_fused_epilogue/tensorcore_gemm_fused_epilogue.cu:53-130):
// one warp per 16x16 C tile, WMMA fragments loaded straight from global every k-step (:80-88),
// accumulator -> 1 KB smem -> 32 lanes apply bias/ReLU/residual (:94-128).  Same operand
// contract: A fp16 [M,K] row-major (K-major UMMA operand), B fp16 [K,N] row-major (an MN-major
// UMMA operand -- no host-side transposition), bias fp32 [N] / residual fp32 [M,N] nullable,
// epilogue order bias -> ReLU -> residual (:109-126).
//
// Two CTAs on the two SMs of a TPC work on one 256x256 output tile.  CTA r of the pair owns rows
// [128r, 128r+128): it loads its own 128x64 A block and HALF of the 64x256 B block (columns
// [128r, 128r+128)) per k-step, and its TMEM holds its 128x256 half of the accumulator.  The leader
// CTA (rank 0) issues tcgen05.mma.cta_group::2 (M = 256, N = 256): the hardware reads A from both
// CTAs' shared memory and assembles B from the two halves, so each SM fetches 32 KB per k-block from L2
// instead of 48 KB and reads 64 B/clk of operands from shared memory instead of 96 B/clk.
//
// CTA = 256 threads: warp 0 TMA producer, warp 1 MMA issuer (leader CTA only), warp 2 TMEM allocator,
// warps 4-7 epilogue (warp & 3 selects the TMEM lane quadrant): accumulator -> registers -> bias/ReLU -> combined
// in shared memory with a residual block fetched by TMA -> TMA store.  While the epilogues drain accumulator
// `a`, the mainloop of the next tile fills accumulator `a ^ 1`.
// Cross-CTA signalling: both producers' TMA loads complete_tx on the LEADER's full barrier
// (cp.async.bulk.tensor ... cta_group::2 with the peer bit of the barrier address cleared); the leader's
// tcgen05.commit multicasts "slot free" / "accumulator full" to both CTAs; the epilogue warps of both
// CTAs arrive remotely on the leader's "accumulator empty" barrier.
#pragma once
#include "sm100_ptx.cuh"

namespace fa {

struct GemmParams {
  int M, N, K, batch;
  int m_blocks, n_blocks;  // ceil(M/256), ceil(N/256): tiles of the CTA pair, per batch entry
  float alpha;             // C = act(alpha * A*op(B) + bias) + residual
  int act;                 // GemmAct applied between bias and residual (1 = ReLU: the reference's fused-epilogue contract)
  const float* bias;       // [N] or null
  const float* residual;   // [M,N] or null
  float* C;                // [M,N]
  DebugChannel dbg;
};

// Epilogue activations: the reference has ReLU only (tensorcore_gemm_fused_epilogue.cu:117, gemm_fused_epilogue.cu:35-37);
// GELU (exact, erf form) and SiLU complete the surface.  Values match FA_ACT_* in fa_b200.h.
enum GemmAct : int { GEMM_ACT_IDENTITY = 0, GEMM_ACT_RELU = 1, GEMM_ACT_GELU = 2, GEMM_ACT_SILU = 3 };

__device__ __forceinline__ float gemm_act_gelu(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gemm_act_silu(float x) { return x / (1.f + expf(-x)); }

// Build knobs of the tile walk and the L2 policy of the operand loads (scripts/sustained_ab_gemm.py ranks them).
// A wave of 74 CTA pairs covers ~9 m-blocks x kGroupN n-blocks; the next wave moves down the SAME n-band, so the B
// panels of a band are re-read wave after wave while every A panel is used by one wave per band: B is worth keeping in
// L2 (evict-last), A is a stream (evict-first).
#ifndef FA_GEMM_GROUP_N
#define FA_GEMM_GROUP_N 8
#endif
#ifndef FA_GEMM_HINT_A
#define FA_GEMM_HINT_A kEvictNormal
#endif
#ifndef FA_GEMM_HINT_B
#define FA_GEMM_HINT_B kEvictNormal
#endif

// kSplit = false: fp16 operands, one tcgen05.mma pass per k-block (the tensor-core GEMM of the reference).
// kSplit = true : fp32 operands as tf32 (hi, lo) PAIRS -- "3xTF32": A*B ~ A_lo*B_hi + A_hi*B_lo + A_hi*B_hi recovers
//                 fp32-level products from the tf32 tensor path (the reference's fp32 sibling gemm_tiled_fused_kernel,
//                 GEMM/gemm_fused_epilogue/gemm_fused_epilogue.cu:44-111, is a scalar-FMA kernel).  Four operand tiles
//                 per stage (A_hi, A_lo, B_hi, B_lo), three MMA passes, both operands K-major (B arrives transposed).
template <bool kSplit>
struct GemmCfgT {
  static constexpr int kBM = 128;   // rows per CTA (256 per pair)
  static constexpr int kBN = 256;
  static constexpr int kElemBytes = kSplit ? 4 : 2;
  static constexpr int kBK = 128 / kElemBytes;             // one 128-byte swizzle row of K per k-block: 64 fp16 / 32 fp32
  static constexpr int kKPerMma = 32 / kElemBytes;         // 16 (f16) / 8 (tf32)
  static constexpr int kStages = kSplit ? 3 : 5;
  static constexpr int kABytes = kBM * 128;                // 16 KB: one 128B-swizzled box
  static constexpr int kBBoxBytes = 64 * 128;              // fp16 [K,N] form: 64 k-rows x 128 B (64 n columns)
  static constexpr int kBBytes = (kBN / 2) * 128;          // this CTA's half of B: 16 KB
  static constexpr int kOperandSets = kSplit ? 2 : 1;      // (hi, lo)
  static constexpr int kStageBytes = kOperandSets * (kABytes + kBBytes);  // 32 KB / 64 KB per CTA
  // epilogue ring: kEpiBufs buffers of 128 rows x 32 fp32 columns (one 128B-swizzled TMA box each).  A buffer
  // receives a residual block by TMA, is updated in place with the accumulator, and leaves by TMA store.
  static constexpr int kEpiCols = 32;
  static constexpr int kEpiBufs = kSplit ? 2 : 4;
  static constexpr int kEpiBufBytes = kBM * kEpiCols * 4;  // 16 KB
  static constexpr int kChunksPerTile = kBN / kEpiCols;    // 8
  static constexpr int kOffBar = kStages * kStageBytes;
  static constexpr int kNumBars = 2 * kStages + 4 + kEpiBufs;
  static constexpr int kOffTmem = kOffBar + kNumBars * 8;
  static constexpr int kOffEpi = (kOffTmem + 16 + 1023) / 1024 * 1024;
  static constexpr int kSmemBytes = kOffEpi + kEpiBufs * kEpiBufBytes;
  static constexpr int kThreads = 256;
  static constexpr int kGroupN = FA_GEMM_GROUP_N;  // rasterisation: bands of n-blocks, n fastest inside a band
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget exceeded");
};
using GemmCfg = GemmCfgT<false>;

enum : uint32_t { SITE_G_EMPTY = 0x21, SITE_G_FULL, SITE_G_TEMPTY, SITE_G_TFULL, SITE_G_RFULL };

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> rank 0 of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes are credited to the barrier at the same offset in the LEADER CTA
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                int c2, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0),
      "r"(c1), "r"(c2), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void umma_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_ss_2sm_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::tf32 instruction descriptor: as umma_idesc_f16 with A / B format 2 (TF32); both operands K-major
__host__ __device__ constexpr uint32_t umma_idesc_tf32(uint32_t M, uint32_t N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// arrive (once all prior tcgen05 ops of this thread completed) on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_2sm_mc(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void gemm_tile_coords(int tile, const GemmParams& p, int& bz, int& mb, int& nb) {
  const int per_batch = p.m_blocks * p.n_blocks;
  bz = tile / per_batch;
  tile -= bz * per_batch;
  const int band_tiles = GemmCfg::kGroupN * p.m_blocks;
  const int band = tile / band_tiles;
  const int rem = tile - band * band_tiles;
  const int n_first = band * GemmCfg::kGroupN;
  const int band_w = min(GemmCfg::kGroupN, p.n_blocks - n_first);
  mb = rem / band_w;
  nb = n_first + rem % band_w;
}

// kTransB = false: B is [K,N] row-major (an MN-major operand, two 64-column boxes per CTA and k-block);
// kTransB = true:  B is [N,K] row-major, C = A * B^T (a K-major operand like A, one 128-row box).
// kSplit (see GemmCfgT): mapA / mapB are the tf32 HI parts, mapA2 / mapB2 the LO parts (fp32 [M,K] and [N,K]).
template <bool kTransB, bool kSplit = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
gemm_fused_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                  const __grid_constant__ CUtensorMap mapR, const __grid_constant__ CUtensorMap mapC,
                  const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB2,
                  const GemmParams p) {
  static_assert(!kSplit || kTransB, "the split-fp32 path takes B transposed (K-major)");
  using Cfg = GemmCfgT<kSplit>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  uint64_t* const full = bars;                        // [kStages] TMA (both CTAs) -> MMA; used in the leader only
  uint64_t* const empty = bars + Cfg::kStages;        // [kStages] MMA -> TMA, multicast to both CTAs
  uint64_t* const tfull = empty + Cfg::kStages;       // [2] MMA -> epilogue, multicast to both CTAs
  uint64_t* const tempty = tfull + 2;                 // [2] epilogues of both CTAs -> MMA; used in the leader only
  uint64_t* const rfull = tempty + 2;                 // [kEpiBufs] residual block landed (CTA-local)
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::kOffTmem);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();            // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const DebugChannel dbg = p.dbg;
  const int num_tiles = p.batch * p.m_blocks * p.n_blocks;
  const int num_kb = (p.K + Cfg::kBK - 1) / Cfg::kBK;

  if (warp == 1 && lane == 0) {
    if ((smem_u32(smem) & 1023u) != 0u) wait_timeout_trap(dbg, 0xA12u, 0, smem_u32(smem));
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full[s], 1);   // the leader's producer arrives once and expects the bytes of both CTAs
      mbar_init(&empty[s], 1);  // one multicast commit from the leader's MMA warp
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 8);  // 4 epilogue warps x 2 CTAs
    }
    for (int b = 0; b < Cfg::kEpiBufs; ++b) mbar_init(&rfull[b], 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&mapA);
    prefetch_tensormap(&mapB);
    prefetch_tensormap(&mapR);
    prefetch_tensormap(&mapC);
    if constexpr (kSplit) {
      prefetch_tensormap(&mapA2);
      prefetch_tensormap(&mapB2);
    }
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_slot, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();  // barrier inits of both CTAs visible before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------- TMA producer (both CTAs) -------------------------
    // whole warp walks the schedule; the lane chosen by elect.sync issues (see fa_fwd.cuh)
    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      int bz, mb, nb;
      gemm_tile_coords(tile, p, bz, mb, nb);
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % Cfg::kStages;
        const int par = (it / Cfg::kStages) & 1;
        mbar_wait(&empty[s], par ^ 1, dbg, SITE_G_EMPTY, it);
        uint8_t* const a_dst = smem + s * Cfg::kStageBytes;
        uint8_t* const b_dst = a_dst + Cfg::kABytes;
        if (elect_one()) {
          if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * Cfg::kStageBytes);
          tma_load_3d_2sm(a_dst, &mapA, &full[s], kb * Cfg::kBK, mb * 256 + static_cast<int>(rank) * Cfg::kBM, bz,
                          FA_GEMM_HINT_A);
          const int n0 = nb * Cfg::kBN + static_cast<int>(rank) * (Cfg::kBN / 2);  // this CTA's half of the B block
          if constexpr (kSplit) {  // stage = A_hi | B_hi | A_lo | B_lo
            tma_load_3d_2sm(b_dst, &mapB, &full[s], kb * Cfg::kBK, n0, bz, FA_GEMM_HINT_B);
            tma_load_3d_2sm(b_dst + Cfg::kBBytes, &mapA2, &full[s], kb * Cfg::kBK,
                            mb * 256 + static_cast<int>(rank) * Cfg::kBM, bz, FA_GEMM_HINT_A);
            tma_load_3d_2sm(b_dst + Cfg::kBBytes + Cfg::kABytes, &mapB2, &full[s], kb * Cfg::kBK, n0, bz, FA_GEMM_HINT_B);
          } else if constexpr (kTransB) {
            tma_load_3d_2sm(b_dst, &mapB, &full[s], kb * Cfg::kBK, n0, bz, FA_GEMM_HINT_B);
          } else {
#pragma unroll
            for (int b = 0; b < Cfg::kBN / 2 / 64; ++b)
              tma_load_3d_2sm(b_dst + b * Cfg::kBBoxBytes, &mapB, &full[s], n0 + b * 64, kb * Cfg::kBK, bz, FA_GEMM_HINT_B);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // -------------------------------- MMA issuer (leader CTA) -------------------------
    if (rank == 0) {
      constexpr uint32_t idesc = kSplit ? umma_idesc_tf32(256, Cfg::kBN) : umma_idesc_f16(256, Cfg::kBN, 0, kTransB ? 0 : 1);
      int it = 0, local = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++local) {
        // fp16: two accumulators alternate (the epilogue of tile t overlaps the mainloop of t + 1).  Split fp32: ONE
        // accumulator pair -- columns [0,256) for A_hi*B_hi, [256,512) for the two small cross terms, so that the
        // truncating fp32 accumulation of the tensor core sees K/8 steps on the large sum instead of 3K/8
        const int acc = kSplit ? 0 : (local & 1);
        const int acc_par = kSplit ? (local & 1) : ((local >> 1) & 1);
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
            if constexpr (kSplit) {
              // small terms first: A_lo*B_hi, A_hi*B_lo, then A_hi*B_hi -- all into the same fp32 accumulator
              const uint32_t a_lo = b0 + Cfg::kBBytes, b_lo = a_lo + Cfg::kABytes;
              const uint32_t pa[3] = {a_lo, a0, a0}, pb[3] = {b0, b_lo, b0};
#pragma unroll
              for (int pass = 0; pass < 3; ++pass)
#pragma unroll
                for (int ks = 0; ks < Cfg::kBK / Cfg::kKPerMma; ++ks)
                  umma_ss_2sm_tf32(pass < 2 ? d + Cfg::kBN : d, umma_desc_sw128(pa[pass] + ks * 32, 16, 1024),
                                   umma_desc_sw128(pb[pass] + ks * 32, 16, 1024), idesc,
                                   pass < 2 ? ((kb | ks | pass) ? 1u : 0u) : ((kb | ks) ? 1u : 0u));
            } else {
#pragma unroll
              for (int ks = 0; ks < Cfg::kBK / Cfg::kKPerMma; ++ks) {
                const uint64_t db = kTransB ? umma_desc_sw128(b0 + ks * 32, 16, 1024)
                                            : umma_desc_sw128(b0 + ks * 2048, Cfg::kBBoxBytes, 1024);
                umma_ss_2sm(d, umma_desc_sw128(a0 + ks * 32, 16, 1024), db, idesc, (kb | ks) ? 1u : 0u);
              }
            }
            umma_commit_2sm_mc(&empty[s]);
            if (kb == num_kb - 1) umma_commit_2sm_mc(&tfull[acc]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4) {
    // --------------------------------- epilogue (both CTAs) ---------------------------
    // A lane owns one accumulator row (tcgen05.ld 32x32b), which is the worst shape for global memory, and four
    // warps cannot keep enough loads in flight to cover DRAM latency.  So TMA moves the data: the residual block of
    // chunk g+3 is in flight while chunk g is combined IN PLACE in shared memory (128B swizzle: the 8 lanes of a
    // quarter-warp hit 8 different 16-byte columns, no bank conflicts) and chunk g-1 drains through a TMA store.
    const int wq = warp & 3;
    const int row_in_tile = wq * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(wq * 32) << 16;
    const bool chief = threadIdx.x == 128;
    const bool has_res = p.residual != nullptr;
    uint8_t* const ebuf = smem + Cfg::kOffEpi;
    const int my_tiles = (num_tiles - pair + num_pairs - 1) / num_pairs;
    const int total_chunks = my_tiles * Cfg::kChunksPerTile;
    const int row0 = static_cast<int>(rank) * Cfg::kBM;
    auto load_residual = [&](int g) {
      int bz, mb, nb;
      gemm_tile_coords(pair + (g / Cfg::kChunksPerTile) * num_pairs, p, bz, mb, nb);
      const int b = g % Cfg::kEpiBufs;
      mbar_arrive_expect_tx(&rfull[b], Cfg::kEpiBufBytes);
      tma_load_3d(ebuf + b * Cfg::kEpiBufBytes, &mapR, &rfull[b], nb * Cfg::kBN + (g % Cfg::kChunksPerTile) * Cfg::kEpiCols,
                  mb * 256 + row0, bz, kEvictFirst);
    };
    if (chief && has_res) {
      for (int g = 0; g < Cfg::kEpiBufs - 1 && g < total_chunks; ++g) load_residual(g);
    }
    int local = 0, g = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++local) {
      int bz, mb, nb;
      gemm_tile_coords(tile, p, bz, mb, nb);
      const int acc = kSplit ? 0 : (local & 1);
      const int acc_par = kSplit ? (local & 1) : ((local >> 1) & 1);
      mbar_wait(&tfull[acc], acc_par, dbg, SITE_G_TFULL, local);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < Cfg::kChunksPerTile; ++cc, ++g) {
        const int b = g % Cfg::kEpiBufs;
        uint32_t r[32];
        tmem_ld_x32(tmem_base + lane_base + acc * Cfg::kBN + cc * Cfg::kEpiCols, r);
        uint32_t r2[kSplit ? 32 : 1];
        if constexpr (kSplit) tmem_ld_x32(tmem_base + lane_base + Cfg::kBN + cc * Cfg::kEpiCols, r2);  // the cross terms
        if (has_res) mbar_wait(&rfull[b], (g / Cfg::kEpiBufs) & 1, dbg, SITE_G_RFULL, g);
        tmem_wait_ld();
        if (cc == Cfg::kChunksPerTile - 1) {  // accumulator drained into registers: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&tempty[acc]);
        }
        uint8_t* const rowp = ebuf + b * Cfg::kEpiBufBytes + row_in_tile * 128;
        const int col0 = nb * Cfg::kBN + cc * Cfg::kEpiCols;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if constexpr (kSplit) {  // large sum + the two small cross terms, rounded to nearest once
#pragma unroll
            for (int e = 0; e < 4; ++e)
              r[4 * q + e] = __float_as_uint(__uint_as_float(r[4 * q + e]) + __uint_as_float(r2[4 * q + e]));
          }
          float4 v = make_float4(__uint_as_float(r[4 * q]) * p.alpha, __uint_as_float(r[4 * q + 1]) * p.alpha,
                                 __uint_as_float(r[4 * q + 2]) * p.alpha, __uint_as_float(r[4 * q + 3]) * p.alpha);
          const int col = col0 + q * 4;
          if (p.bias != nullptr && col < p.N) {  // N % 8 == 0, so a 4-wide group is entirely inside or outside
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
            v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
          }
          if (p.act == GEMM_ACT_RELU) {
            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
          } else if (p.act == GEMM_ACT_GELU) {
            v.x = gemm_act_gelu(v.x); v.y = gemm_act_gelu(v.y); v.z = gemm_act_gelu(v.z); v.w = gemm_act_gelu(v.w);
          } else if (p.act == GEMM_ACT_SILU) {
            v.x = gemm_act_silu(v.x); v.y = gemm_act_silu(v.y); v.z = gemm_act_silu(v.z); v.w = gemm_act_silu(v.w);
          }
          float4* const slot = reinterpret_cast<float4*>(rowp + ((q ^ (row_in_tile & 7)) << 4));
          if (has_res) {
            const float4 q4 = *slot;
            v.x += q4.x; v.y += q4.y; v.z += q4.z; v.w += q4.w;
          }
          *slot = v;
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (chief) {
          tma_store_3d(&mapC, ebuf + b * Cfg::kEpiBufBytes, col0, mb * 256 + row0, bz);  // rows >= M, cols >= N are clipped
          tma_store_commit();
          tma_store_wait_read<1>();  // the store of chunk g-1 has left its buffer, which chunk g+3 reuses
          if (has_res && g + Cfg::kEpiBufs - 1 < total_chunks) load_residual(g + Cfg::kEpiBufs - 1);
        }
      }
    }
    if (chief) tma_store_wait_all<0>();
  }

  tc_fence_before();
  cluster_sync_all();  // no CTA of the pair may exit (or free TMEM) while the other still signals it
  tc_fence_after();
  if (warp == 2) tmem_dealloc_2sm(tmem_base, 512);
}

// ---- pre-pass of the split-fp32 path: x = hi + lo with hi = RN_tf32(x) (low 13 mantissa bits zero, so whatever
// rounding the tensor core applies to its 32-bit inputs leaves hi untouched) and lo = x - hi (exact in fp32).
__device__ __forceinline__ void tf32_split(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  lo = x - hi;
}
// A [rows, cols] -> hi, lo of the same shape; cols % 4 == 0, 16-byte aligned
__global__ void __launch_bounds__(256) gemm_split_kernel(const float4* __restrict__ x, float4* __restrict__ hi,
                                                         float4* __restrict__ lo, size_t n4) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 v = __ldcs(x + i);
    float4 h, l;
    tf32_split(v.x, h.x, l.x);
    tf32_split(v.y, h.y, l.y);
    tf32_split(v.z, h.z, l.z);
    tf32_split(v.w, h.w, l.w);
    hi[i] = h;
    lo[i] = l;
  }
}
// B [K, N] row-major -> hi, lo TRANSPOSED to [N, K] (the K-major operand the tensor core reads like A);
// 32 x 32 tiles through shared memory, block (32, 8)
__global__ void __launch_bounds__(256) gemm_split_transpose_kernel(const float* __restrict__ b, float* __restrict__ hi,
                                                                   float* __restrict__ lo, int K, int N) {
  __shared__ float th[32][33], tl[32][33];
  const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int k = k0 + r, n = n0 + threadIdx.x;
    float h = 0.f, l = 0.f;
    if (k < K && n < N) tf32_split(__ldcs(b + static_cast<size_t>(k) * N + n), h, l);
    th[r][threadIdx.x] = h;
    tl[r][threadIdx.x] = l;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int n = n0 + r, k = k0 + threadIdx.x;
    if (n < N && k < K) {
      hi[static_cast<size_t>(n) * K + k] = th[threadIdx.x][r];
      lo[static_cast<size_t>(n) * K + k] = tl[threadIdx.x][r];
    }
  }
}

}  // namespace fa
