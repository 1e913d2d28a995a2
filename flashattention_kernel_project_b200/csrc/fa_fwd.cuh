// Non-causal FlashAttention forward for sm_100a: TMA -> swizzled smem ring -> tcgen05.mma with
// S and O accumulators in TMEM -> one-thread-per-row exp2-domain online softmax read out of TMEM
// -> lazy O rescale in a correction warpgroup -> normalise -> TMA store.
//
// Replaces the reference's general tensor-core kernels
//   flashattn_forward_wmma_v5_cp_async_kernel
//     (FlashAttention/flashattn_forward_memory_bound/flashattn_forward_wmma_v5_cp_async.cu:99-459)
//   flashattn_forward_wmma_v4_kernel (FlashAttention/flashattn_forward_wmma/flashattn_forward_wmma_v4.cu:52-413)
// which use 32x32 WMMA tiles, cp.async.cg 2-stage K/V ping-pong and 8 + D/32 __syncthreads per
// 32-key tile.  Mapping of subsystems:
//   K/V loader  : cp.async + wait_group 0 + __syncthreads (v5:229-256)  -> one thread issuing
//                 cp.async.bulk.tensor into a STAGES-deep 128B-swizzled ring, full/empty mbarriers
//   contractions: wmma 16x16x16 via smem round trips (v4:197-234,339-400) -> tcgen05.mma
//                 128x128x16 (QK^T, both operands smem) and 128xDx16 (PV, P read from TMEM, V
//                 as an MN-major smem operand so no transposition is needed)
//   softmax     : warp-shuffle row reductions / lanes<32 scalar loop (v4:237-300, v5:312-348) ->
//                 tcgen05.ld 32x32b gives each thread one whole row: no shuffles at all
//   rescale     : O_smem *= alpha every tile (v5:352-360) -> only when the running max grows by
//                 more than 2^8 (lazy), done by a dedicated correction warpgroup on TMEM
//   epilogue    : scalar fp32 stores (v4:403-412) -> 1/l in registers, fp16, swizzled smem, TMA store; fp32 output
//                 and the carried (O_acc) state of a ring step: TMA as well, staged in the idle K/V ring
//   grid        : one CTA per (query block, b*h) (v4:452) -> the same by default; grids that fill less than half the
//                 SMs are cut along K/V into equal spans whose parts are merged by the last CTA to arrive (FaSegment)
//
// CTA = 512 threads = 4 warpgroups, one CTA per SM, 256 query rows per CTA as two 128-row tiles
// that ping-pong on the tensor pipe:
//   WG0 (warps 0-3)   softmax of Q tile 0        WG1 (warps 4-7)  softmax of Q tile 1
//   WG2 (warps 8-11)  correction + epilogue      WG3: warp 12 = MMA issuer (one elected lane),
//                                                     warp 13 = TMA producer,
//                                                     warps 14-15 = spare; in a ring step they pull the NEXT
//                                                     step's K / V shard from its owner GPU (peer loads)
// TMEM (512 columns x 128 lanes x 32 bit): S0 @0, S1 @128, O0 @256, O1 @256+D.  P (fp16, two per
// column) overwrites the first 64 columns of its own S.
#pragma once
#include "sm100_ptx.cuh"

namespace fa {

enum FaMode : int { FA_MODE_F16 = 0, FA_MODE_F32 = 1, FA_MODE_STATE = 2 };

struct FaParams {
  int nq;            // query rows per (b,h)
  int nkv;           // key/value rows per (b,h)
  int q_blocks;      // ceil(nq / 256)
  float scale_log2;  // scale * log2(e)
  int mode;          // FaMode
  int is_first;      // FA_MODE_STATE: incoming state is empty
  int is_last;       // FA_MODE_STATE: also emit normalised fp16 through mapO
  float* o_f32;      // FA_MODE_F32: [bh, nq, D] normalised;  FA_MODE_STATE: un-normalised O_acc
  float* m_state;    // FA_MODE_STATE: [bh, nq]  running max of scale*s (natural-log domain)
  float* l_state;    // FA_MODE_STATE: [bh, nq]
  // Work decomposition (see FaSegment): an ITEM is one 256-row query block of one (b,h) against all K/V tiles.
  // CTAs [0, items_full) own one whole item each; the remaining `tail_items` items -- the partial last wave --
  // are cut along K/V into `tail_ctas` equal spans of key tiles, one per CTA, and merged through `ws`.
  int items_full;
  int tail_items;
  int tail_ctas;
  int tail_rows;           // 1: the tail is cut along QUERY ROWS instead -- two CTAs per item, one 128-row tile each, no merge
  float* ws;               // tail_ctas * 2 slots of (D + 2) * 256 floats: un-normalised O, row max, row sum
  unsigned int* ws_count;  // [tail_items] arrival counters, zero before the launch
  // Ring attention, fused exchange: while this launch computes on one K/V shard, the two otherwise idle warps of
  // every CTA copy the NEXT step's shard out of its owner's memory (a peer mapping: the loads travel over NVLink /
  // NVSwitch) into a local buffer.  [0] = K, [1] = V; pf_words 16-byte words each, 0 = nothing to fetch.
  const uint4* pf_src[2];
  uint4* pf_dst[2];
  unsigned long long pf_words;
  DebugChannel dbg;
#ifdef FA_TRACE
  long long* trace;  // [6 roles][64 iterations][8 points]
  int trace_cta;
#endif
};

template <int D>
struct FaCfg {
  static_assert(D == 64 || D == 128, "head dim must be 64 or 128");
  static constexpr int kQT = 128;                       // query rows per tile
  static constexpr int kKT = 128;                       // keys per tile
  static constexpr int kBoxes = D / 64;                 // 64-column (128 B) TMA boxes per tile
  static constexpr int kBoxBytes = 128 * 128;           // 128 rows x 128 B
  static constexpr int kTileBytes = kBoxes * kBoxBytes; // one Q, K or V tile
  static constexpr int kStages = (D == 128) ? 5 : 10;   // K/V ring depth
  static constexpr int kThreads = 512;
  static constexpr int kTmemCols = 512;
  static constexpr uint32_t kColS0 = 0, kColS1 = 128, kColO0 = 256, kColO1 = 256 + D;
  // smem carve-up (bytes from a 1024-aligned base)
  static constexpr int kOffQ = 0;
  static constexpr int kOffKV = 2 * kTileBytes;
  static constexpr int kOffFac = kOffKV + kStages * kTileBytes;  // float[2][128] rescale factors / final m
  static constexpr int kOffSum = kOffFac + 2 * 128 * 4;          // float[2][128] final l
  static constexpr int kOffBar = kOffSum + 2 * 128 * 4;
  static constexpr int kPChunks = 2;                    // P is published in two chunks: keys [0, 16*kSplitGroups), then the rest
  static constexpr int kNumBars = 2 + 2 * kStages + 8 + 2 * kPChunks + 4;  // + merge_full[2] (stream-K tail), acc_full[2]
  static constexpr int kOffTmem = kOffBar + kNumBars * 8;
  static constexpr int kSmemBytes = kOffTmem + 16;
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget exceeded");
};

// mbarrier wait sites (reported through the debug channel on a timeout)
enum : uint32_t {
  SITE_TMA_EMPTY = 1, SITE_MMA_QFULL, SITE_MMA_KVFULL, SITE_MMA_PFULL,
  SITE_SM_SFULL, SITE_SM_CORR, SITE_CO_STATS, SITE_CO_LFULL, SITE_CO_OFULL, SITE_CO_MERGE, SITE_CO_ACC
};

constexpr float kLazyRescaleTau = 8.0f;  // log2 units: P stays below 2^8, far inside fp16 range
#ifndef FA_POLY_MASK
#define FA_POLY_MASK 0x11
#endif
// Which of the 8 key PAIRS of a 16-key group take the FMA-pipe polynomial exp2 instead of MUFU.EX2 (bit i = pair i;
// pairs inside the speculative head always use the MUFU).  Measured on B200, config 3, interleaved 2 s blocks at the
// power cap (scripts/sustained_ab.py, +-0.2 %): 0x00 -> 1020, 0x11 (25 %) -> 1050, 0x55 (50 %) -> 1025 TFLOP/s.  In
// short bursts the MUFU-only build is as fast or faster (a warp's MUFU stall and its FMA work do not overlap), but a
// launch then costs more energy, and the sustained rate is set by energy per launch.
constexpr unsigned kPolyMask = FA_POLY_MASK;
#ifndef FA_POLY_MASK_D64
#define FA_POLY_MASK_D64 0x49
#endif
// D = 64 halves the tensor work per key, so the MUFU weighs more: same harness at 8 x 32 x 4096 x 64 -- 0x00 -> 676,
// 0x11 -> 686, 0x49 (three pairs of eight, 37.5 %) -> 704, 0x55 -> 667 TFLOP/s.
constexpr unsigned kPolyMaskD64 = FA_POLY_MASK_D64;

// Schedule knobs (build-time; defaults are the measured best on B200, see DESIGN.md 3.1: period per 128 keys and tile
// from the clock64 trace, config 3 -- head 0/8/12/16/20/24/32/48/64 keys x split 2..7 groups: 16 x 4 -> 3234 cycles,
// 32 x 6 (the first shipped setting) -> 3356, 64 x 4 -> 4711)
#ifndef FA_SPEC_KEYS
#define FA_SPEC_KEYS 16   // exps computed against the stale max in the shadow of the row-max chain
#endif
#ifndef FA_SPLIT_GROUPS
#define FA_SPLIT_GROUPS 4  // P is published after this many 16-key groups, then after all 8
#endif
constexpr int kSpecKeys = FA_SPEC_KEYS;
constexpr int kSplitGroups = FA_SPLIT_GROUPS;
static_assert(kSpecKeys % 4 == 0 && kSpecKeys >= 0 && kSpecKeys <= 64, "speculative head is whole key quads");
static_assert(kSplitGroups >= 1 && kSplitGroups <= 7, "first publication covers 1..7 groups of 16 keys");

// Developer timeline trace (scripts/trace_fa.py builds a separate library with -DFA_TRACE): CTA
// `trace_cta` stores clock64() at the role hand-off points of every K/V iteration.  Compiled out of
// the product library.
#ifdef FA_TRACE
#define FA_TRACE_POINT(role, j, k)                                                                   \
  do {                                                                                                \
    if (p.trace != nullptr && blockIdx.x == static_cast<unsigned>(p.trace_cta) && (j) < 64)           \
      p.trace[((role) * 64 + (j)) * 8 + (k)] = clock64();                                             \
  } while (0)
#else
#define FA_TRACE_POINT(role, j, k) do { } while (0)
#endif

// One unit of work of a CTA: key tiles [kv0, kv1) of item `item`.  A whole item has slot = -1 and takes the ordinary
// epilogue.  A partial one parks its un-normalised (O, m, l) in workspace slot `slot` of its CTA (0: the span starts
// inside this item, 1: it started in an earlier item) and bumps the item's counter; the CTA that arrives last merges
// all parts in span order (deterministic whichever CTA that is) and writes the output.  This is the "stream-K" cut
// of the partial LAST WAVE; the host (api.cu, fa_plan) applies it only where it measured faster -- grids that fill
// less than half the SMs -- and otherwise cuts a small last wave along query rows (`half`), or not at all.
struct FaSegment {
  int item;      // global item index = bh * q_blocks + qblk
  int kv0, kv1;  // key-tile range
  int slot;      // -1 = whole item
  int tail_item; // index into ws_count (partial only)
  int half;      // -1 = both 128-row tiles of the item; 0 / 1 = only that tile (FaParams::tail_rows), all key tiles
};
// first query row and number of 128-row tiles of a segment
__device__ __forceinline__ void fa_segment_rows(const FaParams& p, const FaSegment& s, int& q0, int& ntile) {
  q0 = (s.item % p.q_blocks) * 256;
  ntile = (q0 + 128 < p.nq) ? 2 : 1;
  if (s.half >= 0) {
    q0 += s.half * 128;
    ntile = 1;
  }
}

// All of this is 32-bit arithmetic (a 64-bit division is a subroutine call that spills the caller's registers); the
// host only cuts the tail when (tail_items * nkv_tiles + 1) * tail_ctas < 2^31 (fa_plan).
__host__ __device__ __forceinline__ int fa_span_begin(const FaParams& p, int nkv_tiles, int t) {
  return static_cast<int>(static_cast<unsigned>(t) * static_cast<unsigned>(p.tail_items * nkv_tiles) /
                          static_cast<unsigned>(p.tail_ctas));
}
// tail CTA owning iteration x of the tail iteration space: largest t with span_begin(t) <= x
__host__ __device__ __forceinline__ int fa_span_owner(const FaParams& p, int nkv_tiles, int x) {
  const unsigned T = static_cast<unsigned>(p.tail_items * nkv_tiles);
  return static_cast<int>((static_cast<unsigned>(x + 1) * static_cast<unsigned>(p.tail_ctas) + T - 1u) / T) - 1;
}
__host__ __device__ __forceinline__ bool fa_segment(const FaParams& p, int nkv_tiles, int cta, int idx, FaSegment& s) {
  if (cta < p.items_full) {
    if (idx != 0) return false;
    s.item = cta; s.kv0 = 0; s.kv1 = nkv_tiles; s.slot = -1; s.tail_item = 0; s.half = -1;
    return true;
  }
  const int t = cta - p.items_full;
  s.half = -1;
  if (p.tail_rows) {  // CTA pair (2 li, 2 li + 1) shares item li: one query tile each, nothing to merge
    if (idx != 0) return false;
    s.item = p.items_full + (t >> 1); s.kv0 = 0; s.kv1 = nkv_tiles; s.slot = -1; s.tail_item = t >> 1; s.half = t & 1;
    return s.half == 0 || (s.item % p.q_blocks) * 256 + 128 < p.nq;  // a ragged last block may have no second tile
  }
  const int a = fa_span_begin(p, nkv_tiles, t), b = fa_span_begin(p, nkv_tiles, t + 1);
  const int li = static_cast<int>(static_cast<unsigned>(a) / static_cast<unsigned>(nkv_tiles)) + idx;
  const int i0 = li * nkv_tiles;
  if (i0 >= b) return false;
  const int lo = a > i0 ? a : i0, hi = b < i0 + nkv_tiles ? b : i0 + nkv_tiles;
  s.item = p.items_full + li;
  s.kv0 = lo - i0;
  s.kv1 = hi - i0;
  s.tail_item = li;
  s.slot = (s.kv0 == 0 && s.kv1 == nkv_tiles) ? -1 : (a >= i0 ? 0 : 1);
  return true;
}

// (Re-)arms every mbarrier of the CTA; one thread, between two CTA-wide barriers.  Layout as in the kernel:
// q_full[2], kv_full[S], kv_empty[S], s_full[2], p_full[2][chunks], stats_full[2], o_full[2], l_full[2], merge_full[2], acc_full[2].
template <int D>
__device__ __forceinline__ void fa_init_barriers(uint64_t* bars, bool reinit) {
  using Cfg = FaCfg<D>;
  if (reinit)
    for (int b = 0; b < Cfg::kNumBars; ++b) mbar_inval(&bars[b]);
  uint64_t* b = bars;
  for (int i = 0; i < 2 + 2 * Cfg::kStages + 2; ++i) mbar_init(b++, 1);  // q_full, kv_full, kv_empty, s_full
  // softmax / correction warps arrive once per warp (lane 0 after __syncwarp): count 4.  P chunk 0 also collects the
  // correction warps: "P ready" and "O rescaled" are one wait for the MMA warp
  for (int i = 0; i < 2; ++i)
    for (int c = 0; c < Cfg::kPChunks; ++c) mbar_init(b++, c == 0 ? 8 : 4);
  for (int i = 0; i < 2; ++i) mbar_init(b++, 4);  // stats_full
  for (int i = 0; i < 2; ++i) mbar_init(b++, 1);  // o_full
  for (int i = 0; i < 2; ++i) mbar_init(b++, 4);  // l_full
  for (int i = 0; i < 2; ++i) mbar_init(b++, 1);  // merge_full
  for (int i = 0; i < 2; ++i) mbar_init(b++, 1);  // acc_full
  fence_mbar_init();
}
// Between two segments of a tail CTA: every role has finished the previous segment (the epilogue's TMA store has read
// its staging tile), and the barriers start again at phase 0.  Whole-item CTAs never come here.
template <int D>
__device__ __forceinline__ void fa_segment_boundary(uint64_t* bars, int warp, int lane) {
  tc_fence_before();
  named_bar_sync(2, 14 * 32);  // the 14 working warps; warps 14 and 15 only hold registers for the others
  if (warp == 12 && lane == 0) fa_init_barriers<D>(bars, true);
  named_bar_sync(2, 14 * 32);
  tc_fence_after();
}

template <int D>
__global__ void __launch_bounds__(512, 1)
fa_fwd_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
              const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapO,
              const __grid_constant__ CUtensorMap mapO32, const FaParams p) {
  using Cfg = FaCfg<D>;
  extern __shared__ __align__(1024) uint8_t smem[];

  uint8_t* const q_smem = smem + Cfg::kOffQ;
  uint8_t* const kv_smem = smem + Cfg::kOffKV;
  float* const fac_smem = reinterpret_cast<float*>(smem + Cfg::kOffFac);
  float* const sum_smem = reinterpret_cast<float*>(smem + Cfg::kOffSum);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  uint64_t* const q_full = bars;                          // [2]
  uint64_t* const kv_full = bars + 2;                     // [kStages]
  uint64_t* const kv_empty = kv_full + Cfg::kStages;      // [kStages]
  uint64_t* const s_full = kv_empty + Cfg::kStages;       // [2]  MMA -> softmax : S tile ready
  uint64_t* const p_full = s_full + 2;                    // [2][kPChunks]  softmax -> MMA : P chunk in TMEM
  uint64_t* const stats_full = p_full + 2 * Cfg::kPChunks;  // [2]  softmax -> correction : factors
  uint64_t* const o_full = stats_full + 2;                // [2]  MMA -> correction : last PV done
  uint64_t* const l_full = o_full + 2;                    // [2]  softmax -> correction : final l
  uint64_t* const merge_full = l_full + 2;                // [2]  bulk copy -> correction : a parked part landed (tail merge)
  uint64_t* const acc_full = merge_full + 2;              // [2]  TMA -> correction : carried fp32 O_acc tile landed (STATE mode)
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::kOffTmem);  // [0] TMEM base, [1] merge verdict

  const int warp = threadIdx.x >> 5;
  const DebugChannel dbg = p.dbg;
  // per-role copies of the cheap common values: kept across the setmaxnreg boundaries they would be spilled
#define FA_ROLE_COMMON()                                            \
  const int lane = threadIdx.x & 31;                                \
  const int cta = blockIdx.x;                                       \
  const uint32_t tmem_base = *tmem_slot;                            \
  const int nkv_tiles = (p.nkv + Cfg::kKT - 1) / Cfg::kKT

  // ------------------------------- one-time setup -------------------------------------
  if (threadIdx.x == 12 * 32) {
    if ((smem_u32(smem) & 1023u) != 0u) wait_timeout_trap(dbg, 0xA11u, 0, smem_u32(smem));
    fa_init_barriers<D>(bars, false);
  }
  if (threadIdx.x == 13 * 32) {
    prefetch_tensormap(&mapQ);
    prefetch_tensormap(&mapK);
    prefetch_tensormap(&mapV);
    if (p.mode != FA_MODE_F32 && !(p.mode == FA_MODE_STATE && !p.is_last)) prefetch_tensormap(&mapO);
    if (p.mode != FA_MODE_F16) prefetch_tensormap(&mapO32);
  }
  if (warp == 8) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp < 8) {
    // =============================== softmax warpgroups ===============================
    setmaxnreg_inc<192>();
    FA_ROLE_COMMON();
    const int i = warp >> 2;  // Q tile handled by this warpgroup
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t t_s = tmem_base + lane_base + (i == 0 ? Cfg::kColS0 : Cfg::kColS1);
    float c = p.scale_log2;
    const bool neg_scale = c < 0.f;  // max over scale*s == -|scale| * min(s): fold the sign into s
    if (neg_scale) c = -c;
    FaSegment seg;
    for (int si = 0; fa_segment(p, nkv_tiles, cta, si, seg); ++si) {
      if (si > 0) fa_segment_boundary<D>(bars, warp, lane);
      int q0, ntile;
      fa_segment_rows(p, seg, q0, ntile);
      if (i >= ntile) continue;
      const int n_it = seg.kv1 - seg.kv0;
      float m_used = -INFINITY;  // running max of (sign-folded) raw scores the exponent refers to
      float l = 0.f;

      for (int j = 0; j < n_it; ++j) {
        if ((warp & 3) == 0 && lane == 0) FA_TRACE_POINT(i, j, 0);
        mbar_wait(&s_full[i], j & 1, dbg, SITE_SM_SFULL, j);
        tc_fence_after();
        if ((warp & 3) == 0 && lane == 0) FA_TRACE_POINT(i, j, 1);
        uint32_t sr[128];
        tmem_ld_x32(t_s + 0, sr + 0);
        tmem_ld_x32(t_s + 32, sr + 32);
        tmem_ld_x32(t_s + 64, sr + 64);
        tmem_ld_x32(t_s + 96, sr + 96);
        tmem_wait_ld();
        if ((warp & 3) == 0 && lane == 0) FA_TRACE_POINT(i, j, 2);
        float* s = reinterpret_cast<float*>(sr);
        if (neg_scale) {
#pragma unroll
          for (int k = 0; k < 128; ++k) s[k] = -s[k];
        }
        if (seg.kv0 + j == nkv_tiles - 1 && (p.nkv & 127) != 0) {
          const int valid = p.nkv - (nkv_tiles - 1) * 128;  // keys beyond nkv were zero-filled by TMA: mask them
#pragma unroll
          for (int k = 0; k < 128; ++k)
            if (k >= valid) s[k] = -INFINITY;
        }
        // ---- speculative head: the exps of the first kSpecKeys keys are computed against the STALE running
        // max while the 64-instruction row-max chain runs in their shadow (MUFU and ALU pipes overlap).
        // The lazy-rescale rule makes the stale max the common case; when a row does need a new max
        // (always at j == 0, where the stale max is -inf) those exps are simply redone.
        const uint64_t c2 = pack_f32x2(c, c);
        float neg_mc = -m_used * c;
        uint64_t neg_mc2 = pack_f32x2(neg_mc, neg_mc);
        float pe[kSpecKeys > 0 ? kSpecKeys : 4];
#pragma unroll
        for (int k = 0; k < kSpecKeys; k += 2) {
          float x0, x1;
          unpack_f32x2(ffma2(pack_f32x2(s[k], s[k + 1]), c2, neg_mc2), x0, x1);
          pe[k] = ex2_approx(x0);
          pe[k + 1] = ex2_approx(x1);
        }
        // four independent 3-input max chains, one per 32-column quarter
        float mx0 = s[0], mx1 = s[32], mx2 = s[64], mx3 = s[96];
#ifdef FA_MAX2  // 2-input max chains (experiment)
#pragma unroll
        for (int k = 1; k < 32; ++k) {
          mx0 = fmaxf(mx0, s[k]); mx1 = fmaxf(mx1, s[32 + k]); mx2 = fmaxf(mx2, s[64 + k]); mx3 = fmaxf(mx3, s[96 + k]);
        }
        const float m_new = fmaxf(fmaxf(mx0, mx1), fmaxf(fmaxf(mx2, mx3), m_used));
#else
#pragma unroll
        for (int k = 1; k < 31; k += 2) {
          mx0 = fmax3(mx0, s[k], s[k + 1]);
          mx1 = fmax3(mx1, s[32 + k], s[32 + k + 1]);
          mx2 = fmax3(mx2, s[64 + k], s[64 + k + 1]);
          mx3 = fmax3(mx3, s[96 + k], s[96 + k + 1]);
        }
        mx0 = fmax3(mx0, s[31], s[63]);
        mx2 = fmax3(mx2, s[95], s[127]);
        const float m_new = fmaxf(fmax3(mx0, mx1, mx2), fmaxf(mx3, m_used));
#endif
        if ((warp & 3) == 0 && lane == 0) FA_TRACE_POINT(i, j, 6);

        // lazy rescale: keep the stale max while the true max exceeds it by at most 2^tau
        const bool need = (m_new - m_used) * c > kLazyRescaleTau;  // j == 0: (x - -inf) = +inf -> true
        float fac = 1.f;
        if (__any_sync(0xffffffffu, need)) {
          if (need) {
            fac = ex2_approx((m_used - m_new) * c);  // j == 0: exp2(-inf) = 0, and l is 0
            l *= fac;
            m_used = m_new;
            neg_mc = -m_used * c;
            neg_mc2 = pack_f32x2(neg_mc, neg_mc);
#pragma unroll
            for (int k = 0; k < kSpecKeys; k += 2) {
              float x0, x1;
              unpack_f32x2(ffma2(pack_f32x2(s[k], s[k + 1]), c2, neg_mc2), x0, x1);
              pe[k] = ex2_approx(x0);
              pe[k + 1] = ex2_approx(x1);
            }
          }
        }
        if (j > 0) {
          fac_smem[i * 128 + row] = fac;
          __syncwarp();
          if (lane == 0) mbar_arrive(&stats_full[i]);
        }
        if ((warp & 3) == 0 && lane == 0) FA_TRACE_POINT(i, j, 3);

        // ---- exps, packed row sums, fp16 P written back over S; P is published to the MMA warp in two
        // chunks (keys [0,64) then [64,128)) so half of the PV MMAs overlap the tail of the softmax
        uint64_t lsum_a = pack_f32x2(0.f, 0.f), lsum_b = pack_f32x2(0.f, 0.f);
        constexpr unsigned poly_mask = (D == 64) ? kPolyMaskD64 : kPolyMask;
        const float cy = c * (1.f / 128.f);                 // y = sat(s*cy + by) encodes x = 128*y - 120
        const float by = (neg_mc + 120.f) * (1.f / 128.f);
#pragma unroll
        for (int g = 0; g < 8; ++g) {  // 16 keys per group -> 8 packed fp16x2 columns
          uint32_t pk[8];
#pragma unroll
          for (int k = 0; k < 16; k += 4) {
            const int e = g * 16 + k;
            float p0, p1, p2, p3;
            if (e < kSpecKeys) {
              p0 = pe[e]; p1 = pe[e + 1]; p2 = pe[e + 2]; p3 = pe[e + 3];
            } else {
              float x0, x1, x2, x3;
              if ((poly_mask >> (k / 2)) & 1u) {
                exp2_poly_pair(pack_f32x2(ffma_sat(s[e + 0], cy, by), ffma_sat(s[e + 1], cy, by)), p0, p1);
              } else {
                unpack_f32x2(ffma2(pack_f32x2(s[e + 0], s[e + 1]), c2, neg_mc2), x0, x1);
                p0 = ex2_approx(x0); p1 = ex2_approx(x1);
              }
              if ((poly_mask >> (k / 2 + 1)) & 1u) {
                exp2_poly_pair(pack_f32x2(ffma_sat(s[e + 2], cy, by), ffma_sat(s[e + 3], cy, by)), p2, p3);
              } else {
                unpack_f32x2(ffma2(pack_f32x2(s[e + 2], s[e + 3]), c2, neg_mc2), x2, x3);
                p2 = ex2_approx(x2); p3 = ex2_approx(x3);
              }
            }
            lsum_a = fadd2(lsum_a, pack_f32x2(p0, p1));
            lsum_b = fadd2(lsum_b, pack_f32x2(p2, p3));
            pk[k / 2 + 0] = pack_half2(p0, p1);
            pk[k / 2 + 1] = pack_half2(p2, p3);
          }
          tmem_st_x8(t_s + g * 8, pk);  // P (fp16 pairs) overwrites columns [0,64) of its own S
          if (g == kSplitGroups - 1 || g == 7) {
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[i * Cfg::kPChunks + (g == 7 ? 1 : 0)]);
            if (g != 7 && (warp & 3) == 0 && lane == 0) FA_TRACE_POINT(i, j, 4);
          }
        }
        if ((warp & 3) == 0 && lane == 0) FA_TRACE_POINT(i, j, 5);
        {
          float a0, a1;
          unpack_f32x2(fadd2(lsum_a, lsum_b), a0, a1);
          l += a0 + a1;
        }
      }

      // final statistics for the epilogue
      sum_smem[i * 128 + row] = l;
      if (p.mode == FA_MODE_STATE || seg.slot >= 0) {
        // fac_smem doubles as the final-max slot; the correction WG has consumed its last factor
        // once the last phase of p_full[chunk 0] (which its warps also arrive on) has completed.
        mbar_wait(&p_full[i * Cfg::kPChunks], (n_it - 1) & 1, dbg, SITE_SM_CORR, i);
        fac_smem[i * 128 + row] = m_used;  // max of the sign-folded raw scores
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&l_full[i]);
    }
  } else if (warp < 12) {
    // ========================= correction + epilogue warpgroup ========================
    setmaxnreg_dec<80>();
    FA_ROLE_COMMON();
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(wq * 32) << 16;
    const float c_abs = fabsf(p.scale_log2);
    constexpr int kSlotFloats = (D + 2) * 256;  // workspace slot: O as [D/4][256 rows] float4, then m[256], l[256]
    FaSegment seg;
    for (int si = 0; fa_segment(p, nkv_tiles, cta, si, seg); ++si) {
      if (si > 0) {
        if (warp == 8 && lane == 0) tma_store_wait_read<0>();  // the staging tiles become Q tiles again
        fa_segment_boundary<D>(bars, warp, lane);
      }
      const int bh = seg.item / p.q_blocks;
      int q0, ntile;
      fa_segment_rows(p, seg, q0, ntile);
      const int n_it = seg.kv1 - seg.kv0;

      // j == 0: nothing to rescale (PV(0) overwrites O); the arrival only balances the barrier count
      if (lane == 0)
        for (int i = 0; i < ntile; ++i) mbar_arrive(&p_full[i * Cfg::kPChunks]);
      for (int j = 1; j < n_it; ++j) {
        for (int i = 0; i < ntile; ++i) {
          mbar_wait(&stats_full[i], (j - 1) & 1, dbg, SITE_CO_STATS, j);
          if (warp == 8 && lane == 0) FA_TRACE_POINT(2 + i, j, 0);
          const float fac = fac_smem[i * 128 + row];
          if (__any_sync(0xffffffffu, fac != 1.f)) {
            tc_fence_after();
            const uint32_t t_o = tmem_base + lane_base + (i == 0 ? Cfg::kColO0 : Cfg::kColO1);
#pragma unroll
            for (int cc = 0; cc < D / 32; ++cc) {
              uint32_t o[32];
              tmem_ld_x32(t_o + cc * 32, o);
              tmem_wait_ld();
#pragma unroll
              for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * fac);
              tmem_st_x32(t_o + cc * 32, o);
            }
            tmem_wait_st();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[i * Cfg::kPChunks]);
          if (warp == 8 && lane == 0) FA_TRACE_POINT(2 + i, j, 1);
        }
      }

      // ---- a partial segment parks its state; the part that arrives last merges them all ----
      // Slot layout (floats): tile 0 block, tile 1 block, m[256], l[256]; a tile block is D/4 x 128 float4, index
      // (column quad) * 128 + row: consecutive lanes are consecutive float4 on the way out (coalesced) and, after a
      // 1-D bulk copy of the whole block into shared memory, on the way back in (conflict-free).
      if (seg.slot >= 0) {
        constexpr int kTileFloats = D * 128;
        float* const slot = p.ws + (static_cast<size_t>(cta - p.items_full) * 2 + seg.slot) * kSlotFloats;
        for (int i = 0; i < ntile; ++i) {
          mbar_wait(&l_full[i], 0, dbg, SITE_CO_LFULL, i);
          mbar_wait(&o_full[i], 0, dbg, SITE_CO_OFULL, i);
          tc_fence_after();
          const uint32_t t_o = tmem_base + lane_base + (i == 0 ? Cfg::kColO0 : Cfg::kColO1);
          const int ri = i * 128 + row;
          __stcg(slot + 2 * kTileFloats + ri, fac_smem[ri]);
          __stcg(slot + 2 * kTileFloats + 256 + ri, sum_smem[ri]);
          float4* const block = reinterpret_cast<float4*>(slot + i * kTileFloats);
#pragma unroll 1
          for (int cc = 0; cc < D / 32; ++cc) {
            uint32_t o[32];
            tmem_ld_x32(t_o + cc * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4)
              __stcg(block + (cc * 8 + k4) * 128 + row,
                     make_float4(__uint_as_float(o[4 * k4]), __uint_as_float(o[4 * k4 + 1]),
                                 __uint_as_float(o[4 * k4 + 2]), __uint_as_float(o[4 * k4 + 3])));
          }
        }
        // parts of this item: tail CTAs part_first..part_last; every part but the first starts its span inside
        // the item (slot 0), the first one does so only if its span begins exactly at the item
        const int item_begin = seg.tail_item * nkv_tiles;
        const int part_first = fa_span_owner(p, nkv_tiles, item_begin);
        const int nparts = fa_span_owner(p, nkv_tiles, item_begin + nkv_tiles - 1) - part_first + 1;
        const int first_slot = fa_span_begin(p, nkv_tiles, part_first) == item_begin ? 0 : 1;
        __threadfence();
        named_bar_sync(1, 128);
        if (warp == 8 && lane == 0) {
          const unsigned int arrived = atomicAdd(p.ws_count + seg.tail_item, 1u);
          tmem_slot[1] = (arrived == static_cast<unsigned int>(nparts - 1)) ? 1u : 0u;
          __threadfence();
        }
        named_bar_sync(1, 128);
        if (tmem_slot[1] == 0u) continue;  // another CTA finishes this item
        // Last to arrive: reduce ALL parts (its own included: the same arithmetic whichever CTA ends up here), in span
        // order, into this CTA's O accumulator in TMEM and its statistics, then fall through to the ordinary epilogue.
        // The parts come back through 1-D bulk copies into the idle K/V ring, double-buffered: part j + 1 is in
        // flight while part j is accumulated.
        auto part_slot = [&](int k) {
          return p.ws + (static_cast<size_t>(part_first + k) * 2 + (k == 0 ? first_slot : 0)) * kSlotFloats;
        };
        const int steps = ntile * nparts;
        auto fetch = [&](int j) {  // one thread
          const int b = j & 1;
          mbar_arrive_expect_tx(&merge_full[b], kTileFloats * 4);
          bulk_copy_g2s(kv_smem + b * (kTileFloats * 4), part_slot(j % nparts) + (j / nparts) * kTileFloats, kTileFloats * 4,
                        &merge_full[b]);
        };
        if (warp == 8) {  // whole warp, elected lane: a bulk copy under `lane == 0` gets a per-thread serialisation loop
          if (elect_one()) {
            fence_proxy_async_all();
            fetch(0);
          }
          __syncwarp();
        }
        float m_all = 0.f, l_all = 0.f;
        for (int j = 0; j < steps; ++j) {
          const int i = j / nparts, k = j - i * nparts;
          const int ri = i * 128 + row;
          const uint32_t t_o = tmem_base + lane_base + (i == 0 ? Cfg::kColO0 : Cfg::kColO1);
          if (warp == 8 && j + 1 < steps) {  // the buffer of step j + 1 was released at the end of step j - 1
            if (elect_one()) fetch(j + 1);
            __syncwarp();
          }
          if (k == 0) {
            m_all = -INFINITY;
            for (int kk = 0; kk < nparts; ++kk) m_all = fmaxf(m_all, __ldcg(part_slot(kk) + 2 * kTileFloats + ri));
            l_all = 0.f;
            for (int kk = 0; kk < nparts; ++kk) {
              const float* const ps = part_slot(kk) + 2 * kTileFloats;
              l_all = fmaf(__ldcg(ps + 256 + ri), ex2_approx((__ldcg(ps + ri) - m_all) * c_abs), l_all);
            }
          }
          const float w = ex2_approx((__ldcg(part_slot(k) + 2 * kTileFloats + ri) - m_all) * c_abs);
          mbar_wait(&merge_full[j & 1], (j >> 1) & 1, dbg, SITE_CO_MERGE, j);
          const float4* const src = reinterpret_cast<const float4*>(kv_smem + (j & 1) * (kTileFloats * 4));
#pragma unroll 1
          for (int cc = 0; cc < D / 32; ++cc) {
            uint32_t o[32];
            if (k > 0) {
              tmem_ld_x32(t_o + cc * 32, o);
              tmem_wait_ld();
            } else {
#pragma unroll
              for (int e = 0; e < 32; ++e) o[e] = 0u;
            }
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
              const float4 a = src[(cc * 8 + k4) * 128 + row];
              o[4 * k4 + 0] = __float_as_uint(fmaf(a.x, w, __uint_as_float(o[4 * k4 + 0])));
              o[4 * k4 + 1] = __float_as_uint(fmaf(a.y, w, __uint_as_float(o[4 * k4 + 1])));
              o[4 * k4 + 2] = __float_as_uint(fmaf(a.z, w, __uint_as_float(o[4 * k4 + 2])));
              o[4 * k4 + 3] = __float_as_uint(fmaf(a.w, w, __uint_as_float(o[4 * k4 + 3])));
            }
            tmem_st_x32(t_o + cc * 32, o);
          }
          tmem_wait_st();
          if (k == nparts - 1) {
            sum_smem[ri] = l_all;  // read back below by this same thread
            fac_smem[ri] = m_all;
          }
          named_bar_sync(1, 128);  // every thread is done with buffer j & 1
        }
      }

      // ---- epilogue: O / l -> output ----
      // fp16 output leaves through the (dead) Q tiles and a TMA store.  fp32 traffic -- the fp32 output of
      // fa_forward_f32out / _cross and the carried O_acc of a ring step, read AND written -- goes through TMA as well,
      // staged in the K/V ring, which is idle once the last PV of the last tile has completed: a lane owns a whole
      // accumulator row (tcgen05.ld 32x32b), the worst shape for global memory (512-byte stride between lanes), and
      // row-per-thread float4 loads / stores cost a ring step of BASELINE config 4 at 8 GPUs 15 % of its time.
      const bool f32_traffic = (p.mode != FA_MODE_F16);
      const bool have_acc = (p.mode == FA_MODE_STATE) && !p.is_first;
      constexpr int kTileF32Bytes = D * 128 * 4, kBoxF32Bytes = 128 * 128;  // boxes of 32 fp32 columns x 128 rows
      if (f32_traffic) {
        mbar_wait(&o_full[ntile - 1], 0, dbg, SITE_CO_OFULL, ntile - 1);  // nothing reads the ring any more
        if (have_acc && warp == 8) {
          const int bh_u = __shfl_sync(0xffffffffu, bh, 0), q0_u = __shfl_sync(0xffffffffu, q0, 0);
          const int nt_u = __shfl_sync(0xffffffffu, ntile, 0);
          if (elect_one()) {
            for (int i = 0; i < nt_u; ++i) {
              mbar_arrive_expect_tx(&acc_full[i], kTileF32Bytes);
#pragma unroll
              for (int bx = 0; bx < D / 32; ++bx)
                tma_load_3d(kv_smem + i * kTileF32Bytes + bx * kBoxF32Bytes, &mapO32, &acc_full[i], bx * 32, q0_u + i * 128,
                            bh_u, kEvictFirst);  // rows >= nq arrive as zeros
            }
          }
          __syncwarp();
        }
      }
      for (int i = 0; i < ntile; ++i) {
        mbar_wait(&l_full[i], 0, dbg, SITE_CO_LFULL, i);
        mbar_wait(&o_full[i], 0, dbg, SITE_CO_OFULL, i);
        tc_fence_after();
        const float l_new = sum_smem[i * 128 + row];
        const float m_fold = fac_smem[i * 128 + row];  // meaningful in STATE mode and after a merge
        const uint32_t t_o = tmem_base + lane_base + (i == 0 ? Cfg::kColO0 : Cfg::kColO1);
        const int grow = q0 + i * 128 + row;  // query row inside this (b,h)
        const bool row_ok = grow < p.nq;
        float o_scale = 1.f;      // factor applied to this step's un-normalised O
        float acc_scale = 0.f;    // factor applied to the carried O_acc (STATE mode)
        bool emit_f16 = (p.mode == FA_MODE_F16);
        const size_t row_index = static_cast<size_t>(bh) * p.nq + grow;
        if (p.mode == FA_MODE_STATE) {
          // Merge (O, m, l) of this K/V shard into the carried state: the reference's tile-to-tile
          // rule, Streaming_FlashAttention_Forward_Kernel/flashattn_streaming_16x16.cu:247-261.
          const float m_new_s = m_fold * (c_abs * 0.6931471805599453f);  // max of scale*s for this shard
          float m_acc = -INFINITY, l_acc = 0.f;
          if (!p.is_first && row_ok) {
            m_acc = p.m_state[row_index];
            l_acc = p.l_state[row_index];
          }
          const float m_tot = fmaxf(m_acc, m_new_s);
          acc_scale = (m_acc == -INFINITY) ? 0.f : ex2_approx((m_acc - m_tot) * 1.4426950408889634f);
          o_scale = ex2_approx((m_new_s - m_tot) * 1.4426950408889634f);
          const float l_tot = l_acc * acc_scale + l_new * o_scale;
          if (p.is_last) {
            emit_f16 = true;
            const float inv = 1.f / l_tot;
            o_scale *= inv;
            acc_scale *= inv;
          } else if (row_ok) {
            p.m_state[row_index] = m_tot;
            p.l_state[row_index] = l_tot;
          }
        } else {
          o_scale = 1.f / l_new;
        }
        uint8_t* const stage = q_smem + i * Cfg::kTileBytes;           // fp16: Q tile i is dead once o_full fired
        uint8_t* const stage32 = kv_smem + i * kTileF32Bytes;          // fp32: the idle K/V ring
        if (have_acc) mbar_wait(&acc_full[i], 0, dbg, SITE_CO_ACC, i);

#pragma unroll
        for (int cc = 0; cc < D / 32; ++cc) {
          uint32_t o[32];
          tmem_ld_x32(t_o + cc * 32, o);
          tmem_wait_ld();
          float v[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(o[k]) * o_scale;
          // this row's 32 columns inside a 128B-swizzled box of 32 fp32 columns: eight 16-byte chunks, chunk q at q ^ (row & 7)
          uint8_t* const row32 = stage32 + cc * kBoxF32Bytes + row * 128;
          if (have_acc) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 a = *reinterpret_cast<const float4*>(row32 + ((q ^ (row & 7)) << 4));
              v[4 * q + 0] = fmaf(a.x, acc_scale, v[4 * q + 0]);
              v[4 * q + 1] = fmaf(a.y, acc_scale, v[4 * q + 1]);
              v[4 * q + 2] = fmaf(a.z, acc_scale, v[4 * q + 2]);
              v[4 * q + 3] = fmaf(a.w, acc_scale, v[4 * q + 3]);
            }
          }
          if (emit_f16) {
            // 32 columns = 64 B = four 16-byte chunks of this row inside a 128B-swizzled box of 64 fp16 columns
            uint8_t* const box = stage + (cc >> 1) * Cfg::kBoxBytes + row * 128;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 w;
              w.x = pack_half2(v[q * 8 + 0], v[q * 8 + 1]);
              w.y = pack_half2(v[q * 8 + 2], v[q * 8 + 3]);
              w.z = pack_half2(v[q * 8 + 4], v[q * 8 + 5]);
              w.w = pack_half2(v[q * 8 + 6], v[q * 8 + 7]);
              const int chunk = (cc & 1) * 4 + q;
              *reinterpret_cast<uint4*>(box + ((chunk ^ (row & 7)) << 4)) = w;
            }
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4*>(row32 + ((q ^ (row & 7)) << 4)) =
                  make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (warp == 8) {
          // warp-uniform coordinates (a shuffle from lane 0 is what lets ptxas issue the TMA store without a
          // per-thread serialisation loop); lane 0 owns the bulk group it later waits on
          const int row0 = __shfl_sync(0xffffffffu, q0 + i * 128, 0), bh_u = __shfl_sync(0xffffffffu, bh, 0);
          const int i_u = __shfl_sync(0xffffffffu, i, 0);
          if (lane == 0) {
            if (emit_f16) {
#pragma unroll
              for (int b = 0; b < Cfg::kBoxes; ++b)
                tma_store_3d(&mapO, q_smem + i_u * Cfg::kTileBytes + b * Cfg::kBoxBytes, b * 64, row0, bh_u);  // clips rows >= nq
            } else {
#pragma unroll
              for (int b = 0; b < D / 32; ++b)
                tma_store_3d(&mapO32, kv_smem + i_u * kTileF32Bytes + b * kBoxF32Bytes, b * 32, row0, bh_u);
            }
            tma_store_commit();
          }
        }
      }
    }
    if (warp == 8 && lane == 0) tma_store_wait_all<0>();
  } else if (warp == 12) {
    // ================================ MMA issuer ======================================
    // The whole warp walks the schedule (uniform control flow: every lane waits on the barriers);
    // one lane chosen by elect.sync issues the tcgen05 instructions.  (Putting the loop under
    // `if (lane == 0)` makes ptxas wrap every UTCHMMA in a per-thread serialisation loop.)
    // Static order: QK0(0) QK1(0) | PV0(j) QK0(j+1) PV1(j) QK1(j+1) | ...  -- tcgen05 ops of one thread
    // execute in order, which gives PV_i(j) before QK_i(j+1) (whose S overwrites P_i) for free.
    // (Two issuing warps, one per tile, were measured slower: the tiles' MMAs interleave on the pipe
    //  and lengthen each tile's S -> softmax -> P chain.)
    setmaxnreg_dec<48>();
    {
      FA_ROLE_COMMON();
      constexpr uint32_t idesc_qk = umma_idesc_f16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_f16(128, D, 0, 1);
      const uint32_t q_addr = smem_u32(q_smem);
      const uint32_t kv_addr = smem_u32(kv_smem);

      // S_i = Q_i K^T; signals s_full[i]; optionally releases the K slot afterwards
      auto issue_qk = [&](int i, int slot, uint64_t* release) {
        const uint32_t a0 = q_addr + i * Cfg::kTileBytes;
        const uint32_t b0 = kv_addr + slot * Cfg::kTileBytes;
        const uint32_t d = tmem_base + (i == 0 ? Cfg::kColS0 : Cfg::kColS1);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < D / 16; ++ks) {
            const uint32_t off = (ks >> 2) * Cfg::kBoxBytes + (ks & 3) * 32;
            umma_ss(d, umma_desc_sw128(a0 + off, 16, 1024), umma_desc_sw128(b0 + off, 16, 1024), idesc_qk,
                    ks > 0 ? 1u : 0u);
          }
          umma_commit(&s_full[i]);
          if (release != nullptr) umma_commit(release);
        }
        __syncwarp();
      };
      // O_i (+)= P_i[:, chunk c] V[chunk c, :]: chunk 0 = the first kSplitGroups 16-key MMAs, chunk 1 = the rest
      auto issue_pv_chunk = [&](int i, int slot, int c, bool accumulate, uint64_t* done0, uint64_t* done1) {
        const uint32_t b0 = kv_addr + slot * Cfg::kTileBytes;
        const uint32_t d = tmem_base + (i == 0 ? Cfg::kColO0 : Cfg::kColO1);
        const uint32_t a = tmem_base + (i == 0 ? Cfg::kColS0 : Cfg::kColS1);
        if (elect_one()) {
          const int ks_begin = (c == 0) ? 0 : kSplitGroups, ks_end = (c == 0) ? kSplitGroups : 8;
#pragma unroll
          for (int ks = ks_begin; ks < ks_end; ++ks) {
            // V tile: rows = keys (128 B of head dim each): MN-major B operand, 16 keys = 2048 B per step
            umma_ts(d, a + ks * 8, umma_desc_sw128(b0 + ks * 2048, Cfg::kBoxBytes, 1024), idesc_pv,
                    (accumulate || ks > 0) ? 1u : 0u);
          }
          if (done0 != nullptr) umma_commit(done0);
          if (done1 != nullptr) umma_commit(done1);
        }
        __syncwarp();
      };
      auto slot_of = [](int t) { return t % Cfg::kStages; };
      auto par_of = [](int t) { return (t / Cfg::kStages) & 1; };

      FaSegment seg;
      for (int si = 0; fa_segment(p, nkv_tiles, cta, si, seg); ++si) {
        if (si > 0) fa_segment_boundary<D>(bars, warp, lane);
        int q0, ntile;
        fa_segment_rows(p, seg, q0, ntile);
        const int n_it = seg.kv1 - seg.kv0;

        // prologue: S_i(0) = Q_i K(0)^T
        mbar_wait(&q_full[0], 0, dbg, SITE_MMA_QFULL, 0);
        mbar_wait(&kv_full[slot_of(0)], par_of(0), dbg, SITE_MMA_KVFULL, 0);
        tc_fence_after();
        issue_qk(0, slot_of(0), ntile == 2 ? nullptr : &kv_empty[slot_of(0)]);
        if (ntile == 2) {
          mbar_wait(&q_full[1], 0, dbg, SITE_MMA_QFULL, 1);
          tc_fence_after();
          issue_qk(1, slot_of(0), &kv_empty[slot_of(0)]);
        }

        for (int j = 0; j < n_it; ++j) {
          const int tv = 2 * j + 1, tk = 2 * j + 2;
          const bool more = (j + 1 < n_it);
          if (lane == 0) FA_TRACE_POINT(4, j, 5);
          mbar_wait(&kv_full[slot_of(tv)], par_of(tv), dbg, SITE_MMA_KVFULL, tv);
          if (more) mbar_wait(&kv_full[slot_of(tk)], par_of(tk), dbg, SITE_MMA_KVFULL, tk);
          if (lane == 0) FA_TRACE_POINT(4, j, 6);
          for (int i = 0; i < ntile; ++i) {
            const bool last_tile = (i == ntile - 1);
            uint64_t* const p_full_i = &p_full[i * Cfg::kPChunks];
            if (lane == 0) FA_TRACE_POINT(4 + i, j, 0);
            mbar_wait(&p_full_i[0], j & 1, dbg, SITE_MMA_PFULL, j);  // P chunk 0 written and O rescaled
            tc_fence_after();
            if (lane == 0) FA_TRACE_POINT(4 + i, j, 1);
            issue_pv_chunk(i, slot_of(tv), 0, j > 0, nullptr, nullptr);
            mbar_wait(&p_full_i[1], j & 1, dbg, SITE_MMA_PFULL, j);
            tc_fence_after();
            // V(j) is released by the commit that follows its last reader; the final PV also signals o_full
            issue_pv_chunk(i, slot_of(tv), 1, true, last_tile ? &kv_empty[slot_of(tv)] : nullptr,
                           more ? nullptr : &o_full[i]);
            if (lane == 0) FA_TRACE_POINT(4 + i, j, 3);
            if (more) {
              issue_qk(i, slot_of(tk), last_tile ? &kv_empty[slot_of(tk)] : nullptr);
              if (lane == 0) FA_TRACE_POINT(4 + i, j, 4);
            }
          }
        }
      }
    }
  } else if (warp == 13) {
    // ================================ TMA producer ====================================
    setmaxnreg_dec<48>();
    {
      FA_ROLE_COMMON();
      (void)tmem_base;
      FaSegment seg;
      for (int si = 0; fa_segment(p, nkv_tiles, cta, si, seg); ++si) {
        if (si > 0) fa_segment_boundary<D>(bars, warp, lane);
        const int bh = seg.item / p.q_blocks;
        int q0, ntile;
        fa_segment_rows(p, seg, q0, ntile);
        auto load_tile = [&](const CUtensorMap* map, uint8_t* dst, uint64_t* bar, int row0, uint64_t hint) {
          if (elect_one()) {
            mbar_arrive_expect_tx(bar, Cfg::kTileBytes);
#pragma unroll
            for (int b = 0; b < Cfg::kBoxes; ++b)
              tma_load_3d(dst + b * Cfg::kBoxBytes, map, bar, b * 64, row0, bh, hint);
          }
          __syncwarp();
        };
        load_tile(&mapQ, q_smem, &q_full[0], q0, kEvictFirst);
        const int total = 2 * (seg.kv1 - seg.kv0);
        for (int t = 0; t < total; ++t) {
          const int slot = t % Cfg::kStages;
          const int par = (t / Cfg::kStages) & 1;
          mbar_wait(&kv_empty[slot], par ^ 1, dbg, SITE_TMA_EMPTY, t);
          load_tile((t & 1) ? &mapV : &mapK, kv_smem + slot * Cfg::kTileBytes, &kv_full[slot],
                    (seg.kv0 + (t >> 1)) * Cfg::kKT, kEvictLast);
          if (t == 0 && ntile == 2) load_tile(&mapQ, q_smem + Cfg::kTileBytes, &q_full[1], q0 + 128, kEvictFirst);
        }
      }
    }
  } else {
    // ============================ warps 14 / 15: peer prefetch =========================
    // Idle in a plain attention call (they only hold registers for the softmax warps).  In a ring step they are the
    // exchange: warp 14 pulls K, warp 15 pulls V of the next step, each CTA its 1 / gridDim.x slice of the shard --
    // 134 MB per step of BASELINE config 4 on 8 GPUs is 64 KB per warp, a few loads in flight per lane.  No SM is
    // taken from the attention kernel (NCCL's send/recv CTAs cannot share an SM with a 227 KB CTA), no second
    // stream, no events: the next step's launch follows this one in stream order and finds its shard in place.
    setmaxnreg_dec<48>();
    if (p.pf_words != 0ull) {
      const int lane = threadIdx.x & 31;
      const uint4* const src = p.pf_src[warp - 14];
      uint4* const dst = p.pf_dst[warp - 14];
      const unsigned long long per = (p.pf_words + gridDim.x - 1) / gridDim.x;
      const unsigned long long w0 = per * blockIdx.x;
      const unsigned long long w1 = w0 + per < p.pf_words ? w0 + per : p.pf_words;
      constexpr int kInFlight = 4;  // 16-byte loads in flight per lane: 296 warps x 2 KB cover ~2 us of NVLink latency at 300 GB/s
      for (unsigned long long w = w0 + lane; w < w1; w += 32ull * kInFlight) {
        uint4 r[kInFlight];
#pragma unroll
        for (int u = 0; u < kInFlight; ++u)
          if (w + 32ull * u < w1) r[u] = __ldcv(src + w + 32ull * u);  // .cv: never a stale line of an earlier call
#pragma unroll
        for (int u = 0; u < kInFlight; ++u)
          if (w + 32ull * u < w1) __stcs(dst + w + 32ull * u, r[u]);
      }
    }
  }

  // ---------------------------------- teardown ----------------------------------------
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 8) tmem_dealloc(*tmem_slot, Cfg::kTmemCols);
#undef FA_ROLE_COMMON
}

}  // namespace fa
