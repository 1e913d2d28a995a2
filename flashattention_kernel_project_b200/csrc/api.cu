// C-ABI entry points (include/fa_b200.h) and the thin C++ host layer behind them: argument
// validation, CUtensorMap construction, shared-memory opt-in, launch.  No dispatch table, no CPU
// fallback: on anything but an sm_100 device the compute entry points return FA_ERR_NO_DEVICE.
//
// The reference's host side is a main() per program (constexpr sizes -> cudaMalloc -> H2D ->
// <<<>>> on the default stream -> D2H, e.g. flashattn_warp_spc/flashattn_streaming_16x32_2warp_softmax_v12i.cu:298-393)
// that aborts through CHECK_CUDA (:21-29).  Here the same flow is behind *_host entry points and
// every failure is a return code.
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/fa_b200.h"
#include "fa_fwd.cuh"
#include "fa_small.cuh"
#include "gemm_fused.cuh"
#include "softmax_rows.cuh"

namespace {

using namespace fa;

#define FA_CUDA_TRY(expr)                               \
  do {                                                  \
    cudaError_t e__ = (expr);                           \
    if (e__ != cudaSuccess) return static_cast<int>(e__); \
  } while (0)

constexpr int kMaxDevices = 64;
constexpr int kMaxRingWorld = 64;
std::atomic<unsigned long long> g_launches{0};

// ---- driver entry point: cuTensorMapEncodeTiled without linking libcuda -----------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
  static std::once_flag once;
  static EncodeTiledFn fn = nullptr;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

// Bounded, mutex-protected cache of encoded tensor maps.  A map is a pure function of
// (base, element type, extents, box), so an entry never goes stale; a small-shape call (BASELINE config 1 is 4 us of
// GPU time) otherwise spends more host time in four cuTensorMapEncodeTiled calls than the GPU spends computing.
struct MapKey {
  const void* base;
  uint64_t dims[3];
  uint64_t ld[2];
  uint32_t box[3];
  int dt, rank;
  bool operator==(const MapKey& o) const {
    return base == o.base && dt == o.dt && rank == o.rank && dims[0] == o.dims[0] && dims[1] == o.dims[1] &&
           dims[2] == o.dims[2] && ld[0] == o.ld[0] && ld[1] == o.ld[1] && box[0] == o.box[0] && box[1] == o.box[1] &&
           box[2] == o.box[2];
  }
};
constexpr int kMapCacheEntries = 64;
struct MapCache {
  std::mutex mu;
  MapKey keys[kMapCacheEntries];
  CUtensorMap maps[kMapCacheEntries];
  int used = 0, next = 0;
};
MapCache g_map_cache;
std::atomic<unsigned long long> g_map_hits{0}, g_map_misses{0};

int encode_map(CUtensorMap* map, CUtensorMapDataType dt, int elem_bytes, const void* base, int rank,
               const uint64_t* dims, const uint32_t* box, const uint64_t* ld);

// rank-2/3 tiled map over a row-major tensor whose innermost extent is dims[0]; 128B swizzle.  ld (may be null =
// densely packed): element pitch of dimension 1 (the row pitch) and of dimension 2 (the batch stride).
int make_map(CUtensorMap* map, CUtensorMapDataType dt, int elem_bytes, const void* base, int rank,
             const uint64_t* dims, const uint32_t* box, const uint64_t* ld = nullptr) {
  MapKey key{};
  key.base = base;
  key.dt = static_cast<int>(dt);
  key.rank = rank;
  for (int i = 0; i < rank; ++i) {
    key.dims[i] = dims[i];
    key.box[i] = box[i];
  }
  key.ld[0] = ld != nullptr ? ld[0] : 0;
  key.ld[1] = ld != nullptr ? ld[1] : 0;
  {
    std::lock_guard<std::mutex> lock(g_map_cache.mu);
    for (int i = 0; i < g_map_cache.used; ++i)
      if (g_map_cache.keys[i] == key) {
        *map = g_map_cache.maps[i];
        g_map_hits.fetch_add(1, std::memory_order_relaxed);
        return FA_OK;
      }
  }
  const int rc = encode_map(map, dt, elem_bytes, base, rank, dims, box, ld);
  if (rc != FA_OK) return rc;
  g_map_misses.fetch_add(1, std::memory_order_relaxed);
  std::lock_guard<std::mutex> lock(g_map_cache.mu);
  const int slot = g_map_cache.used < kMapCacheEntries ? g_map_cache.used++ : (g_map_cache.next++ % kMapCacheEntries);
  g_map_cache.keys[slot] = key;
  g_map_cache.maps[slot] = *map;
  return FA_OK;
}

int encode_map(CUtensorMap* map, CUtensorMapDataType dt, int elem_bytes, const void* base, int rank,
               const uint64_t* dims, const uint32_t* box, const uint64_t* ld) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return FA_ERR_DRIVER_ENTRY;
  cuuint64_t gdim[3];
  cuuint64_t gstride[2];
  cuuint32_t bdim[3];
  cuuint32_t estride[3] = {1, 1, 1};
  uint64_t pitch = static_cast<uint64_t>(elem_bytes);
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    pitch *= dims[i];
    if (i < rank - 1) gstride[i] = (ld != nullptr && ld[i] != 0) ? ld[i] * static_cast<uint64_t>(elem_bytes) : pitch;
  }
  const CUresult r = fn(map, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstride, bdim, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? FA_OK : FA_ERR_DRIVER_ENTRY;
}

// ---- per-device state -------------------------------------------------------------------
struct DeviceState {
  bool probed = false;
  bool is_sm100 = false;
  int sm_count = 0;
  bool attr_fa64 = false, attr_fa128 = false, attr_gemm = false, attr_gemm_t = false, attr_gemm_split = false;
  // stream-ordered pool for the attention kernel's tail workspace (allocated and freed on the caller's stream
  // around the launch, so concurrent launches on different streams and captured graphs each own theirs)
  bool pool_tried = false;
  cudaMemPool_t pool = nullptr;
  // ring attention: a high-priority side stream for the K/V exchange and the events that tie it to the caller's stream
  bool ring_ready = false;
  cudaStream_t s_ring = nullptr;
  cudaEvent_t ev_ring_ready = nullptr, ev_ring_recv[kMaxRingWorld], ev_ring_step[kMaxRingWorld];
  // scratch for the *_host entry points
  void* scratch[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t scratch_bytes[6] = {0, 0, 0, 0, 0, 0};
  // the *_host entry points pipeline H2D copy / kernel / D2H copy over chunks of the batch on three streams
  std::mutex host_mu;  // one *_host call at a time per device: they share the scratch buffers and the three streams
  bool pipe_ready = false;
  cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[64], ev_run[64];
};
constexpr int kHostChunksMax = 64;  // events allocated per device; the chunk count actually used is host_chunk_limit()
DeviceState g_dev[kMaxDevices];
std::mutex g_mutex;
unsigned int* g_dbg_host = nullptr;  // mapped pinned page shared by all devices

int current_device_state(DeviceState** out) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return FA_ERR_NO_DEVICE;
  }
  if (dev < 0 || dev >= kMaxDevices) return FA_ERR_NO_DEVICE;
  std::lock_guard<std::mutex> lock(g_mutex);
  DeviceState& st = g_dev[dev];
  if (!st.probed) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
      (void)cudaGetLastError();
      return FA_ERR_NO_DEVICE;
    }
    cudaDeviceGetAttribute(&st.sm_count, cudaDevAttrMultiProcessorCount, dev);
    st.is_sm100 = (major == 10);
    st.probed = true;
  }
  if (st.is_sm100 && !st.pool_tried) {
    st.pool_tried = true;
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    if (cudaMemPoolCreate(&st.pool, &props) == cudaSuccess) {
      unsigned long long keep = ~0ull;  // never hand the (<= 40 MB) workspace back to the driver between launches
      cudaMemPoolSetAttribute(st.pool, cudaMemPoolAttrReleaseThreshold, &keep);
    } else {
      (void)cudaGetLastError();
      st.pool = nullptr;  // no pool: the tail of the grid is not cut (fa_plan)
    }
  }
  if (!st.is_sm100) return FA_ERR_NO_DEVICE;
  if (g_dbg_host == nullptr) {
    void* ptr = nullptr;
    if (cudaHostAlloc(&ptr, 256, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess) {
      std::memset(ptr, 0, 256);
      g_dbg_host = static_cast<unsigned int*>(ptr);
    } else {
      (void)cudaGetLastError();
    }
  }
  *out = &st;
  return FA_OK;
}

DebugChannel debug_channel() {
  DebugChannel d;
  d.words = nullptr;
  if (g_dbg_host != nullptr) {
    void* dptr = nullptr;
    if (cudaHostGetDevicePointer(&dptr, g_dbg_host, 0) == cudaSuccess)
      d.words = static_cast<unsigned int*>(dptr);
    else
      (void)cudaGetLastError();
  }
  return d;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename Kernel>
int ensure_smem_attr(bool* done, Kernel kernel, int bytes) {
  if (*done) return FA_OK;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return static_cast<int>(e);
  *done = true;
  return FA_OK;
}

#ifdef FA_TRACE
long long* g_trace_buf = nullptr;  // developer timeline trace (scripts/trace_fa.py), not in the product build
int g_trace_cta = 0;
#endif

// ---- attention ----------------------------------------------------------------------------
struct FaCall {
  const void *Q, *K, *V;
  void* O16;       // fp16 output (F16 mode, or STATE + is_last)
  float* O32;      // F32 output or O_acc
  float *m, *l;
  int B, H, Nq, Nkv, D;
  float scale;
  int mode, is_first, is_last;
  // fused ring exchange (fa_ring_forward_pull): next step's K / V shard, peer source -> local destination
  const void* pf_src_k = nullptr;
  const void* pf_src_v = nullptr;
  void* pf_dst_k = nullptr;
  void* pf_dst_v = nullptr;
  size_t pf_bytes = 0;
};

// How the grid is cut (fa_fwd.cuh, FaSegment).  The kernel can cut the items of a partial last wave along K/V into
// equal spans of key tiles ("stream-K"); when that pays was measured on B200 (scripts/tail_split_ab.py, both settings
// interleaved in one process, CUDA-graph time):
//   * multi-wave grids (BASELINE configs 2-4, 256 ... 8192 items on 148 SMs): the cut is neutral to 2.6 % SLOWER --
//     the part is power-capped, so the SMs that do run the partial last wave clock higher while the others idle, and
//     512 whole items (3.46 waves, config 3 on 8 GPUs) already take 3.5 item-times, not 4;
//   * grids that fill less than half the SMs (small batch x long sequence, e.g. one (b,h) of N = 4096: 16 items):
//     1.3-1.6x faster, limited by the merge, which costs the last CTA of an item about 2.2 key-tile iterations per
//     part (two 64 KB bulk copies + accumulate).
// So only sub-half-wave grids are cut, into p parts per item with p ~ sqrt(L / 2.2) minimising L/p + 2.2 p
// (L = key tiles per item), 2 <= p <= min(8, SMs / items), spans of at least 4 key tiles, and only if the model
// says the cut wins.  FA_TAIL_COST (default 3) is the fixed cost of a cut span in key-tile iterations.
// The other cut is along QUERY ROWS: an item of a last wave that fills at most half the SMs becomes two CTAs with one
// 128-row tile each.  Nothing is merged and every row is computed exactly as in the whole item (bit-identical
// output); a single-tile CTA runs 2761 cycles per key tile against 3234 for two tiles, so the wave ends ~15 % sooner:
// 32 items of 8 key tiles (B=1,H=8,N=1024,D=64) 16.5 -> 13.7 us, 16 items of 32 key tiles 53.6 -> 49.0 us (K/V cut:
// 44.6), 512 items (config 3 on 8 GPUs) 233.5 -> 232.3 us.
struct FaPlan {
  int items_full, tail_items, tail_ctas;
  int tail_rows;  // the tail is cut along query rows (two single-tile CTAs per item), not along K/V
};
// -1: the environment decides (FA_TAIL_SPLIT = 0 ... 3, default 1); 0: never cut; 1: cut where the model above says
// it pays; 2: cut the partial last wave of ANY grid along K/V into equal spans, whatever the model says (the general
// stream-K plan -- CTAs whose span straddles two items, items shared by up to 8 CTAs); 3: only the row cut.  2 and 3
// are for the tests and the A/B script.
std::atomic<int> g_tail_split{-1};
int tail_split_mode() {
  static const int env_mode = [] {
    const char* e = std::getenv("FA_TAIL_SPLIT");
    const int v = e == nullptr ? 1 : std::atoi(e);
    return v < 0 || v > 3 ? 1 : v;
  }();
  const int v = g_tail_split.load(std::memory_order_relaxed);
  return v < 0 ? env_mode : v;
}
int tail_split_cost() {
  static const int cost = [] {
    const char* e = std::getenv("FA_TAIL_COST");
    return e != nullptr ? std::atoi(e) : 3;
  }();
  return cost;
}
FaPlan fa_plan(long long items, int nkv_tiles, int nq, int sms, bool have_pool) {
  FaPlan plan{static_cast<int>(items), 0, 0, 0};
  const int mode = tail_split_mode();
  if (mode == 0 || sms <= 0) return plan;
  const long long rest = items % sms;
  // Row cut: the items of a last wave that fills at most half the SMs become two single-tile CTAs each.  A CTA with
  // one query tile runs a shorter period per key tile than one with two (2761 against 3234 cycles),
  // every row is computed exactly as before (outputs are bit-identical) and nothing is merged.
  const bool rows_ok = rest > 0 && rest * 2 <= sms && nq > 128;
  auto row_cut = [&] {
    plan.items_full = static_cast<int>(items - rest);
    plan.tail_items = static_cast<int>(rest);
    plan.tail_ctas = static_cast<int>(rest * 2);
    plan.tail_rows = 1;
    return plan;
  };
  if (mode == 3) return rows_ok ? row_cut() : plan;
  if (mode == 2) {
    if (!have_pool) return plan;
    const long long total = rest * nkv_tiles;         // key-tile iterations of the partial wave
    long long n = total / 2 < sms ? total / 2 : sms;  // spans of at least two key tiles ...
    if (n > rest * 8) n = rest * 8;                   // ... and at most 8 parts per item
    if (rest == 0 || n <= rest || (total + 1) * n >= 0x7fffffffll) return plan;
    plan.items_full = static_cast<int>(items - rest);
    plan.tail_items = static_cast<int>(rest);
    plan.tail_ctas = static_cast<int>(n);
    return plan;
  }
  if (items * 2 > sms) return rows_ok ? row_cut() : plan;  // at least half a wave: never cut along K/V (see above)
  // sub-half-wave grid: K/V cut if the model says it beats both the whole item (L) and the row cut (~0.85 L)
  const double L = nkv_tiles;
  const double other = rows_ok ? 0.85 * L : L;
  long long p = static_cast<long long>(std::sqrt(L / 2.2) + 0.5);
  const long long p_max = sms / items < 8 ? sms / items : 8;
  if (p > p_max) p = p_max;
  if (p > nkv_tiles / 4) p = nkv_tiles / 4;
  const long long total = items * nkv_tiles, n = items * p;
  if (!have_pool || p < 2 || L / p + tail_split_cost() + 2.2 * p >= other || (total + 1) * n >= 0x7fffffffll)
    return rows_ok ? row_cut() : plan;  // (the kernel's span arithmetic is 32-bit)
  plan.items_full = 0;
  plan.tail_items = static_cast<int>(items);
  plan.tail_ctas = static_cast<int>(n);
  return plan;
}

template <int D>
int launch_fa(DeviceState* st, const FaCall& c, cudaStream_t stream) {
  using Cfg = FaCfg<D>;
  const uint64_t BH = static_cast<uint64_t>(c.B) * c.H;
  CUtensorMap mq, mk, mv, mo, mo32;
  const uint32_t box[3] = {64, 128, 1};
  const uint64_t dq[3] = {static_cast<uint64_t>(D), static_cast<uint64_t>(c.Nq), BH};
  const uint64_t dk[3] = {static_cast<uint64_t>(D), static_cast<uint64_t>(c.Nkv), BH};
  int rc;
  if ((rc = make_map(&mq, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, c.Q, 3, dq, box)) != FA_OK) return rc;
  if ((rc = make_map(&mk, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, c.K, 3, dk, box)) != FA_OK) return rc;
  if ((rc = make_map(&mv, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, c.V, 3, dk, box)) != FA_OK) return rc;
  if (c.O16 != nullptr) {
    if ((rc = make_map(&mo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, c.O16, 3, dq, box)) != FA_OK) return rc;
  } else {
    mo = mq;  // never dereferenced in the modes that pass no fp16 output
  }
  if (c.O32 != nullptr) {  // fp32 output / carried O_acc: boxes of 32 fp32 columns (one 128-byte swizzle row) x 128 rows
    const uint32_t box32[3] = {32, 128, 1};
    if ((rc = make_map(&mo32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, c.O32, 3, dq, box32)) != FA_OK) return rc;
  } else {
    mo32 = mq;  // never dereferenced in FA_MODE_F16
  }
  bool* flag = (D == 64) ? &st->attr_fa64 : &st->attr_fa128;
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    if ((rc = ensure_smem_attr(flag, fa_fwd_kernel<D>, Cfg::kSmemBytes)) != FA_OK) return rc;
  }
  FaParams p;
  p.nq = c.Nq;
  p.nkv = c.Nkv;
  p.q_blocks = (c.Nq + 255) / 256;
  float sl2 = c.scale * 1.4426950408889634f;
  if (fabsf(sl2) < 1e-30f) sl2 = (sl2 < 0.f || (sl2 == 0.f && std::signbit(sl2))) ? -1e-30f : 1e-30f;
  p.scale_log2 = sl2;
  p.mode = c.mode;
  p.is_first = c.is_first;
  p.is_last = c.is_last;
  p.o_f32 = c.O32;
  p.m_state = c.m;
  p.l_state = c.l;
  p.dbg = debug_channel();
#ifdef FA_TRACE
  p.trace = g_trace_buf;
  p.trace_cta = g_trace_cta;
#endif
  const uint64_t items = BH * static_cast<uint64_t>(p.q_blocks);
  if (items == 0 || items > 0x7fffffffull) return FA_ERR_BAD_SHAPE;
  const FaPlan plan = fa_plan(static_cast<long long>(items), (c.Nkv + Cfg::kKT - 1) / Cfg::kKT, c.Nq, st->sm_count,
                              st->pool != nullptr);
  p.items_full = plan.items_full;
  p.tail_items = plan.tail_items;
  p.tail_ctas = plan.tail_ctas;
  p.tail_rows = plan.tail_rows;
  p.ws = nullptr;
  p.ws_count = nullptr;
  p.pf_src[0] = static_cast<const uint4*>(c.pf_src_k);
  p.pf_src[1] = static_cast<const uint4*>(c.pf_src_v);
  p.pf_dst[0] = static_cast<uint4*>(c.pf_dst_k);
  p.pf_dst[1] = static_cast<uint4*>(c.pf_dst_v);
  p.pf_words = c.pf_bytes / 16;
  void* ws = nullptr;
  if (plan.tail_ctas > 0 && !plan.tail_rows) {
    // [counters, padded to 256 B][tail_ctas * 2 slots of (D + 2) * 256 floats]
    const size_t count_bytes = (static_cast<size_t>(plan.tail_items) * 4 + 255) & ~static_cast<size_t>(255);
    const size_t bytes = count_bytes + static_cast<size_t>(plan.tail_ctas) * 2 * (D + 2) * 256 * 4;
    cudaError_t e = cudaMallocFromPoolAsync(&ws, bytes, st->pool, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(ws, 0, count_bytes, stream);
    if (e != cudaSuccess) {
      if (ws != nullptr) cudaFreeAsync(ws, stream);
      return static_cast<int>(e);
    }
    p.ws_count = static_cast<unsigned int*>(ws);
    p.ws = reinterpret_cast<float*>(static_cast<char*>(ws) + count_bytes);
  }
  const unsigned grid = static_cast<unsigned>(plan.items_full + plan.tail_ctas);
  fa_fwd_kernel<D><<<grid, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(mq, mk, mv, mo, mo32, p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (ws != nullptr) {
    const cudaError_t ef = cudaFreeAsync(ws, stream);
    if (e == cudaSuccess) e = ef;
  }
  return e == cudaSuccess ? FA_OK : static_cast<int>(e);
}

int run_fa(const FaCall& c, void* stream) {
  if (c.B <= 0 || c.H <= 0 || c.Nq <= 0 || c.Nkv <= 0) return FA_ERR_BAD_SHAPE;
  if (c.D != 64 && c.D != 128) return FA_ERR_BAD_HEAD_DIM;
  if (!(c.scale == c.scale) || std::isinf(c.scale)) return FA_ERR_BAD_ARGUMENT;
  if (c.Q == nullptr || c.K == nullptr || c.V == nullptr) return FA_ERR_NULL_POINTER;
  if (c.mode == FA_MODE_F16 && c.O16 == nullptr) return FA_ERR_NULL_POINTER;
  if (c.mode == FA_MODE_F32 && c.O32 == nullptr) return FA_ERR_NULL_POINTER;
  if (c.mode == FA_MODE_STATE) {
    if (c.O32 == nullptr || c.m == nullptr || c.l == nullptr) return FA_ERR_NULL_POINTER;
    if (c.is_last && c.O16 == nullptr) return FA_ERR_NULL_POINTER;
    if (!aligned16(c.m) || !aligned16(c.l)) return FA_ERR_MISALIGNED;
  }
  if (!aligned16(c.Q) || !aligned16(c.K) || !aligned16(c.V) || !aligned16(c.O16) || !aligned16(c.O32))
    return FA_ERR_MISALIGNED;
  const uint64_t BH = static_cast<uint64_t>(c.B) * c.H;
  if (BH > 0x7fffffffull) return FA_ERR_BAD_SHAPE;
  DeviceState* st = nullptr;
  int rc = current_device_state(&st);
  if (rc != FA_OK) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return c.D == 64 ? launch_fa<64>(st, c, s) : launch_fa<128>(st, c, s);
}

// ---- GEMM ---------------------------------------------------------------------------------
// C[z] = act(alpha * A[z] * op(B[z]) + bias) + residual[z] for z in [0, batch)
struct GemmCall {
  const void* A;
  const void* Bm;
  const float* bias;
  const float* residual;
  float* C;
  int batch, M, N, K;
  bool trans_b;
  int act;
  float alpha;
  // element pitches; 0 = densely packed.  lda/ldb/ldc: row pitch of A, B (its stored rows: K-rows of [K,N], N-rows of
  // [N,K]) and C / residual; sa/sb/sc: distance between consecutive batch entries.
  long long lda = 0, ldb = 0, ldc = 0, sa = 0, sb = 0, sc = 0;
};

int run_gemm(const GemmCall& c, void* stream) {
  const int M = c.M, N = c.N, K = c.K;
  if (M <= 0 || N <= 0 || K <= 0 || c.batch <= 0) return FA_ERR_BAD_SHAPE;
  if (c.A == nullptr || c.Bm == nullptr || c.C == nullptr) return FA_ERR_NULL_POINTER;
  if ((N & 7) != 0 || (K & 7) != 0) return FA_ERR_MISALIGNED;
  if (!aligned16(c.A) || !aligned16(c.Bm) || !aligned16(c.C) || !aligned16(c.bias) || !aligned16(c.residual))
    return FA_ERR_MISALIGNED;
  if (c.act < GEMM_ACT_IDENTITY || c.act > GEMM_ACT_SILU) return FA_ERR_BAD_ARGUMENT;
  // leading dimensions: at least the row they hold, 16-byte multiples (TMA global strides)
  const long long b_row = c.trans_b ? K : N, b_rows = c.trans_b ? N : K;
  if (c.lda < 0 || c.ldb < 0 || c.ldc < 0 || c.sa < 0 || c.sb < 0 || c.sc < 0) return FA_ERR_BAD_ARGUMENT;
  if ((c.lda != 0 && c.lda < K) || (c.ldb != 0 && c.ldb < b_row) || (c.ldc != 0 && c.ldc < N)) return FA_ERR_BAD_ARGUMENT;
  if ((c.lda & 7) || (c.ldb & 7) || (c.ldc & 3) || (c.sa & 7) || (c.sb & 7) || (c.sc & 3)) return FA_ERR_MISALIGNED;
  const uint64_t lda = c.lda ? c.lda : K, ldb = c.ldb ? c.ldb : b_row, ldc = c.ldc ? c.ldc : N;
  const uint64_t sa = c.sa ? c.sa : lda * M, sb = c.sb ? c.sb : ldb * b_rows, sc = c.sc ? c.sc : ldc * M;
  const uint64_t ld_a[2] = {lda, sa}, ld_b[2] = {ldb, sb}, ld_c[2] = {ldc, sc};
  DeviceState* st = nullptr;
  int rc = current_device_state(&st);
  if (rc != FA_OK) return rc;
  const uint64_t Z = static_cast<uint64_t>(c.batch);
  CUtensorMap ma, mb, mr, mc;
  const uint64_t da[3] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M), Z};
  const uint32_t ba[3] = {64, 128, 1};
  if ((rc = make_map(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, c.A, 3, da, ba, ld_a)) != FA_OK) return rc;
  if (c.trans_b) {  // [N,K] rows: the same K-major boxes as A
    const uint64_t db[3] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N), Z};
    if ((rc = make_map(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, c.Bm, 3, db, ba, ld_b)) != FA_OK) return rc;
  } else {
    const uint64_t db[3] = {static_cast<uint64_t>(N), static_cast<uint64_t>(K), Z};
    const uint32_t bb[3] = {64, 64, 1};
    if ((rc = make_map(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, c.Bm, 3, db, bb, ld_b)) != FA_OK) return rc;
  }
  // epilogue blocks: 128 rows x 32 fp32 columns of C / residual
  const uint64_t dc[3] = {static_cast<uint64_t>(N), static_cast<uint64_t>(M), Z};
  const uint32_t bc[3] = {GemmCfg::kEpiCols, GemmCfg::kBM, 1};
  if ((rc = make_map(&mc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, c.C, 3, dc, bc, ld_c)) != FA_OK) return rc;
  if (c.residual != nullptr) {  // the residual shares C's layout
    if ((rc = make_map(&mr, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, c.residual, 3, dc, bc, ld_c)) != FA_OK) return rc;
  } else {
    mr = mc;  // never dereferenced without a residual
  }
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    if ((rc = ensure_smem_attr(&st->attr_gemm, gemm_fused_kernel<false>, GemmCfg::kSmemBytes)) != FA_OK) return rc;
    if ((rc = ensure_smem_attr(&st->attr_gemm_t, gemm_fused_kernel<true>, GemmCfg::kSmemBytes)) != FA_OK) return rc;
  }
  GemmParams p;
  p.M = M;
  p.N = N;
  p.K = K;
  p.batch = c.batch;
  p.m_blocks = (M + 255) / 256;  // 256 x 256 output tiles, one per CTA pair
  p.n_blocks = (N + GemmCfg::kBN - 1) / GemmCfg::kBN;
  p.alpha = c.alpha;
  p.act = c.act;
  p.bias = c.bias;
  p.residual = c.residual;
  p.C = c.C;
  p.dbg = debug_channel();
  const long long tiles = static_cast<long long>(p.m_blocks) * p.n_blocks * c.batch;
  if (tiles > 0x7fffffffll) return FA_ERR_BAD_SHAPE;
  const long long pairs_max = st->sm_count / 2;  // one CTA pair (cluster of 2) per TPC, persistent
  const int grid = 2 * static_cast<int>(tiles < pairs_max ? tiles : pairs_max);
  const cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (c.trans_b)
    gemm_fused_kernel<true><<<grid, GemmCfg::kThreads, GemmCfg::kSmemBytes, s>>>(ma, mb, mr, mc, ma, mb, p);
  else
    gemm_fused_kernel<false><<<grid, GemmCfg::kThreads, GemmCfg::kSmemBytes, s>>>(ma, mb, mr, mc, ma, mb, p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? FA_OK : static_cast<int>(e);
}

int run_gemm(const void* A, const void* Bm, const float* bias, const float* residual, float* C, int M, int N, int K,
             void* stream) {
  const GemmCall c{A, Bm, bias, residual, C, 1, M, N, K, false, GEMM_ACT_RELU, 1.0f};
  return run_gemm(c, stream);
}

// fp32 operands through the tf32 tensor path, split into (hi, lo) pairs ("3xTF32"): see GemmCfgT in gemm_fused.cuh.
// Scratch for the four split operands (2 x (M*K + N*K) floats) comes from the stream-ordered pool and is released on
// the stream after the GEMM.
int run_gemm_f32(const float* A, const float* Bm, const float* bias, const float* residual, float* C, int M, int N,
                 int K, int act, void* stream) {
  using Cfg = GemmCfgT<true>;
  if (M <= 0 || N <= 0 || K <= 0) return FA_ERR_BAD_SHAPE;
  if (A == nullptr || Bm == nullptr || C == nullptr) return FA_ERR_NULL_POINTER;
  if ((N & 3) != 0 || (K & 3) != 0) return FA_ERR_MISALIGNED;  // fp32 rows must be 16-byte multiples
  if (!aligned16(A) || !aligned16(Bm) || !aligned16(C) || !aligned16(bias) || !aligned16(residual)) return FA_ERR_MISALIGNED;
  if (act < GEMM_ACT_IDENTITY || act > GEMM_ACT_SILU) return FA_ERR_BAD_ARGUMENT;
  DeviceState* st = nullptr;
  int rc = current_device_state(&st);
  if (rc != FA_OK) return rc;
  if (st->pool == nullptr) return FA_ERR_NO_DEVICE;
  const cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t a_elems = static_cast<size_t>(M) * K, b_elems = static_cast<size_t>(N) * K;
  const size_t a_bytes = (a_elems * 4 + 255) & ~static_cast<size_t>(255), b_bytes = (b_elems * 4 + 255) & ~static_cast<size_t>(255);
  void* scratch = nullptr;
  FA_CUDA_TRY(cudaMallocFromPoolAsync(&scratch, 2 * a_bytes + 2 * b_bytes, st->pool, s));
  float* const a_hi = static_cast<float*>(scratch);
  float* const a_lo = reinterpret_cast<float*>(static_cast<char*>(scratch) + a_bytes);
  float* const b_hi = reinterpret_cast<float*>(static_cast<char*>(scratch) + 2 * a_bytes);
  float* const b_lo = reinterpret_cast<float*>(static_cast<char*>(scratch) + 2 * a_bytes + b_bytes);
  auto fail = [&](int code) {
    cudaFreeAsync(scratch, s);
    return code;
  };
  {
    const size_t n4 = a_elems / 4;
    const unsigned blocks = static_cast<unsigned>(std::min<size_t>((n4 + 255) / 256, static_cast<size_t>(st->sm_count) * 16));
    gemm_split_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(A), reinterpret_cast<float4*>(a_hi),
                                             reinterpret_cast<float4*>(a_lo), n4);
    const dim3 tgrid((N + 31) / 32, (K + 31) / 32);
    gemm_split_transpose_kernel<<<tgrid, dim3(32, 8), 0, s>>>(Bm, b_hi, b_lo, K, N);
    g_launches.fetch_add(2, std::memory_order_relaxed);
  }
  CUtensorMap ma, mb, ma2, mb2, mr, mc;
  const uint64_t da[3] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M), 1};
  const uint64_t db[3] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N), 1};
  const uint32_t box[3] = {32, 128, 1};  // 32 fp32 = one 128-byte swizzle row
  if ((rc = make_map(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a_hi, 3, da, box)) != FA_OK) return fail(rc);
  if ((rc = make_map(&ma2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a_lo, 3, da, box)) != FA_OK) return fail(rc);
  if ((rc = make_map(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, b_hi, 3, db, box)) != FA_OK) return fail(rc);
  if ((rc = make_map(&mb2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, b_lo, 3, db, box)) != FA_OK) return fail(rc);
  const uint64_t dc[3] = {static_cast<uint64_t>(N), static_cast<uint64_t>(M), 1};
  const uint32_t bc[3] = {Cfg::kEpiCols, Cfg::kBM, 1};
  if ((rc = make_map(&mc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, C, 3, dc, bc)) != FA_OK) return fail(rc);
  if (residual != nullptr) {
    if ((rc = make_map(&mr, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, residual, 3, dc, bc)) != FA_OK) return fail(rc);
  } else {
    mr = mc;
  }
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    if ((rc = ensure_smem_attr(&st->attr_gemm_split, gemm_fused_kernel<true, true>, Cfg::kSmemBytes)) != FA_OK)
      return fail(rc);
  }
  GemmParams p;
  p.M = M;
  p.N = N;
  p.K = K;
  p.batch = 1;
  p.m_blocks = (M + 255) / 256;
  p.n_blocks = (N + Cfg::kBN - 1) / Cfg::kBN;
  p.alpha = 1.0f;
  p.act = act;
  p.bias = bias;
  p.residual = residual;
  p.C = C;
  p.dbg = debug_channel();
  const long long tiles = static_cast<long long>(p.m_blocks) * p.n_blocks;
  const long long pairs_max = st->sm_count / 2;
  const int grid = 2 * static_cast<int>(tiles < pairs_max ? tiles : pairs_max);
  gemm_fused_kernel<true, true><<<grid, Cfg::kThreads, Cfg::kSmemBytes, s>>>(ma, mb, mr, mc, ma2, mb2, p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  const cudaError_t ef = cudaFreeAsync(scratch, s);
  if (e == cudaSuccess) e = ef;
  return e == cudaSuccess ? FA_OK : static_cast<int>(e);
}

// ---- softmax ------------------------------------------------------------------------------
template <int TPR, int V4, typename TOut>
void launch_softmax_reg(const float* X, TOut* P, int rows, int N, int sm_count, cudaStream_t s) {
  constexpr int kGroups = 256 / TPR;
  long long ctas = (static_cast<long long>(rows) + kGroups - 1) / kGroups;
  const long long cap = static_cast<long long>(sm_count) * 8;  // up to 8 resident CTAs per SM, grid-stride beyond
  if (ctas > cap) ctas = cap;
  softmax_rows_reg_kernel<TPR, V4, TOut><<<static_cast<unsigned>(ctas), 256, 0, s>>>(X, P, rows, N);
}

template <typename TOut>
int run_softmax(const float* X, TOut* P, int rows, int N, void* stream) {
  if (rows <= 0 || N <= 0) return FA_ERR_BAD_SHAPE;
  if (X == nullptr || P == nullptr) return FA_ERR_NULL_POINTER;
  if ((reinterpret_cast<uintptr_t>(X) & 3u) || (reinterpret_cast<uintptr_t>(P) & (sizeof(TOut) - 1))) return FA_ERR_MISALIGNED;
  DeviceState* st = nullptr;
  int rc = current_device_state(&st);
  if (rc != FA_OK) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // 128-bit loads of X and 4-element stores of P (16 B of float, 8 B of half)
  const bool vec_ok = (N % 4 == 0) && aligned16(X) && (reinterpret_cast<uintptr_t>(P) & (4 * sizeof(TOut) - 1)) == 0;
  const int sm = st->sm_count;
  if (vec_ok && N <= 128) launch_softmax_reg<32, 1>(X, P, rows, N, sm, s);
  else if (vec_ok && N <= 256) launch_softmax_reg<32, 2>(X, P, rows, N, sm, s);
  else if (vec_ok && N <= 512) launch_softmax_reg<32, 4>(X, P, rows, N, sm, s);
  else if (vec_ok && N <= 1024) launch_softmax_reg<32, 8>(X, P, rows, N, sm, s);
  else if (vec_ok && N <= 2048) launch_softmax_reg<256, 2>(X, P, rows, N, sm, s);
  else if (vec_ok && N <= 4096) launch_softmax_reg<256, 4>(X, P, rows, N, sm, s);
  else if (vec_ok && N <= 8192) launch_softmax_reg<256, 8>(X, P, rows, N, sm, s);
  else {
    long long ctas = rows;
    const long long cap = static_cast<long long>(sm) * 8;
    if (ctas > cap) ctas = cap;
    softmax_rows_3pass_kernel<TOut><<<static_cast<unsigned>(ctas), 256, 0, s>>>(X, P, rows, N);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? FA_OK : static_cast<int>(e);
}

// ---- host-buffer variants -----------------------------------------------------------------
int scratch_get(DeviceState* st, int slot, size_t bytes, void** out) {
  if (st->scratch_bytes[slot] < bytes) {
    if (st->scratch[slot] != nullptr) cudaFree(st->scratch[slot]);
    st->scratch[slot] = nullptr;
    st->scratch_bytes[slot] = 0;
    cudaError_t e = cudaMalloc(&st->scratch[slot], bytes);
    if (e != cudaSuccess) return static_cast<int>(e);
    st->scratch_bytes[slot] = bytes;
  }
  *out = st->scratch[slot];
  return FA_OK;
}


// Inside the chunk loops of the *_host entry points: a failing call must not leave earlier chunks' copies running
// against the caller's buffers (and the scratch, which the next call reuses) -- drain the pipeline, then return.
#define FA_PIPE_TRY(st, expr)                           \
  do {                                                  \
    cudaError_t e__ = (expr);                           \
    if (e__ != cudaSuccess) {                           \
      host_pipe_drain(st);                              \
      return static_cast<int>(e__);                     \
    }                                                   \
  } while (0)

// Streams and events of the host-buffer pipeline, created once per device.
int host_pipe_init(DeviceState* st) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (st->pipe_ready) return FA_OK;
  FA_CUDA_TRY(cudaStreamCreateWithFlags(&st->s_in, cudaStreamNonBlocking));
  FA_CUDA_TRY(cudaStreamCreateWithFlags(&st->s_run, cudaStreamNonBlocking));
  FA_CUDA_TRY(cudaStreamCreateWithFlags(&st->s_out, cudaStreamNonBlocking));
  for (int i = 0; i < kHostChunksMax; ++i) {
    FA_CUDA_TRY(cudaEventCreateWithFlags(&st->ev_in[i], cudaEventDisableTiming));
    FA_CUDA_TRY(cudaEventCreateWithFlags(&st->ev_run[i], cudaEventDisableTiming));
  }
  st->pipe_ready = true;
  return FA_OK;
}
// Everything enqueued by a *_host call has finished when this returns (also on its error paths, so that no copy is
// still reading or writing the caller's buffers).
int host_pipe_drain(DeviceState* st) {
  const cudaError_t a = cudaStreamSynchronize(st->s_in), b = cudaStreamSynchronize(st->s_run),
                    c = cudaStreamSynchronize(st->s_out);
  if (a != cudaSuccess) return static_cast<int>(a);
  if (b != cudaSuccess) return static_cast<int>(b);
  return c == cudaSuccess ? FA_OK : static_cast<int>(c);
}
// Upper bound on the chunks of one *_host call: 16 unless FA_HOST_CHUNKS (1..64) says otherwise.
int host_chunk_limit() {
  static const int limit = [] {
    const char* e = std::getenv("FA_HOST_CHUNKS");
    const int v = e != nullptr ? std::atoi(e) : 0;
    return (v >= 1 && v <= kHostChunksMax) ? v : 16;
  }();
  return limit;
}
// As many chunks as host_chunk_limit() allows (16) while a copy stays at least FA_HOST_CHUNK_MB (default 2) MB of one
// tensor, at most one per unit.  Measured on 8xB200, head shards of BASELINE config 3, settings interleaved in one job
// (scripts/e2e_chunk_sweep.py; ms per call by number of chunks):
//   1 GPU, 256 MB per tensor:  8 chunks 16.20, 16 chunks 16.31, 32 / 64 chunks 17.06
//   2 GPUs, 128 MB:            4 chunks 9.12, 8 chunks 8.99, 16 chunks 9.11, 64 chunks 9.77
//   4 GPUs, 64 MB:             2 chunks 10.37, 4 chunks 9.98, 8 chunks 9.70, 16 chunks 9.40
//   8 GPUs, 32 MB:             1 chunk 7.18, 2 chunks 6.96, 8 chunks 6.80, 16 chunks 6.39
// The ranks of a job run in lockstep, so with few chunks every rank copies in the same direction at the same time
// and the host's other direction idles; 8-16 chunks interleave the directions, 64 pay per-copy cost.
int host_chunk_min_bytes_log2() {  // read per call: a tuning script changes it between calls (scripts/e2e_chunk_sweep.py)
  const char* e = std::getenv("FA_HOST_CHUNK_MB");
  const int mb = e != nullptr ? std::atoi(e) : 2;
  int s = 20;
  while ((1 << (s - 20)) < mb && s < 30) ++s;
  return s;
}
int host_chunks(size_t tensor_bytes, long long units) {
  long long n = static_cast<long long>(tensor_bytes >> host_chunk_min_bytes_log2());
  if (n > host_chunk_limit()) n = host_chunk_limit();
  if (n > units) n = units;
  return n < 1 ? 1 : static_cast<int>(n);
}


// ---- ring / context-parallel attention: NCCL resolved at run time ---------------------------------------------
// The library does not link NCCL: a process that never calls fa_ring_forward (every C program above) needs none.
// The symbols come from the libnccl.so.2 already loaded into the process (the one that created the caller's
// communicator -- torch's bundled copy under torchrun), else from the system's.
struct NcclApi {
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  bool ok = false;
};
const NcclApi& nccl_api() {
  static NcclApi api = [] {
    NcclApi a;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (h == nullptr) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) return a;
    a.GroupStart = reinterpret_cast<int (*)()>(dlsym(h, "ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<int (*)()>(dlsym(h, "ncclGroupEnd"));
    a.Send = reinterpret_cast<int (*)(const void*, size_t, int, int, void*, cudaStream_t)>(dlsym(h, "ncclSend"));
    a.Recv = reinterpret_cast<int (*)(void*, size_t, int, int, void*, cudaStream_t)>(dlsym(h, "ncclRecv"));
    a.ok = a.GroupStart != nullptr && a.GroupEnd != nullptr && a.Send != nullptr && a.Recv != nullptr;
    return a;
  }();
  return api;
}
constexpr int kNcclUint8 = 1;  // ncclDataType_t: ncclInt8 = 0, ncclUint8 = 1 (nccl.h); shards travel as bytes

int ring_init(DeviceState* st) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (st->ring_ready) return FA_OK;
  int lo = 0, hi = 0;
  FA_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  // The attention kernel holds every SM with one 227 KB CTA; NCCL's send/recv kernel only gets SMs while it runs if
  // its stream outranks the compute stream (measured: 1.5 ms vs 14-19 ms for a 512 MiB exchange next to a 15 ms launch).
  FA_CUDA_TRY(cudaStreamCreateWithPriority(&st->s_ring, cudaStreamNonBlocking, hi));
  FA_CUDA_TRY(cudaEventCreateWithFlags(&st->ev_ring_ready, cudaEventDisableTiming));
  for (int i = 0; i < kMaxRingWorld; ++i) {
    FA_CUDA_TRY(cudaEventCreateWithFlags(&st->ev_ring_recv[i], cudaEventDisableTiming));
    FA_CUDA_TRY(cudaEventCreateWithFlags(&st->ev_ring_step[i], cudaEventDisableTiming));
  }
  st->ring_ready = true;
  return FA_OK;
}

struct RingLayout {
  size_t state_o, state_m, state_l, kv, shard_bytes, total;
  int depth;
};
// [O_acc fp32 | m | l | depth x (K shard, V shard)], every part 256-byte aligned
RingLayout ring_layout(int B, int H, int n, int D, int world, size_t workspace_bytes) {
  auto up = [](size_t x) { return (x + 255) & ~static_cast<size_t>(255); };
  RingLayout L{};
  const size_t rows = static_cast<size_t>(B) * H * n;
  L.shard_bytes = up(rows * D * 2);
  L.state_o = 0;
  L.state_m = up(rows * D * 4);
  L.state_l = L.state_m + up(rows * 4);
  L.kv = L.state_l + up(rows * 4);
  const int want = world - 1;
  int depth = want;
  if (workspace_bytes != 0) {
    const size_t room = workspace_bytes > L.kv ? workspace_bytes - L.kv : 0;
    const size_t fit = room / (2 * L.shard_bytes);
    depth = fit < static_cast<size_t>(want) ? static_cast<int>(fit) : want;
  }
  L.depth = depth;
  L.total = L.kv + static_cast<size_t>(depth) * 2 * L.shard_bytes;
  return L;
}

}  // namespace

// =========================================== C ABI ==========================================
extern "C" {

int fa_version(void) { return 200; }

int fa_init(void) {
  DeviceState* st = nullptr;
  return current_device_state(&st);
}

const char* fa_strerror(int code) {
  switch (code) {
    case FA_OK: return "ok";
    case FA_ERR_BAD_SHAPE: return "bad shape (non-positive extent or index overflow)";
    case FA_ERR_BAD_HEAD_DIM: return "head dim must be 64 or 128";
    case FA_ERR_MISALIGNED: return "pointer or row pitch not 16-byte aligned";
    case FA_ERR_NULL_POINTER: return "required pointer is null";
    case FA_ERR_NO_DEVICE: return "current device is not an sm_100 (B200) GPU; there is no CPU fallback";
    case FA_ERR_DRIVER_ENTRY: return "cuTensorMapEncodeTiled unavailable or failed";
    case FA_ERR_BAD_ARGUMENT: return "bad argument";
    case FA_ERR_NCCL: return "NCCL call failed";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "unknown error";
}

int fa_forward(const void* Q, const void* K, const void* V, void* O, int B, int H, int N, int D, float scale,
               void* stream) {
  FaCall c{Q, K, V, O, nullptr, nullptr, nullptr, B, H, N, N, D, scale, FA_MODE_F16, 1, 1};
  return run_fa(c, stream);
}

int fa_forward_f32out(const void* Q, const void* K, const void* V, float* O, int B, int H, int N, int D, float scale,
                      void* stream) {
  FaCall c{Q, K, V, nullptr, O, nullptr, nullptr, B, H, N, N, D, scale, FA_MODE_F32, 1, 1};
  return run_fa(c, stream);
}

int fa_forward_partial(const void* Q, const void* K, const void* V, float* O_acc, float* m, float* l, int B, int H,
                       int Nq, int Nkv, int D, float scale, int is_first, int is_last, void* O_final, void* stream) {
  FaCall c{Q, K, V, is_last ? O_final : nullptr, O_acc, m, l, B, H, Nq, Nkv, D, scale, FA_MODE_STATE,
           is_first ? 1 : 0, is_last ? 1 : 0};
  return run_fa(c, stream);
}

size_t fa_ring_workspace_bytes(int B, int H, int n, int D, int world, int depth) {
  if (B <= 0 || H <= 0 || n <= 0 || D <= 0 || world <= 0) return 0;
  RingLayout L = ring_layout(B, H, n, D, world, 0);
  if (depth > 0 && depth < L.depth) L.total = L.kv + static_cast<size_t>(depth) * 2 * L.shard_bytes;
  return L.total;
}

int fa_ring_forward(const void* Q, const void* K, const void* V, void* O, int B, int H, int n, int D, float scale,
                    int rank, int world, void* nccl_comm, void* workspace, size_t workspace_bytes, void* stream) {
  if (B <= 0 || H <= 0 || n <= 0) return FA_ERR_BAD_SHAPE;
  if (D != 64 && D != 128) return FA_ERR_BAD_HEAD_DIM;
  if (world <= 0 || world > kMaxRingWorld || rank < 0 || rank >= world) return FA_ERR_BAD_ARGUMENT;
  if (Q == nullptr || K == nullptr || V == nullptr || O == nullptr || workspace == nullptr) return FA_ERR_NULL_POINTER;
  if (world > 1 && nccl_comm == nullptr) return FA_ERR_NULL_POINTER;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return FA_ERR_MISALIGNED;
  const RingLayout L = ring_layout(B, H, n, D, world, workspace_bytes);
  if (workspace_bytes < L.kv || (world > 1 && L.depth < 1)) return FA_ERR_BAD_ARGUMENT;
  DeviceState* st = nullptr;
  int rc = current_device_state(&st);
  if (rc != FA_OK) return rc;
  char* const ws = static_cast<char*>(workspace);
  float* const o_acc = reinterpret_cast<float*>(ws + L.state_o);
  float* const m = reinterpret_cast<float*>(ws + L.state_m);
  float* const l = reinterpret_cast<float*>(ws + L.state_l);
  const cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (world == 1) return fa_forward_partial(Q, K, V, o_acc, m, l, B, H, n, n, D, scale, 1, 1, O, stream);

  const NcclApi& nccl = nccl_api();
  if (!nccl.ok) return FA_ERR_DRIVER_ENTRY;
  if ((rc = ring_init(st)) != FA_OK) return rc;
  std::lock_guard<std::mutex> ring_lock(st->host_mu);  // one enqueue at a time per device: the events are shared
  const size_t bytes = static_cast<size_t>(B) * H * n * D * 2;
  const int depth = L.depth;
  auto kbuf = [&](int slot) { return ws + L.kv + static_cast<size_t>(slot) * 2 * L.shard_bytes; };
  // Exchange x (1 <= x < world): this rank's OWN shard goes to rank + x and the shard of rank - x arrives.  K/V are
  // read-only, so nothing is passed along: the world - 1 exchanges do not depend on each other and are enqueued as
  // far ahead as the workspace has buffers (all of them with the default size), each one grouped send/recv call.
  auto exchange = [&](int x) -> int {
    const int to = (rank + x) % world, from = (rank - x + world) % world;
    char* const dst = kbuf((x - 1) % depth);
    int e = nccl.GroupStart();
    if (e == 0) e = nccl.Send(K, bytes, kNcclUint8, to, nccl_comm, st->s_ring);
    if (e == 0) e = nccl.Send(V, bytes, kNcclUint8, to, nccl_comm, st->s_ring);
    if (e == 0) e = nccl.Recv(dst, bytes, kNcclUint8, from, nccl_comm, st->s_ring);
    if (e == 0) e = nccl.Recv(dst + L.shard_bytes, bytes, kNcclUint8, from, nccl_comm, st->s_ring);
    const int g = nccl.GroupEnd();
    if (e == 0) e = g;
    if (e != 0) return FA_ERR_NCCL;
    FA_CUDA_TRY(cudaEventRecord(st->ev_ring_recv[x], st->s_ring));
    return FA_OK;
  };
  // K, V (and the workspace, whatever used it last) are ready where the caller's stream stands now
  FA_CUDA_TRY(cudaEventRecord(st->ev_ring_ready, s));
  FA_CUDA_TRY(cudaStreamWaitEvent(st->s_ring, st->ev_ring_ready, 0));
  for (int x = 1; x <= depth && x < world; ++x)
    if ((rc = exchange(x)) != FA_OK) return rc;
  // step 0 needs no exchange: the first exchanges run under it
  if ((rc = fa_forward_partial(Q, K, V, o_acc, m, l, B, H, n, n, D, scale, 1, 0, nullptr, stream)) != FA_OK) return rc;
  for (int x = 1; x < world; ++x) {
    char* const src = kbuf((x - 1) % depth);
    FA_CUDA_TRY(cudaStreamWaitEvent(s, st->ev_ring_recv[x], 0));
    if ((rc = fa_forward_partial(Q, src, src + L.shard_bytes, o_acc, m, l, B, H, n, n, D, scale, 0, x == world - 1 ? 1 : 0,
                                 O, stream)) != FA_OK)
      return rc;
    if (x + depth < world) {  // the buffer step x just read is the one exchange x + depth fills
      FA_CUDA_TRY(cudaEventRecord(st->ev_ring_step[x], s));
      FA_CUDA_TRY(cudaStreamWaitEvent(st->s_ring, st->ev_ring_step[x], 0));
      if ((rc = exchange(x + depth)) != FA_OK) return rc;
    }
  }
  return FA_OK;
}

int fa_ipc_export(const void* ptr, unsigned char* handle, unsigned long long* offset) {
  if (ptr == nullptr || handle == nullptr || offset == nullptr) return FA_ERR_NULL_POINTER;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the header promises 64-byte handles");
  using RangeFn = CUresult (*)(CUdeviceptr*, size_t*, CUdeviceptr);
  static RangeFn range_fn = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      fn = nullptr;
    return reinterpret_cast<RangeFn>(fn);
  }();
  if (range_fn == nullptr) return FA_ERR_DRIVER_ENTRY;
  CUdeviceptr base = 0;
  size_t size = 0;
  if (range_fn(&base, &size, reinterpret_cast<CUdeviceptr>(ptr)) != CUDA_SUCCESS) return FA_ERR_BAD_ARGUMENT;
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base));  // the handle names the allocation
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return static_cast<int>(e);
  }
  std::memcpy(handle, &h, sizeof(h));
  *offset = static_cast<unsigned long long>(reinterpret_cast<CUdeviceptr>(ptr) - base);
  return FA_OK;
}

int fa_ipc_open(const unsigned char* handle, void** base) {
  if (handle == nullptr || base == nullptr) return FA_ERR_NULL_POINTER;
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  const cudaError_t e = cudaIpcOpenMemHandle(base, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return static_cast<int>(e);
  }
  return FA_OK;
}

int fa_ipc_close(void* base) {
  if (base == nullptr) return FA_ERR_NULL_POINTER;
  const cudaError_t e = cudaIpcCloseMemHandle(base);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return static_cast<int>(e);
  }
  return FA_OK;
}

int fa_enable_peer_access(int peer_device) {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    (void)cudaGetLastError();
    return FA_ERR_NO_DEVICE;
  }
  if (peer_device == dev) return FA_OK;
  int can = 0;
  cudaError_t e = cudaDeviceCanAccessPeer(&can, dev, peer_device);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return static_cast<int>(e);
  }
  if (!can) return FA_ERR_BAD_ARGUMENT;
  e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    (void)cudaGetLastError();
    return FA_OK;
  }
  return e == cudaSuccess ? FA_OK : static_cast<int>(e);
}

int fa_ring_forward_pull(const void* Q, const void* K, const void* V, void* O, int B, int H, int n, int D, float scale,
                         int rank, int world, const void* const* peer_k, const void* const* peer_v, void* workspace,
                         size_t workspace_bytes, void* stream) {
  if (B <= 0 || H <= 0 || n <= 0) return FA_ERR_BAD_SHAPE;
  if (D != 64 && D != 128) return FA_ERR_BAD_HEAD_DIM;
  if (world <= 0 || world > kMaxRingWorld || rank < 0 || rank >= world) return FA_ERR_BAD_ARGUMENT;
  if (Q == nullptr || K == nullptr || V == nullptr || O == nullptr || workspace == nullptr) return FA_ERR_NULL_POINTER;
  if (world > 1 && (peer_k == nullptr || peer_v == nullptr)) return FA_ERR_NULL_POINTER;
  for (int r = 0; r < world && world > 1; ++r)
    if (r != rank && (peer_k[r] == nullptr || peer_v[r] == nullptr)) return FA_ERR_NULL_POINTER;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return FA_ERR_MISALIGNED;
  const RingLayout L = ring_layout(B, H, n, D, world, workspace_bytes);
  if (workspace_bytes < L.kv || (world > 1 && L.depth < 1)) return FA_ERR_BAD_ARGUMENT;
  DeviceState* st = nullptr;
  int rc = current_device_state(&st);
  if (rc != FA_OK) return rc;
  char* const ws = static_cast<char*>(workspace);
  float* const o_acc = reinterpret_cast<float*>(ws + L.state_o);
  float* const m = reinterpret_cast<float*>(ws + L.state_m);
  float* const l = reinterpret_cast<float*>(ws + L.state_l);
  if (world == 1) return fa_forward_partial(Q, K, V, o_acc, m, l, B, H, n, n, D, scale, 1, 1, O, stream);
  const size_t bytes = static_cast<size_t>(B) * H * n * D * 2;  // a multiple of 16: D is 64 or 128
  const int depth = L.depth < 2 ? L.depth : 2;  // the step after next reuses the buffer this step reads
  if (world > 2 && depth < 2) return FA_ERR_BAD_ARGUMENT;
  auto kbuf = [&](int slot) { return ws + L.kv + static_cast<size_t>(slot) * 2 * L.shard_bytes; };
  // Step x attends to the shard of rank - x (step 0: the local one) and, with the two spare warps of every CTA,
  // pulls the shard of rank - (x + 1) out of its owner's memory into the other buffer.  One launch per step, all
  // on the caller's stream; nothing else is enqueued.
  for (int x = 0; x < world; ++x) {
    FaCall c{};
    c.Q = Q;
    c.K = x == 0 ? K : kbuf((x - 1) % depth);
    c.V = x == 0 ? V : kbuf((x - 1) % depth) + L.shard_bytes;
    c.O16 = x == world - 1 ? O : nullptr;
    c.O32 = o_acc;
    c.m = m;
    c.l = l;
    c.B = B; c.H = H; c.Nq = n; c.Nkv = n; c.D = D;
    c.scale = scale;
    c.mode = FA_MODE_STATE;
    c.is_first = x == 0 ? 1 : 0;
    c.is_last = x == world - 1 ? 1 : 0;
    if (x + 1 < world) {
      const int from = (rank - (x + 1) + world) % world;
      c.pf_src_k = peer_k[from];
      c.pf_src_v = peer_v[from];
      c.pf_dst_k = kbuf(x % depth);
      c.pf_dst_v = kbuf(x % depth) + L.shard_bytes;
      c.pf_bytes = bytes;
    }
    if ((rc = run_fa(c, stream)) != FA_OK) return rc;
  }
  return FA_OK;
}

int fa_forward_cross(const void* Q, const void* K, const void* V, float* O, int B, int H, int Nq, int Nkv, int D,
                     float scale, void* stream) {
  if (D == 16) {
    if (B <= 0 || H <= 0 || Nq <= 0 || Nkv <= 0) return FA_ERR_BAD_SHAPE;
    if (!(scale == scale) || std::isinf(scale)) return FA_ERR_BAD_ARGUMENT;
    if (Q == nullptr || K == nullptr || V == nullptr || O == nullptr) return FA_ERR_NULL_POINTER;
    if (!aligned16(Q) || !aligned16(K) || !aligned16(V) || !aligned16(O)) return FA_ERR_MISALIGNED;
    DeviceState* st = nullptr;
    int rc = current_device_state(&st);
    if (rc != FA_OK) return rc;
    const long long rows = static_cast<long long>(B) * H * Nq;
    if (Nq >= 4) {  // four query rows per warp, K / V staged in shared memory
      const int q_blocks = (Nq + 15) / 16;
      const long long ctas = static_cast<long long>(B) * H * q_blocks;
      if (ctas > 0x7fffffffll) return FA_ERR_BAD_SHAPE;
      fa_fwd_d16_rows4_kernel<<<static_cast<unsigned>(ctas), 128, 0, static_cast<cudaStream_t>(stream)>>>(
          static_cast<const __half*>(Q), static_cast<const __half*>(K), static_cast<const __half*>(V), O, q_blocks, Nq, Nkv,
          scale * 1.4426950408889634f);
    } else {  // decode shapes (v7_5: one query per batch element): one warp per query row
      const long long ctas = (rows + 7) / 8;
      if (ctas > 0x7fffffffll) return FA_ERR_BAD_SHAPE;
      fa_fwd_d16_kernel<<<static_cast<unsigned>(ctas), 256, 0, static_cast<cudaStream_t>(stream)>>>(
          static_cast<const __half*>(Q), static_cast<const __half*>(K), static_cast<const __half*>(V), O, rows, Nq, Nkv,
          scale * 1.4426950408889634f);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? FA_OK : static_cast<int>(e);
  }
  FaCall c{Q, K, V, nullptr, O, nullptr, nullptr, B, H, Nq, Nkv, D, scale, FA_MODE_F32, 1, 1};
  return run_fa(c, stream);
}

int gemm_fused_epilogue(const void* A, const void* Bm, const float* bias, const float* residual, float* C, int M,
                        int N, int K, void* stream) {
  return run_gemm(A, Bm, bias, residual, C, M, N, K, stream);
}

int gemm_batched(const void* A, const void* Bm, float* C, int batch, int M, int N, int K, int trans_b, float alpha,
                 void* stream) {
  const GemmCall c{A, Bm, nullptr, nullptr, C, batch, M, N, K, trans_b != 0, GEMM_ACT_IDENTITY, alpha};
  return run_gemm(c, stream);
}

int gemm_fused_epilogue_act(const void* A, const void* Bm, const float* bias, const float* residual, float* C, int M,
                            int N, int K, int act, void* stream) {
  const GemmCall c{A, Bm, bias, residual, C, 1, M, N, K, false, act, 1.0f};
  return run_gemm(c, stream);
}

int gemm_fused_epilogue_f32(const float* A, const float* Bm, const float* bias, const float* residual, float* C, int M,
                            int N, int K, int act, void* stream) {
  return run_gemm_f32(A, Bm, bias, residual, C, M, N, K, act, stream);
}

int gemm_batched_strided(const void* A, const void* Bm, float* C, int batch, int M, int N, int K, long long lda,
                         long long ldb, long long ldc, long long stride_a, long long stride_b, long long stride_c,
                         int trans_b, float alpha, void* stream) {
  GemmCall c{A, Bm, nullptr, nullptr, C, batch, M, N, K, trans_b != 0, GEMM_ACT_IDENTITY, alpha};
  c.lda = lda; c.ldb = ldb; c.ldc = ldc;
  c.sa = stride_a; c.sb = stride_b; c.sc = stride_c;
  return run_gemm(c, stream);
}

int softmax_rows(const float* X, float* P, int rows, int N, void* stream) { return run_softmax(X, P, rows, N, stream); }

int softmax_rows_f16out(const float* X, void* P, int rows, int N, void* stream) {
  return run_softmax(X, static_cast<__half*>(P), rows, N, stream);
}

int fa_forward_unfused(const void* Q, const void* K, const void* V, float* O, float* S, void* P, int B, int H, int N,
                       int D, float scale, void* stream) {
  if (B <= 0 || H <= 0 || N <= 0 || D <= 0) return FA_ERR_BAD_SHAPE;
  if (Q == nullptr || K == nullptr || V == nullptr || O == nullptr || S == nullptr || P == nullptr)
    return FA_ERR_NULL_POINTER;
  const long long bh = static_cast<long long>(B) * H;
  if (bh > 0x7fffffffll || bh * N > 0x7fffffffll) return FA_ERR_BAD_SHAPE;
  // stage A: S = scale * Q K^T  (K as [N, D] rows = the transposed-B form)
  const GemmCall qk{Q, K, nullptr, nullptr, S, static_cast<int>(bh), N, N, D, true, GEMM_ACT_IDENTITY, scale};
  int rc = run_gemm(qk, stream);
  if (rc != FA_OK) return rc;
  // stage B: P = softmax_rows(S), rounded to fp16 for the tensor core
  if ((rc = run_softmax(S, static_cast<__half*>(P), static_cast<int>(bh * N), N, stream)) != FA_OK) return rc;
  // stage C: O = P V
  const GemmCall pv{P, V, nullptr, nullptr, O, static_cast<int>(bh), N, D, N, false, GEMM_ACT_IDENTITY, 1.0f};
  return run_gemm(pv, stream);
}

int fa_forward_host(const void* Q, const void* K, const void* V, void* O, int B, int H, int N, int D, float scale) {
  if (B <= 0 || H <= 0 || N <= 0) return FA_ERR_BAD_SHAPE;
  if (D != 64 && D != 128) return FA_ERR_BAD_HEAD_DIM;
  if (Q == nullptr || K == nullptr || V == nullptr || O == nullptr) return FA_ERR_NULL_POINTER;
  const long long bh = static_cast<long long>(B) * H;
  if (bh > 0x7fffffffll) return FA_ERR_BAD_SHAPE;
  DeviceState* st = nullptr;
  int rc = current_device_state(&st);
  if (rc != FA_OK) return rc;
  std::lock_guard<std::mutex> host_lock(st->host_mu);
  if ((rc = host_pipe_init(st)) != FA_OK) return rc;
  const size_t unit = static_cast<size_t>(N) * D * 2;  // bytes of one (b,h) of one tensor
  const size_t bytes = static_cast<size_t>(bh) * unit;
  void *dq, *dk, *dv, *dout;
  if ((rc = scratch_get(st, 0, bytes, &dq)) != FA_OK) return rc;
  if ((rc = scratch_get(st, 1, bytes, &dk)) != FA_OK) return rc;
  if ((rc = scratch_get(st, 2, bytes, &dv)) != FA_OK) return rc;
  if ((rc = scratch_get(st, 3, bytes, &dout)) != FA_OK) return rc;
  // (b,h) units are independent: while chunk c runs, chunk c+1 is on its way in and chunk c-1 on its way out
  // (PCIe is full duplex), so the call costs about the larger of the two copy directions, not their sum + kernel.
  const int nchunk = host_chunks(bytes, bh);
  for (int c = 0; c < nchunk; ++c) {
    const long long u0 = bh * c / nchunk, u1 = bh * (c + 1) / nchunk;
    const size_t off = static_cast<size_t>(u0) * unit, len = static_cast<size_t>(u1 - u0) * unit;
    FA_PIPE_TRY(st, cudaMemcpyAsync(static_cast<char*>(dq) + off, static_cast<const char*>(Q) + off, len, cudaMemcpyHostToDevice, st->s_in));
    FA_PIPE_TRY(st, cudaMemcpyAsync(static_cast<char*>(dk) + off, static_cast<const char*>(K) + off, len, cudaMemcpyHostToDevice, st->s_in));
    FA_PIPE_TRY(st, cudaMemcpyAsync(static_cast<char*>(dv) + off, static_cast<const char*>(V) + off, len, cudaMemcpyHostToDevice, st->s_in));
    FA_PIPE_TRY(st, cudaEventRecord(st->ev_in[c], st->s_in));
    FA_PIPE_TRY(st, cudaStreamWaitEvent(st->s_run, st->ev_in[c], 0));
    if ((rc = fa_forward(static_cast<char*>(dq) + off, static_cast<char*>(dk) + off, static_cast<char*>(dv) + off,
                         static_cast<char*>(dout) + off, 1, static_cast<int>(u1 - u0), N, D, scale, st->s_run)) != FA_OK) {
      host_pipe_drain(st);
      return rc;
    }
    FA_PIPE_TRY(st, cudaEventRecord(st->ev_run[c], st->s_run));
    FA_PIPE_TRY(st, cudaStreamWaitEvent(st->s_out, st->ev_run[c], 0));
    FA_PIPE_TRY(st, cudaMemcpyAsync(static_cast<char*>(O) + off, static_cast<char*>(dout) + off, len, cudaMemcpyDeviceToHost, st->s_out));
  }
  return host_pipe_drain(st);
}

int gemm_fused_epilogue_host(const void* A, const void* Bm, const float* bias, const float* residual, float* C, int M,
                             int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return FA_ERR_BAD_SHAPE;
  if (A == nullptr || Bm == nullptr || C == nullptr) return FA_ERR_NULL_POINTER;
  DeviceState* st = nullptr;
  int rc = current_device_state(&st);
  if (rc != FA_OK) return rc;
  std::lock_guard<std::mutex> host_lock(st->host_mu);
  if ((rc = host_pipe_init(st)) != FA_OK) return rc;
  const size_t ab = static_cast<size_t>(M) * K * 2, bb = static_cast<size_t>(K) * N * 2;
  const size_t cb = static_cast<size_t>(M) * N * 4, biasb = static_cast<size_t>(N) * 4;
  void *da, *db, *dbias = nullptr, *dres = nullptr, *dc;
  if ((rc = scratch_get(st, 0, ab, &da)) != FA_OK) return rc;
  if ((rc = scratch_get(st, 1, bb, &db)) != FA_OK) return rc;
  if ((rc = scratch_get(st, 3, cb, &dc)) != FA_OK) return rc;
  if (bias != nullptr && (rc = scratch_get(st, 4, biasb, &dbias)) != FA_OK) return rc;
  if (residual != nullptr && (rc = scratch_get(st, 2, cb, &dres)) != FA_OK) return rc;
  // B and bias are shared by every row block and go first; then row blocks of A / residual / C are pipelined
  FA_PIPE_TRY(st, cudaMemcpyAsync(db, Bm, bb, cudaMemcpyHostToDevice, st->s_in));
  if (bias != nullptr) FA_PIPE_TRY(st, cudaMemcpyAsync(dbias, bias, biasb, cudaMemcpyHostToDevice, st->s_in));
  const long long row_blocks = (static_cast<long long>(M) + 255) / 256;  // chunks are whole 256-row output tiles
  const int nchunk = host_chunks(ab + 2 * cb, row_blocks);
  for (int c = 0; c < nchunk; ++c) {
    const long long r0 = row_blocks * c / nchunk * 256;
    long long r1 = row_blocks * (c + 1) / nchunk * 256;
    if (r1 > M) r1 = M;
    const int rows = static_cast<int>(r1 - r0);
    const size_t a_off = static_cast<size_t>(r0) * K * 2, c_off = static_cast<size_t>(r0) * N * 4;
    FA_PIPE_TRY(st, cudaMemcpyAsync(static_cast<char*>(da) + a_off, static_cast<const char*>(A) + a_off,
                                static_cast<size_t>(rows) * K * 2, cudaMemcpyHostToDevice, st->s_in));
    if (residual != nullptr)
      FA_PIPE_TRY(st, cudaMemcpyAsync(static_cast<char*>(dres) + c_off, reinterpret_cast<const char*>(residual) + c_off,
                                  static_cast<size_t>(rows) * N * 4, cudaMemcpyHostToDevice, st->s_in));
    FA_PIPE_TRY(st, cudaEventRecord(st->ev_in[c], st->s_in));
    FA_PIPE_TRY(st, cudaStreamWaitEvent(st->s_run, st->ev_in[c], 0));
    rc = gemm_fused_epilogue(static_cast<char*>(da) + a_off, db, static_cast<const float*>(dbias),
                             residual != nullptr ? reinterpret_cast<const float*>(static_cast<char*>(dres) + c_off) : nullptr,
                             reinterpret_cast<float*>(static_cast<char*>(dc) + c_off), rows, N, K, st->s_run);
    if (rc != FA_OK) {
      host_pipe_drain(st);
      return rc;
    }
    FA_PIPE_TRY(st, cudaEventRecord(st->ev_run[c], st->s_run));
    FA_PIPE_TRY(st, cudaStreamWaitEvent(st->s_out, st->ev_run[c], 0));
    FA_PIPE_TRY(st, cudaMemcpyAsync(reinterpret_cast<char*>(C) + c_off, static_cast<char*>(dc) + c_off,
                                static_cast<size_t>(rows) * N * 4, cudaMemcpyDeviceToHost, st->s_out));
  }
  return host_pipe_drain(st);
}

int softmax_rows_host(const float* X, float* P, int rows, int N) {
  if (rows <= 0 || N <= 0) return FA_ERR_BAD_SHAPE;
  if (X == nullptr || P == nullptr) return FA_ERR_NULL_POINTER;
  DeviceState* st = nullptr;
  int rc = current_device_state(&st);
  if (rc != FA_OK) return rc;
  std::lock_guard<std::mutex> host_lock(st->host_mu);
  if ((rc = host_pipe_init(st)) != FA_OK) return rc;
  const size_t pitch = static_cast<size_t>(N) * 4;
  const size_t bytes = static_cast<size_t>(rows) * pitch;
  void* dx;
  if ((rc = scratch_get(st, 0, bytes, &dx)) != FA_OK) return rc;
  const int nchunk = host_chunks(2 * bytes, rows);
  for (int c = 0; c < nchunk; ++c) {
    const long long r0 = static_cast<long long>(rows) * c / nchunk, r1 = static_cast<long long>(rows) * (c + 1) / nchunk;
    const size_t off = static_cast<size_t>(r0) * pitch, len = static_cast<size_t>(r1 - r0) * pitch;
    float* const dchunk = reinterpret_cast<float*>(static_cast<char*>(dx) + off);
    FA_PIPE_TRY(st, cudaMemcpyAsync(dchunk, reinterpret_cast<const char*>(X) + off, len, cudaMemcpyHostToDevice, st->s_in));
    FA_PIPE_TRY(st, cudaEventRecord(st->ev_in[c], st->s_in));
    FA_PIPE_TRY(st, cudaStreamWaitEvent(st->s_run, st->ev_in[c], 0));
    if ((rc = softmax_rows(dchunk, dchunk, static_cast<int>(r1 - r0), N, st->s_run)) != FA_OK) {
      host_pipe_drain(st);
      return rc;
    }
    FA_PIPE_TRY(st, cudaEventRecord(st->ev_run[c], st->s_run));
    FA_PIPE_TRY(st, cudaStreamWaitEvent(st->s_out, st->ev_run[c], 0));
    FA_PIPE_TRY(st, cudaMemcpyAsync(reinterpret_cast<char*>(P) + off, dchunk, len, cudaMemcpyDeviceToHost, st->s_out));
  }
  return host_pipe_drain(st);
}

int fa_host_release(void) {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
    (void)cudaGetLastError();
    return FA_ERR_NO_DEVICE;
  }
  DeviceState& st = g_dev[dev];
  std::lock_guard<std::mutex> host_lock(st.host_mu);
  for (int i = 0; i < 6; ++i) {
    if (st.scratch[i] != nullptr) cudaFree(st.scratch[i]);
    st.scratch[i] = nullptr;
    st.scratch_bytes[i] = 0;
  }
  return FA_OK;
}

int fa_last_debug(unsigned int* words, int n) {
  if (words == nullptr || n <= 0 || g_dbg_host == nullptr) return 0;
  const int k = n < 8 ? n : 8;
  for (int i = 0; i < k; ++i) words[i] = reinterpret_cast<volatile unsigned int*>(g_dbg_host)[i];
  return k;
}

unsigned long long fa_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int fa_plan_segments(int B, int H, int Nq, int Nkv, int sms, int* out, int max_segments) {
  if (B <= 0 || H <= 0 || Nq <= 0 || Nkv <= 0 || sms <= 0) return FA_ERR_BAD_SHAPE;
  FaParams p{};
  p.nq = Nq;
  p.nkv = Nkv;
  p.q_blocks = (Nq + 255) / 256;
  const long long items = static_cast<long long>(B) * H * p.q_blocks;
  if (items > 0x7fffffffll) return FA_ERR_BAD_SHAPE;
  const int nkv_tiles = (Nkv + 127) / 128;
  const FaPlan plan = fa_plan(items, nkv_tiles, Nq, sms, true);
  p.items_full = plan.items_full;
  p.tail_items = plan.tail_items;
  p.tail_ctas = plan.tail_ctas;
  p.tail_rows = plan.tail_rows;
  int n = 0;
  for (int cta = 0; cta < plan.items_full + plan.tail_ctas; ++cta) {
    FaSegment seg;
    for (int si = 0; fa_segment(p, nkv_tiles, cta, si, seg); ++si, ++n) {
      if (out != nullptr && n < max_segments) {
        int* row = out + static_cast<size_t>(n) * 6;
        row[0] = cta; row[1] = seg.item; row[2] = seg.kv0; row[3] = seg.kv1; row[4] = seg.half >= 0 ? -2 - seg.half : seg.slot; row[5] = seg.tail_item;
      }
    }
  }
  return n;
}

int fa_set_tail_split(int mode) {
  const int before = g_tail_split.exchange(mode < 0 ? -1 : (mode > 3 ? 3 : mode), std::memory_order_relaxed);
  return before;
}

int fa_map_cache_stats(unsigned long long* hits, unsigned long long* misses) {
  if (hits != nullptr) *hits = g_map_hits.load(std::memory_order_relaxed);
  if (misses != nullptr) *misses = g_map_misses.load(std::memory_order_relaxed);
  return kMapCacheEntries;
}

#ifdef FA_TRACE
void fa_trace_config(long long* device_buf, int cta) {
  g_trace_buf = device_buf;
  g_trace_cta = cta;
}
#endif

}  // extern "C"
