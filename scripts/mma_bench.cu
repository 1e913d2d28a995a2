// Developer microbenchmark: cycles per tcgen05.mma for the operand flavours the attention and GEMM
// kernels use (operands resident in smem / TMEM, no TMA, no softmax).  One CTA per SM.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o mma_bench scripts/mma_bench.cu
#include <cstdio>
#include "../flashattention_kernel_project_b200/csrc/sm100_ptx.cuh"
using namespace fa;

// mode 0: SS K-major x K-major N=128, one accumulator   (QK^T of one tile)
// mode 1: SS K-major x K-major N=128, two accumulators alternating every 8 MMAs
// mode 2: TS (A in TMEM) x MN-major B, N=128            (PV)
// mode 3: SS K-major x MN-major N=256                   (GEMM tile)
// mode 4: attention mix: 8 QK->S0, 8 PV->O0, 8 QK->S1, 8 PV->O1
// mode 5: SS K-major x K-major N=256
// mode 6: TS x MN-major B, N=128, two accumulators alternating
template <int mode>
__global__ void __launch_bounds__(384, 1) mma_bench(int iters, long long* out, int traffic, int pace) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[8];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) {
    uint32_t v = 0x3c003c00u;  // 1.0h pairs
    if (pace < 0) {            // pseudo-random fp16 in (-1, 1): realistic operand toggling
      uint32_t h = (i + 1) * 2654435761u + blockIdx.x * 40503u;
      h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      const float a = ((h & 0xffff) / 32768.0f) - 1.0f, b = ((h >> 16) / 32768.0f) - 1.0f;
      v = pack_half2(a, b);
    }
    reinterpret_cast<uint32_t*>(smem)[i] = v;
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) mbar_init(&bar2[i], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tmem_slot;
  const uint32_t sa = smem_u32(smem), sb = sa + 64 * 1024;
  long long t0 = 0, t1 = 0;
  __shared__ volatile int done_flag;
  if (threadIdx.x == 0) done_flag = 0;
  __syncthreads();
  if (warp >= 4 && traffic != 0) {
    // softmax-like side traffic until the MMA warp finishes
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t col = (warp >= 8) ? 128u : 0u;
    uint32_t r[32];
    for (int k = 0; k < 32; ++k) r[k] = k;
    while (done_flag == 0) {
      if (traffic == 1) {
        tmem_ld_x32(tb + lane_base + col + 0, r); tmem_ld_x32(tb + lane_base + col + 32, r);
        tmem_ld_x32(tb + lane_base + col + 64, r); tmem_ld_x32(tb + lane_base + col + 96, r);
        tmem_wait_ld();
        tmem_st_x16(tb + lane_base + col + 64, r); tmem_st_x16(tb + lane_base + col + 80, r);
        tmem_st_x16(tb + lane_base + col + 96, r); tmem_st_x16(tb + lane_base + col + 112, r);
        tmem_wait_st();
      } else {
        uint4 v = make_uint4(r[0], r[1], r[2], r[3]);
        for (int k = 0; k < 16; ++k)
          *reinterpret_cast<uint4*>(smem + 128 * 1024 + ((threadIdx.x * 16 + k * 4096) & 32767)) = v;
      }
      for (int w = 0; w < pace; ++w) __nanosleep(20);
    }
    if (r[5] == 0xdeadbeef) out[0] = r[7];
  }
  if (warp == 1) {
    const uint32_t id_qk = umma_idesc_f16(128, 128, 0, 0), id_pv = umma_idesc_f16(128, 128, 0, 1);
    const uint32_t id_g = umma_idesc_f16(128, 256, 0, 1), id_k256 = umma_idesc_f16(128, 256, 0, 0);
    const uint32_t id_n64 = umma_idesc_f16(128, 64, 0, 0), id_pv64 = umma_idesc_f16(128, 64, 0, 1);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t off = (ks >> 2) * 16384 + (ks & 3) * 32;
            const uint32_t acc = (it | g | ks) ? 1u : 0u;
            if constexpr (mode == 0) umma_ss(tb, umma_desc_sw128(sa + off, 16, 1024), umma_desc_sw128(sb + off, 16, 1024), id_qk, acc);
            else if constexpr (mode == 1) umma_ss(tb + (g & 1) * 128, umma_desc_sw128(sa + off, 16, 1024), umma_desc_sw128(sb + off, 16, 1024), id_qk, acc);
            else if constexpr (mode == 2) umma_ts(tb + 256, tb + ks * 8, umma_desc_sw128(sb + ks * 2048, 16384, 1024), id_pv, acc);
            else if constexpr (mode == 3) umma_ss(tb, umma_desc_sw128(sa + (ks & 3) * 32, 16, 1024), umma_desc_sw128(sb + (ks & 3) * 2048, 8192, 1024), id_g, acc);
            else if constexpr (mode == 5) umma_ss(tb, umma_desc_sw128(sa + off, 16, 1024), umma_desc_sw128(sb + off, 16, 1024), id_k256, acc);
            else if constexpr (mode == 6) umma_ts(tb + 256 + (g & 1) * 128, tb + (g & 1) * 128 + ks * 8, umma_desc_sw128(sb + ks * 2048, 16384, 1024), id_pv, acc);
            else if constexpr (mode == 7) umma_ss(tb + (ks & 1) * 128, umma_desc_sw128(sa + off, 16, 1024), umma_desc_sw128(sb + off, 16, 1024), id_qk, (it | g | (ks >> 1)) ? 1u : 0u);
            else if constexpr (mode == 8) umma_ss(tb + (ks & 3) * 64, umma_desc_sw128(sa + off, 16, 1024), umma_desc_sw128(sb + off + (ks & 1) * 8192, 16, 1024), id_n64, (it | g | (ks >> 2)) ? 1u : 0u);
            else if constexpr (mode == 9) umma_ss(tb, umma_desc_sw128(sa + off, 16, 1024), umma_desc_sw128(sb + off, 16, 1024), id_n64, acc);
            else if constexpr (mode == 10) umma_ts(tb + 256 + (ks & 1) * 64, tb + (ks >> 1) * 8, umma_desc_sw128(sb + (ks & 1) * 16384 + (ks >> 1) * 2048, 16384, 1024), id_pv64, (it | g | (ks >> 1)) ? 1u : 0u);
            else if constexpr (mode == 11) umma_ss(tb + (ks & 3) * 128, umma_desc_sw128(sa + off, 16, 1024), umma_desc_sw128(sb + off, 16, 1024), id_qk, (it | g | (ks >> 2)) ? 1u : 0u);
            else {
              (void)id_n64; (void)id_pv64; (void)id_g; (void)id_k256;
              if ((g & 1) == 0) umma_ss(tb + (g >> 1) * 128, umma_desc_sw128(sa + (g >> 1) * 32768 + off, 16, 1024), umma_desc_sw128(sb + off, 16, 1024), id_qk, ks ? 1u : 0u);
              else umma_ts(tb + 256 + (g >> 1) * 128, tb + (g >> 1) * 128 + ks * 8, umma_desc_sw128(sb + 32768 + ks * 2048, 16384, 1024), id_pv, acc);
            }
          }
          if constexpr (mode >= 12) umma_commit(&bar2[g]);
          if constexpr (mode == 14) umma_commit(&bar2[4 + g]);
        }
      }
      __syncwarp();
      if constexpr (mode == 13 || mode == 15) { tc_fence_after(); }
      if constexpr (mode == 15) { (void)mbar_try_wait(&bar2[0], 1); (void)mbar_try_wait(&bar2[1], 1); (void)mbar_try_wait(&bar2[2], 1); (void)mbar_try_wait(&bar2[3], 1); }
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    DebugChannel d{nullptr};
    mbar_wait(&bar, 0, d, 0);
    t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
    done_flag = 1;
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 0) tmem_dealloc(tb, 512);
}

template <int MODE>
float run_mode(int grid, int iters, long long* d_out, int smem_bytes, int traffic = 0, int pace = 0) {
  cudaFuncSetAttribute(mma_bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  mma_bench<MODE><<<grid, 384, smem_bytes>>>(iters, d_out, traffic, pace);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  mma_bench<MODE><<<grid, 384, smem_bytes>>>(iters, d_out, traffic, pace);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("mode %d failed: %s\n", MODE, cudaGetErrorString(err)); exit(1); }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 148 * sizeof(long long));
  const int smem_bytes = 160 * 1024 + 1024;
  const char* names[] = {"SS Kmaj x Kmaj N=128, 1 accumulator (QK)", "SS Kmaj x Kmaj N=128, 2 accumulators", "TS x MNmaj N=128 (PV)",
                         "SS Kmaj x MNmaj N=256 (GEMM tile)", "attention mix QK0 PV0 QK1 PV1", "SS Kmaj x Kmaj N=256", "TS x MNmaj N=128, 2 accumulators",
                         "SS N=128 alternate 2 accumulators per MMA", "SS N=64 round-robin 4 accumulators", "SS N=64 one accumulator",
                         "TS N=64 alternate 2 accumulators per MMA", "SS N=128 round-robin 4 accumulators",
                         "attention mix + commit per 8 MMAs", "attention mix + commit per 8 + fence per 32", "attention mix + 2 commits per 8 MMAs",
                         "attention mix + commit per 8 + fence + 4 try_wait per 32"};
  for (int grid : {148}) {
    for (int mode = 0; mode < 16; ++mode) {
      const int iters = 200;
      float ms = 0;
      switch (mode) {
        case 0: ms = run_mode<0>(grid, iters, d_out, smem_bytes); break;
        case 1: ms = run_mode<1>(grid, iters, d_out, smem_bytes); break;
        case 2: ms = run_mode<2>(grid, iters, d_out, smem_bytes); break;
        case 3: ms = run_mode<3>(grid, iters, d_out, smem_bytes); break;
        case 4: ms = run_mode<4>(grid, iters, d_out, smem_bytes); break;
        case 5: ms = run_mode<5>(grid, iters, d_out, smem_bytes); break;
        case 6: ms = run_mode<6>(grid, iters, d_out, smem_bytes); break;
        case 7: ms = run_mode<7>(grid, iters, d_out, smem_bytes); break;
        case 8: ms = run_mode<8>(grid, iters, d_out, smem_bytes); break;
        case 9: ms = run_mode<9>(grid, iters, d_out, smem_bytes); break;
        case 10: ms = run_mode<10>(grid, iters, d_out, smem_bytes); break;
        case 11: ms = run_mode<11>(grid, iters, d_out, smem_bytes); break;
        case 12: ms = run_mode<12>(grid, iters, d_out, smem_bytes); break;
        case 13: ms = run_mode<13>(grid, iters, d_out, smem_bytes); break;
        case 14: ms = run_mode<14>(grid, iters, d_out, smem_bytes); break;
        default: ms = run_mode<15>(grid, iters, d_out, smem_bytes); break;
      }
      long long h[148]; cudaMemcpy(h, d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      double cyc = 0; for (int i = 0; i < grid; ++i) cyc += h[i]; cyc /= grid;
      const double nmma = iters * 32.0;
      const double flop_per = 2.0 * 128 * ((mode == 3 || mode == 5) ? 256 : (mode == 8 || mode == 9 || mode == 10) ? 64 : 128) * 16;
      printf("grid %3d mode %d %-45s: %7.1f cycles/MMA  %6.1f flop/clk/SM  (%.3f ms -> %.0f TFLOP/s chip)\n", grid, mode, names[mode],
             cyc / nmma, flop_per * nmma / cyc, ms, flop_per * nmma * grid / ms / 1e9);
    }
  }
  for (int rep = 0; rep < 3; ++rep) {
    float ms = run_mode<4>(148, 2000, d_out, smem_bytes, 0, rep == 1 ? -1 : 0);
    long long h[148]; cudaMemcpy(h, d_out, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
    double cyc = 0; for (int i = 0; i < 148; ++i) cyc += h[i]; cyc /= 148;
    printf("attention mix, 2000 iters, %s operands: %.1f cycles/MMA (%.3f ms -> %.0f TFLOP/s, %.0f MHz)\n", rep == 1 ? "RANDOM" : "constant",
           cyc / 64000.0, ms, 2.0 * 128 * 128 * 16 * 64000.0 * 148 / ms / 1e9, cyc / ms / 1e3);
  }
  for (int traffic : {1}) for (int pace : {0}) {
    float ms = run_mode<4>(148, 200, d_out, smem_bytes, traffic, pace);
    long long h[148]; cudaMemcpy(h, d_out, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
    double cyc = 0; for (int i = 0; i < 148; ++i) cyc += h[i]; cyc /= 148;
    printf("attention mix with %s side traffic (8 warps, pace %d): %.1f cycles/MMA (%.3f ms)\n", traffic == 1 ? "TMEM ld/st" : "smem store", pace, cyc / 6400.0, ms);
  }
  return 0;
}
