// Include-wrapper around the UNMODIFIED reference file
//   flashattn_warp_spc/flashattn_streaming_16x32_2warp_softmax_v12i.cu
// (compiled where it lies under $(REF); nothing is copied).  Exposes its CPU oracle
// cpu_reference (:243-286), rel_l2_error (:288-296), its input generator (:30-33,317-320) and its
// kernel flashattn_2warp_nosmemW_online_softmax<16,16,4,32> (:60-240) behind extern "C".
// Test infrastructure only (oracle/_ref/libref_v12i.so).
#define main ref_v12i_unused_main
#include "flashattn_warp_spc/flashattn_streaming_16x32_2warp_softmax_v12i.cu"
#undef main

extern "C" {

void ref_v12i_cpu_reference(const unsigned short* q, const unsigned short* k, const unsigned short* v, float* o,
                            int B, int M_TOTAL, int SEQ_LEN, int KDIM, int DV, float scale) {
  std::vector<half> hQ((size_t)B * M_TOTAL * KDIM), hK((size_t)B * SEQ_LEN * KDIM), hV((size_t)B * SEQ_LEN * DV);
  memcpy(hQ.data(), q, hQ.size() * 2);
  memcpy(hK.data(), k, hK.size() * 2);
  memcpy(hV.data(), v, hV.size() * 2);
  std::vector<float> hO((size_t)B * M_TOTAL * DV, 0.0f);
  cpu_reference(hQ, hK, hV, hO, B, M_TOTAL, SEQ_LEN, KDIM, DV, scale);
  memcpy(o, hO.data(), hO.size() * 4);
}

double ref_v12i_rel_l2(const float* a, const float* b, size_t n) {
  std::vector<float> va(a, a + n), vb(b, b + n);
  return rel_l2_error(va, vb);
}

// The reference's fixture: mt19937(seed), uniform(-1,1) rounded to half, fill order Q, K, V (:317-320).
void ref_v12i_fill_uniform(unsigned seed, unsigned short* q, size_t nq, unsigned short* k, size_t nk,
                           unsigned short* v, size_t nv) {
  std::mt19937 gen(seed);
  auto fill = [&](unsigned short* dst, size_t n) {
    for (size_t i = 0; i < n; ++i) {
      half h = __float2half(frand_uniform(gen));
      memcpy(dst + i, &h, 2);
    }
  };
  fill(q, nq);
  fill(k, nk);
  fill(v, nv);
}

// Native instantiation (KDIM = DV = 16, 4 rows per block, 32-key tiles), device pointers.
int ref_v12i_launch(const void* dQ, const void* dK, const void* dV, float* dO, int B, int M_TOTAL, int SEQ_LEN,
                    float scale, void* stream) {
  constexpr int KDIM = 16, DV = 16, ROWS = 4, N_TILE = 32;
  dim3 grid(B, (M_TOTAL + ROWS - 1) / ROWS, 1), block(64, 1, 1);
  size_t shmem = (size_t)N_TILE * KDIM * sizeof(half) + (size_t)N_TILE * DV * sizeof(half) + 3 * ROWS * sizeof(float);
  flashattn_2warp_nosmemW_online_softmax<KDIM, DV, ROWS, N_TILE><<<grid, block, shmem, (cudaStream_t)stream>>>(
      (const half*)dQ, (const half*)dK, (const half*)dV, dO, SEQ_LEN, M_TOTAL, scale);
  return (int)cudaGetLastError();
}

}  // extern "C"
