/*
 * fa_oracle.c -- CPU restatement of the reference's host-side oracles for the attention /
 * fused-GEMM / row-softmax path of jeehun98/FlashAttention_Kernel_Project.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it; the product path
 * (flashattention_kernel_project_b200/) never does and has no CPU fallback.
 *
 * Parity status: PINNED.  Every function here is checked bit-for-bit (same operation order, fp32
 * or double accumulation exactly as the reference writes it) against the reference's own
 * functions compiled unmodified from /root/reference into oracle/_ref/ (see oracle/Makefile and
 * tests/test_oracle.py), and against the committed fixtures in tests/golden/ that those compiled
 * functions produced (tests/golden/make_golden.py).  One exception, marked where it is defined:
 * fa_oracle_streaming_state_f16 restates a rule that exists in the reference as DEVICE code only (no
 * host function to compile); it is pinned through its invariants and against the reference's
 * streaming kernel run on the GPU.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (no -ffast-math: summation order and
 * rounding must match the reference's g++-compiled host code).
 *
 * Citations are relative to the reference repository root.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------------------------------------------------------------------------------------
 * IEEE binary16 <-> binary32 (the reference uses __half2float / __float2half from cuda_fp16.h,
 * which are exact / round-to-nearest-even on the host).
 * ------------------------------------------------------------------------------------- */
static inline float half_bits_to_float(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1fu;
  uint32_t man = h & 0x3ffu;
  uint32_t bits;
  if (exp == 0) {
    if (man == 0) {
      bits = sign;
    } else { /* subnormal: normalise */
      int e = -1;
      do {
        ++e;
        man <<= 1;
      } while ((man & 0x400u) == 0);
      man &= 0x3ffu;
      bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
    }
  } else if (exp == 31) {
    bits = sign | 0x7f800000u | (man << 13);
  } else {
    bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
  }
  float f;
  memcpy(&f, &bits, 4);
  return f;
}

static inline uint16_t float_to_half_bits(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  const uint32_t absx = x & 0x7fffffffu;
  if (absx >= 0x7f800000u) { /* inf / nan */
    return (uint16_t)(sign | 0x7c00u | ((absx > 0x7f800000u) ? 0x200u : 0u));
  }
  if (absx >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u); /* rounds to >= 65520 -> inf */
  if (absx < 0x33000001u) return (uint16_t)sign;             /* < 2^-25 (or == 2^-25 tie-to-even) -> 0 */
  int32_t e = (int32_t)(absx >> 23) - 127;
  uint32_t man = (absx & 0x7fffffu) | 0x800000u;
  uint32_t shift, half_man, rem, halfway;
  if (e < -14) { /* subnormal half */
    shift = (uint32_t)(13 + (-14 - e));
    half_man = man >> shift;
    rem = man & ((1u << shift) - 1u);
    halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (half_man & 1u))) ++half_man;
    return (uint16_t)(sign | half_man);
  }
  half_man = (man >> 13) & 0x3ffu;
  rem = man & 0x1fffu;
  uint32_t out = ((uint32_t)(e + 15) << 10) | half_man;
  if (rem > 0x1000u || (rem == 0x1000u && (out & 1u))) ++out; /* carry may bump the exponent: correct */
  return (uint16_t)(sign | out);
}

void fa_oracle_half_to_float(const uint16_t* in, float* out, size_t n) {
  for (size_t i = 0; i < n; ++i) out[i] = half_bits_to_float(in[i]);
}
void fa_oracle_float_to_half(const float* in, uint16_t* out, size_t n) {
  for (size_t i = 0; i < n; ++i) out[i] = float_to_half_bits(in[i]);
}

/* ---------------------------------------------------------------------------------------
 * Attention, fp16 inputs, fp32 math -- restates cpu_reference
 * (flashattn_warp_spc/flashattn_streaming_16x32_2warp_softmax_v12i.cu:243-286):
 *   scores[t] = (sum_k q*k, fp32, k ascending) * scale;  mmax = max;  denom = sum_t exp(s-mmax);
 *   out[d] = sum_t (exp(s-mmax)/denom) * v[t][d], t ascending.
 * The reference re-evaluates exp(scores[t]-mmax)/denom inside the d loop (:275-281); the value
 * does not depend on d, so it is computed once per t here -- identical results, D x fewer exps.
 * One row = one (b, m) pair.  `rows` (optional) selects a subset of flat row ids b*M + m, which
 * is how the BASELINE configs too large for a full CPU pass are sampled.
 * ------------------------------------------------------------------------------------- */
static void attention_row_f16(const uint16_t* Q, const uint16_t* K, const uint16_t* V, float* out_row, int b, int m,
                              int M_TOTAL, int SEQ_LEN, int KDIM, int DV, float scale, float* scores, float* qf) {
  const uint16_t* q = Q + ((size_t)b * M_TOTAL + m) * KDIM;
  for (int k = 0; k < KDIM; ++k) qf[k] = half_bits_to_float(q[k]);
  float mmax = -INFINITY;
  for (int t = 0; t < SEQ_LEN; ++t) {
    const uint16_t* kr = K + ((size_t)b * SEQ_LEN + t) * KDIM;
    float acc = 0.0f;
    for (int k = 0; k < KDIM; ++k) acc += qf[k] * half_bits_to_float(kr[k]);
    const float s = acc * scale;
    scores[t] = s;
    mmax = (mmax < s) ? s : mmax; /* std::max(mmax, s) */
  }
  float denom = 0.0f;
  for (int t = 0; t < SEQ_LEN; ++t) denom += expf(scores[t] - mmax);
  for (int t = 0; t < SEQ_LEN; ++t) scores[t] = expf(scores[t] - mmax) / denom; /* w, as at :277 */
  for (int d = 0; d < DV; ++d) {
    float out = 0.0f;
    for (int t = 0; t < SEQ_LEN; ++t) out += scores[t] * half_bits_to_float(V[((size_t)b * SEQ_LEN + t) * DV + d]);
    out_row[d] = out;
  }
}

/* Full tensor: O fp32 [B, M_TOTAL, DV]. nthreads <= 0: all cores. */
void fa_oracle_attention_f16(const uint16_t* Q, const uint16_t* K, const uint16_t* V, float* O, int B, int M_TOTAL,
                             int SEQ_LEN, int KDIM, int DV, float scale, int nthreads) {
  const long long total = (long long)B * M_TOTAL;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
#pragma omp parallel
  {
    float* scores = (float*)malloc(sizeof(float) * (size_t)SEQ_LEN);
    float* qf = (float*)malloc(sizeof(float) * (size_t)KDIM);
#pragma omp for schedule(dynamic, 16)
    for (long long r = 0; r < total; ++r) {
      const int b = (int)(r / M_TOTAL), m = (int)(r % M_TOTAL);
      attention_row_f16(Q, K, V, O + (size_t)r * DV, b, m, M_TOTAL, SEQ_LEN, KDIM, DV, scale, scores, qf);
    }
    free(scores);
    free(qf);
  }
}

/* Sampled rows: out fp32 [nrows, DV], rows[i] = b*M_TOTAL + m. */
void fa_oracle_attention_f16_rows(const uint16_t* Q, const uint16_t* K, const uint16_t* V, float* out,
                                  const long long* rows, int nrows, int B, int M_TOTAL, int SEQ_LEN, int KDIM, int DV,
                                  float scale, int nthreads) {
  (void)B;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
#pragma omp parallel
  {
    float* scores = (float*)malloc(sizeof(float) * (size_t)SEQ_LEN);
    float* qf = (float*)malloc(sizeof(float) * (size_t)KDIM);
#pragma omp for schedule(dynamic, 1)
    for (int i = 0; i < nrows; ++i) {
      const int b = (int)(rows[i] / M_TOTAL), m = (int)(rows[i] % M_TOTAL);
      attention_row_f16(Q, K, V, out + (size_t)i * DV, b, m, M_TOTAL, SEQ_LEN, KDIM, DV, scale, scores, qf);
    }
    free(scores);
    free(qf);
  }
}

/* ---------------------------------------------------------------------------------------
 * Attention, fp32 inputs, double accumulation, scale = 1/sqrt(D) -- restates flashattn_cpu_ref
 * (GEMM/FlashAttention Forward Fused/flashattn_forward_fused_5_4_2.cu:224-271).
 * ------------------------------------------------------------------------------------- */
void fa_oracle_attention_f32_dacc(const float* Q, const float* K, const float* V, float* O, int BH, int N, int D) {
  const float scale = 1.0f / sqrtf((float)D);
  float* scores = (float*)malloc(sizeof(float) * (size_t)N);
  double* w = (double*)malloc(sizeof(double) * (size_t)N);
  for (int bh = 0; bh < BH; ++bh) {
    const float* Qb = Q + (size_t)bh * N * D;
    const float* Kb = K + (size_t)bh * N * D;
    const float* Vb = V + (size_t)bh * N * D;
    float* Ob = O + (size_t)bh * N * D;
    for (int i = 0; i < N; ++i) {
      float max_s = -INFINITY;
      for (int j = 0; j < N; ++j) {
        double acc = 0.0;
        for (int k = 0; k < D; ++k) acc += (double)Qb[i * D + k] * Kb[j * D + k];
        acc *= scale;
        scores[j] = (float)acc;
        max_s = (max_s < scores[j]) ? scores[j] : max_s;
      }
      double sum_e = 0.0;
      for (int j = 0; j < N; ++j) sum_e += exp((double)(scores[j] - max_s));
      for (int j = 0; j < N; ++j) w[j] = exp((double)(scores[j] - max_s)) / sum_e; /* :262, hoisted out of k */
      for (int k = 0; k < D; ++k) {
        double out = 0.0;
        for (int j = 0; j < N; ++j) out += w[j] * Vb[j * D + k];
        Ob[i * D + k] = (float)out;
      }
    }
  }
  free(scores);
  free(w);
}

/* ---------------------------------------------------------------------------------------
 * Attention, fp32 inputs, fp32 math, explicit scale, -1e30 sentinel -- restates flashattn_ref
 * (Softmax micro kernel/flashattn_fused_softmax_block/flashattn_fused_softmax_block.cu:214-264).
 * ------------------------------------------------------------------------------------- */
void fa_oracle_attention_f32(const float* Q, const float* K, const float* V, float* O, int N, int D, int BH,
                             float scale) {
  float* scores = (float*)malloc(sizeof(float) * (size_t)N);
  for (int bh = 0; bh < BH; ++bh) {
    for (int i = 0; i < N; ++i) {
      for (int j = 0; j < N; ++j) {
        float sum = 0.0f;
        for (int d = 0; d < D; ++d) sum += Q[((size_t)bh * N + i) * D + d] * K[((size_t)bh * N + j) * D + d];
        scores[j] = sum * scale;
      }
      float max_x = -1e30f;
      for (int j = 0; j < N; ++j) max_x = (max_x < scores[j]) ? scores[j] : max_x;
      float sum_e = 0.0f;
      for (int j = 0; j < N; ++j) {
        scores[j] = expf(scores[j] - max_x);
        sum_e += scores[j];
      }
      for (int j = 0; j < N; ++j) scores[j] /= sum_e;
      for (int d = 0; d < D; ++d) {
        float acc = 0.0f;
        for (int j = 0; j < N; ++j) acc += scores[j] * V[((size_t)bh * N + j) * D + d];
        O[((size_t)bh * N + i) * D + d] = acc;
      }
    }
  }
  free(scores);
}

/* ---------------------------------------------------------------------------------------
 * Row softmax -- the softmax step of flashattn_ref (same file, :242-252): max with a -1e30
 * sentinel, e = expf(x - max), p = e / sum.  The three standalone softmax programs of the
 * reference print only and hold no CPU oracle.
 * ------------------------------------------------------------------------------------- */
void fa_oracle_softmax_rows(const float* X, float* P, long long rows, int N, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
#pragma omp parallel for schedule(static)
  for (long long r = 0; r < rows; ++r) {
    const float* x = X + (size_t)r * N;
    float* p = P + (size_t)r * N;
    float max_x = -1e30f;
    for (int j = 0; j < N; ++j) max_x = (max_x < x[j]) ? x[j] : max_x;
    float sum_e = 0.0f;
    for (int j = 0; j < N; ++j) {
      p[j] = expf(x[j] - max_x);
      sum_e += p[j];
    }
    for (int j = 0; j < N; ++j) p[j] /= sum_e;
  }
}

/* ---------------------------------------------------------------------------------------
 * C = ReLU(A*B + bias) + residual, fp32, k ascending -- restates cpu_gemm_fused_ref
 * (GEMM/tensorcore_gemm_fused_epilogue/tensorcore_gemm_fused_epilogue.cu:136-155).  bias /
 * residual may be NULL as in the kernel (:112,120).  row_begin/row_end select a row block so
 * the 8192^3 configuration can be sampled.
 * ------------------------------------------------------------------------------------- */
void fa_oracle_gemm_fused(const float* A, const float* B, const float* bias, const float* residual, float* C, int M,
                          int N, int K, int row_begin, int row_end, int nthreads) {
  (void)M;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = row_begin; i < row_end; ++i) {
    for (int j = 0; j < N; ++j) {
      float acc = 0.0f;
      for (int k = 0; k < K; ++k) acc += A[(size_t)i * K + k] * B[(size_t)k * N + j];
      float val = acc + (bias ? bias[j] : 0.0f);
      val = val > 0.0f ? val : 0.0f;
      val += residual ? residual[(size_t)i * N + j] : 0.0f;
      C[(size_t)i * N + j] = val;
    }
  }
}

/* Activation between bias and residual.  1 = the reference's ReLU (`val > 0.0f ? val : 0.0f`,
 * tensorcore_gemm_fused_epilogue.cu:150).  0 / 2 / 3 (identity, exact GELU, SiLU) are this engine's extension of
 * the epilogue surface: the reference has no such code, so they are PARITY-UNPINNED and
 * checked against their defining formulas in float64 (tests/test_oracle.py). */
static float oracle_act(float val, int act) {
  switch (act) {
    case 1: return val > 0.0f ? val : 0.0f;
    case 2: return 0.5f * val * (1.0f + erff(val * 0.70710678118654752f));
    case 3: return val / (1.0f + expf(-val));
    default: return val;
  }
}

/* The same contraction on fp16 operands widened to fp32 (what the GPU kernel is fed, :192-195);
 * rows [row_begin,row_end) of A only; out is [row_end-row_begin, N]. */
void fa_oracle_gemm_fused_f16_act(const uint16_t* A, const uint16_t* B, const float* bias, const float* residual,
                                  float* out, int M, int N, int K, int row_begin, int row_end, int act, int nthreads) {
  (void)M;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
  float* Bf = (float*)malloc(sizeof(float) * (size_t)K * N);
  fa_oracle_half_to_float(B, Bf, (size_t)K * N);
#pragma omp parallel
  {
    float* arow = (float*)malloc(sizeof(float) * (size_t)K);
    float* acc = (float*)malloc(sizeof(float) * (size_t)N);
#pragma omp for schedule(dynamic, 1)
    for (int i = row_begin; i < row_end; ++i) {
      fa_oracle_half_to_float(A + (size_t)i * K, arow, (size_t)K);
      /* k-outer order keeps every acc[j] a plain k-ascending fp32 sum, as in the reference */
      for (int j = 0; j < N; ++j) acc[j] = 0.0f;
      for (int k = 0; k < K; ++k) {
        const float a = arow[k];
        const float* brow = Bf + (size_t)k * N;
        for (int j = 0; j < N; ++j) acc[j] += a * brow[j];
      }
      float* o = out + (size_t)(i - row_begin) * N;
      for (int j = 0; j < N; ++j) {
        float val = acc[j] + (bias ? bias[j] : 0.0f);
        val = oracle_act(val, act);
        val += residual ? residual[(size_t)i * N + j] : 0.0f;
        o[j] = val;
      }
    }
    free(arow);
    free(acc);
  }
  free(Bf);
}
void fa_oracle_gemm_fused_f16(const uint16_t* A, const uint16_t* B, const float* bias, const float* residual,
                              float* out, int M, int N, int K, int row_begin, int row_end, int nthreads) {
  fa_oracle_gemm_fused_f16_act(A, B, bias, residual, out, M, N, K, row_begin, row_end, 1, nthreads);
}

/* ---------------------------------------------------------------------------------------
 * Batched plain GEMM, C[b] = A[b]*B[b], fp32 accumulate, k ascending -- restates
 * cpu_batched_gemm_ref (GEMM/gemm_tensorcore_batched/gemm_4_9_tensorcore_batched.cu:94-121).
 * A [batch,M,K], B [batch,K,N], C [batch,M,N], fp32.  The reference harness hands it the UNROUNDED
 * fp32 operands (:145) while its kernel gets fp16 copies (:150-151); the parity tests here feed
 * it the half-rounded values widened to float, i.e. exactly what the GPU sees.
 * ------------------------------------------------------------------------------------- */
void fa_oracle_gemm_batched(const float* A, const float* B, float* C, int batch, int M, int N, int K, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
#pragma omp parallel for collapse(2) schedule(dynamic, 8)
  for (int b = 0; b < batch; ++b) {
    for (int i = 0; i < M; ++i) {
      const float* Ab = A + (size_t)b * M * K;
      const float* Bb = B + (size_t)b * K * N;
      float* Cb = C + (size_t)b * M * N;
      for (int j = 0; j < N; ++j) {
        float acc = 0.0f;
        for (int k = 0; k < K; ++k) acc += Ab[(size_t)i * K + k] * Bb[(size_t)k * N + j];
        Cb[(size_t)i * N + j] = acc;
      }
    }
  }
}

/* ---------------------------------------------------------------------------------------
 * Scores[b,i,j] = scale * sum_k Q[b,i,k] K[b,j,k], double accumulate, one final rounding --
 * restates qk_cpu_ref (GEMM/QK Tensor Core/qk_wmma_5_1_2.cu:101-122).  Q [BH,N,D], Kmat [BH,Nk,D];
 * the reference has Nk == N, the extra extent costs nothing here.
 * ------------------------------------------------------------------------------------- */
void fa_oracle_qk_scores(const float* Q, const float* Kmat, float* S, int BH, int N, int Nk, int D, float scale,
                         int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
#pragma omp parallel for collapse(2) schedule(dynamic, 8)
  for (int b = 0; b < BH; ++b) {
    for (int i = 0; i < N; ++i) {
      for (int j = 0; j < Nk; ++j) {
        double acc = 0.0;
        for (int k = 0; k < D; ++k) {
          const float qv = Q[((size_t)b * N + i) * D + k];
          const float kv = Kmat[((size_t)b * Nk + j) * D + k];
          acc += (double)qv * kv;
        }
        acc *= scale;
        S[((size_t)b * N + i) * Nk + j] = (float)acc;
      }
    }
  }
}

/* ---------------------------------------------------------------------------------------
 * The streaming family's host oracle: S = (Q K) * scale materialised, row softmax with the
 * `+EPS` normaliser (EPS = 1e-6), O = P V -- restates flashattn_streaming_cpu_ref
 * (Streaming_FlashAttention_Forward_Kernel/flashattn_streaming_16x16.cu:280-348).
 * Layouts as in that file: Q half [B, M, Kd], K half [B, Kd, L] (TRANSPOSED: key index fastest),
 * V half [B, L, Dv]; the file fixes M = Kd = Dv = 16, the restatement keeps them as arguments.
 * ------------------------------------------------------------------------------------- */
void fa_oracle_streaming_ref_f16(const uint16_t* Q, const uint16_t* Kt, const uint16_t* V, float* O, int B, int M, int L,
                                 int Kd, int Dv, float scale) {
  const float NEG_LARGE = -1e30f, EPS = 1e-6f;
  float* S = (float*)malloc(sizeof(float) * (size_t)M * L);
  float* P = (float*)malloc(sizeof(float) * (size_t)M * L);
  for (int b = 0; b < B; ++b) {
    const uint16_t* Qb = Q + (size_t)b * M * Kd;
    const uint16_t* Kb = Kt + (size_t)b * Kd * L;
    const uint16_t* Vb = V + (size_t)b * L * Dv;
    float* Ob = O + (size_t)b * M * Dv;
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < L; ++j) {
        float acc = 0.0f;
        for (int k = 0; k < Kd; ++k) acc += half_bits_to_float(Qb[i * Kd + k]) * half_bits_to_float(Kb[(size_t)k * L + j]);
        S[(size_t)i * L + j] = acc * scale;
      }
    for (int i = 0; i < M; ++i) {
      float maxv = NEG_LARGE;
      for (int j = 0; j < L; ++j)
        if (S[(size_t)i * L + j] > maxv) maxv = S[(size_t)i * L + j];
      float sumv = 0.0f;
      for (int j = 0; j < L; ++j) {
        const float e = expf(S[(size_t)i * L + j] - maxv);
        P[(size_t)i * L + j] = e;
        sumv += e;
      }
      const float inv = 1.0f / (sumv + EPS);
      for (int j = 0; j < L; ++j) P[(size_t)i * L + j] *= inv;
    }
    for (int i = 0; i < M; ++i)
      for (int d = 0; d < Dv; ++d) {
        float acc = 0.0f;
        for (int j = 0; j < L; ++j) acc += P[(size_t)i * L + j] * half_bits_to_float(Vb[(size_t)j * Dv + d]);
        Ob[i * Dv + d] = acc;
      }
  }
  free(S);
  free(P);
}

/* ---------------------------------------------------------------------------------------
 * STAGE oracle: the streaming (online-softmax) state after every key tile.  Per tile of `tile`
 * keys: m_t = max_j s_ij, l_t = sum_j exp(s_ij - m_t), y_t[d] = sum_j exp(s_ij - m_t) v_jd; then
 * the tile-to-tile merge  m' = max(m, m_t), alpha = exp(m - m'), beta = exp(m_t - m'),
 * l' = l alpha + l_t beta, y' = y alpha + y_t beta, with the first tile copied and empty tiles
 * skipped -- restates the kernel-side rule of
 * Streaming_FlashAttention_Forward_Kernel/flashattn_streaming_16x16.cu:170-261 (tile statistics
 * :186-223, merge :228-261; `__expf` there, expf here; warp-butterfly sums there, j-ascending
 * here).  Q [B,M,D], K [B,L,D] (row per key), V [B,L,Dv] half.  Outputs hold one snapshot per tile:
 * m, l [T, B, M], y [T, B, M, Dv] with T = ceil(L / tile); the last tile may be partial.
 * The reference has no host function for this rule (it lives in device code only): the
 * restatement is pinned through its invariants -- y/l of the last snapshot equals cpu_reference,
 * m equals the running row max, snapshots do not depend on the tile size beyond rounding -- and
 * against the reference's own streaming kernel run on the GPU (tests/test_gpu_stage_ladder.py).
 * ------------------------------------------------------------------------------------- */
void fa_oracle_streaming_state_f16(const uint16_t* Q, const uint16_t* K, const uint16_t* V, float* m_out, float* l_out,
                                   float* y_out, int B, int M, int L, int D, int Dv, int tile, float scale) {
  const float NEG_LARGE = -1e30f;
  const int T = (L + tile - 1) / tile;
  float* s = (float*)malloc(sizeof(float) * (size_t)tile);
  float* yt = (float*)malloc(sizeof(float) * (size_t)Dv);
  float* y = (float*)malloc(sizeof(float) * (size_t)Dv);
  for (int b = 0; b < B; ++b) {
    for (int i = 0; i < M; ++i) {
      const uint16_t* q = Q + ((size_t)b * M + i) * D;
      float m = NEG_LARGE, l = 0.0f;
      for (int d = 0; d < Dv; ++d) y[d] = 0.0f;
      for (int t = 0; t < T; ++t) {
        const int j0 = t * tile, j1 = (j0 + tile < L) ? j0 + tile : L;
        float mt = NEG_LARGE;
        for (int j = j0; j < j1; ++j) {
          const uint16_t* k = K + ((size_t)b * L + j) * D;
          float acc = 0.0f;
          for (int c = 0; c < D; ++c) acc += half_bits_to_float(q[c]) * half_bits_to_float(k[c]);
          s[j - j0] = acc * scale;
          mt = fmaxf(mt, s[j - j0]);
        }
        float lt = 0.0f;
        for (int d = 0; d < Dv; ++d) yt[d] = 0.0f;
        for (int j = j0; j < j1; ++j) {
          const float e = expf(s[j - j0] - mt);
          lt += e;
          const uint16_t* v = V + ((size_t)b * L + j) * Dv;
          for (int d = 0; d < Dv; ++d) yt[d] += e * half_bits_to_float(v[d]);
        }
        if (lt > 0.0f) {
          if (m == NEG_LARGE) {
            m = mt;
            l = lt;
            for (int d = 0; d < Dv; ++d) y[d] = yt[d];
          } else {
            const float mn = fmaxf(m, mt);
            const float alpha = expf(m - mn), beta = expf(mt - mn);
            l = l * alpha + lt * beta;
            for (int d = 0; d < Dv; ++d) y[d] = y[d] * alpha + yt[d] * beta;
            m = mn;
          }
        }
        const size_t row = ((size_t)t * B + b) * M + i;
        m_out[row] = m;
        l_out[row] = l;
        for (int d = 0; d < Dv; ++d) y_out[row * Dv + d] = y[d];
      }
    }
  }
  free(s);
  free(yt);
  free(y);
}

/* ---------------------------------------------------------------------------------------
 * Error metrics, adopted verbatim: rel_l2 = sqrt(sum (a-b)^2 / (sum b^2 + 1e-30))
 * (flashattn_streaming_16x32_2warp_softmax_v12i.cu:288-296); max_abs = max |a-b| in double
 * (tensorcore_gemm_fused_epilogue.cu:247-254).
 * ------------------------------------------------------------------------------------- */
double fa_oracle_rel_l2(const float* a, const float* b, size_t n) {
  double num = 0.0, den = 0.0;
  for (size_t i = 0; i < n; ++i) {
    const double diff = (double)a[i] - (double)b[i];
    num += diff * diff;
    den += (double)b[i] * (double)b[i];
  }
  return sqrt(num / (den + 1e-30));
}

double fa_oracle_max_abs(const float* a, const float* b, size_t n) {
  double m = 0.0;
  for (size_t i = 0; i < n; ++i) {
    const double d = fabs((double)a[i] - (double)b[i]);
    if (d > m) m = d;
  }
  return m;
}

int fa_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
