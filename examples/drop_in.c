/* A plain-C consumer of libfa_b200.so: the flow of the reference's main()s (allocate on the device, copy in, one
 * launch line, copy out -- e.g. flashattn_warp_spc/flashattn_streaming_16x32_2warp_softmax_v12i.cu:325-368) with the
 * launch line replaced by the C-ABI call.  Built by tests/test_abi.py with gcc (no CUDA headers needed: the *_host
 * entry points take host pointers); on a box without an sm_100 GPU it reports FA_ERR_NO_DEVICE and exits 2.
 *   gcc -std=c99 -Iinclude examples/drop_in.c -o drop_in -Lflashattention_kernel_project_b200 -lfa_b200 -lm
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "fa_b200.h"

/* IEEE binary16 from float, round to nearest even (inputs here are small normal numbers) */
static unsigned short to_half(float f) {
  union { float f; unsigned u; } v;
  v.f = f;
  unsigned sign = (v.u >> 16) & 0x8000u, mant = v.u & 0x7fffffu;
  int exp = (int)((v.u >> 23) & 0xff) - 127 + 15;
  if (exp <= 0) return (unsigned short)sign;
  if (exp >= 31) return (unsigned short)(sign | 0x7c00u);
  unsigned h = sign | ((unsigned)exp << 10) | (mant >> 13);
  unsigned rem = mant & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;
  return (unsigned short)h;
}
static float from_half(unsigned short h) {
  int exp = (h >> 10) & 31;
  float m = (float)(h & 0x3ff) / 1024.0f;
  float v = exp == 0 ? ldexpf(m, -14) : ldexpf(1.0f + m, exp - 15);
  return (h & 0x8000) ? -v : v;
}

int main(void) {
  const int B = 1, H = 2, N = 256, D = 64;
  const size_t n = (size_t)B * H * N * D;
  unsigned short *q = malloc(n * 2), *k = malloc(n * 2), *v = malloc(n * 2), *o = malloc(n * 2);
  unsigned s = 123u;
  for (size_t i = 0; i < 3 * n; ++i) {
    s = s * 1664525u + 1013904223u;
    float x = (float)(s >> 8) / 8388608.0f - 1.0f; /* uniform(-1,1) */
    (i < n ? q : i < 2 * n ? k : v)[i % n] = to_half(x);
  }
  printf("fa_version %d\n", fa_version());
  int rc = fa_forward_host(q, k, v, o, B, H, N, D, 0.125f);
  if (rc != FA_OK) {
    printf("fa_forward_host: %d (%s)\n", rc, fa_strerror(rc));
    return rc == FA_ERR_NO_DEVICE ? 2 : 1;
  }
  /* check one row against a direct evaluation: O[0,0,5,:] */
  double max_err = 0.0;
  const int row = 5;
  float sc[256];
  float mx = -1e30f, sum = 0.0f;
  for (int j = 0; j < N; ++j) {
    float acc = 0.0f;
    for (int d = 0; d < D; ++d) acc += from_half(q[row * D + d]) * from_half(k[j * D + d]);
    sc[j] = acc * 0.125f;
    if (sc[j] > mx) mx = sc[j];
  }
  for (int j = 0; j < N; ++j) { sc[j] = expf(sc[j] - mx); sum += sc[j]; }
  for (int d = 0; d < D; ++d) {
    float acc = 0.0f;
    for (int j = 0; j < N; ++j) acc += sc[j] * from_half(v[j * D + d]);
    double e = fabs((double)acc / sum - (double)from_half(o[row * D + d]));
    if (e > max_err) max_err = e;
  }
  printf("row %d max abs error %.3e (limit 2e-3), kernels launched %llu\n", row, max_err, fa_launch_count());
  fa_host_release();
  free(q); free(k); free(v); free(o);
  return max_err <= 2e-3 ? 0 : 1;
}
