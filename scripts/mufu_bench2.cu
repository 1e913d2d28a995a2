// Which softmax-loop instruction interferes with MUFU.EX2 on a sub-partition?  Warps w and w+4 share a sub-partition.
// Group 0 always runs EX2 chains; group 1 runs one other instruction type; then both groups run the realistic
// per-element mix of the attention softmax (per 2 keys: 1 FFMA2, 2 EX2, 1 FADD2, 1 F2FP, 1 FMNMX3).
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long y; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(y) : "l"(a), "l"(b), "l"(c)); return y; }
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
  unsigned long long y; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(y) : "l"(a), "l"(b)); return y; }
__device__ __forceinline__ unsigned f2fp(float a, float b) { unsigned y; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(b), "f"(a)); return y; }
__device__ __forceinline__ float fmax3(float a, float b, float c) { float y; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c)); return y; }
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }

constexpr int kIters = 1024;

// 1 EX2, 2 FFMA2, 3 FADD2, 4 F2FP, 5 FMNMX3, 6 realistic mix (8 element pairs per iteration), 0 idle
template <int ROLE>
__device__ float work(float seed) {
  float a[8], m = seed;
  unsigned long long p[8], acc = pk(0.f, 0.f);
  unsigned h = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed + i * 0.01f; p[i] = pk(a[i], a[i] * 0.5f); }
  for (int it = 0; it < kIters; ++it) {
    if (ROLE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = ex2(a[i]);
    } else if (ROLE == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = ffma2(p[i], 0x3f7fbe773f7fbe77ull, 0x3a83126f3a83126full);
    } else if (ROLE == 3) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fadd2(p[i], 0x3a83126f3a83126full);
    } else if (ROLE == 4) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { h ^= f2fp(a[i], a[(i + 1) & 7]); a[i] = __uint_as_float(__float_as_uint(a[i]) ^ (h & 1)); }
    } else if (ROLE == 5) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fmax3(a[i], a[(i + 3) & 7], m);
    } else if (ROLE == 6) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {  // one key pair
        float x0, x1;
        upk(ffma2(p[i], 0x3f7fbe773f7fbe77ull, 0xbf000000bf000000ull), x0, x1);
        const float e0 = ex2(x0), e1 = ex2(x1);
        acc = fadd2(acc, pk(e0, e1));
        h ^= f2fp(e0, e1);
        m = fmax3(m, x0, x1);
        p[i] = pk(e0, e1);
      }
    }
  }
  float s = m + __uint_as_float(h), t0, t1;
  upk(acc, t0, t1);
#pragma unroll
  for (int i = 0; i < 8; ++i) { float u0, u1; upk(p[i], u0, u1); s += a[i] + u0 + u1; }
  return s + t0 + t1;
}

template <int R0, int R1>
__global__ void __launch_bounds__(256, 1) k(float* out, long long* cycles, float seed) {
  const int g = threadIdx.x >> 7;
  __syncthreads();
  const long long t0 = clock64();
  float r = (g == 0) ? work<R0>(seed) : work<R1>(seed);
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)&cycles[blockIdx.x * 2 + g], (unsigned long long)(t1 - t0));
  if (r == 12345.678f) out[0] = r;
}

template <int R0, int R1>
void run(const char* name) {
  float* out; long long* cyc;
  cudaMalloc(&out, 4); cudaMalloc(&cyc, 16 * 148);
  for (int rep = 0; rep < 2; ++rep) { cudaMemset(cyc, 0, 16 * 148); k<R0, R1><<<148, 256>>>(out, cyc, 0.3f); }
  long long h[2];
  cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-40s g0 %8lld cyc = %6.2f / 8-group   g1 %8lld cyc = %6.2f / 8-group  %s\n", name, h[0], (double)h[0] / kIters, h[1],
         (double)h[1] / kIters, cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<1, 0>("EX2 alone");
  run<2, 0>("FFMA2 alone");
  run<3, 0>("FADD2 alone");
  run<4, 0>("F2FP(+LOP) alone");
  run<5, 0>("FMNMX3 alone");
  run<1, 2>("EX2 | FFMA2");
  run<1, 3>("EX2 | FADD2");
  run<1, 4>("EX2 | F2FP");
  run<1, 5>("EX2 | FMNMX3");
  run<6, 0>("softmax mix alone (16 keys / group)");
  run<6, 6>("softmax mix | softmax mix");
  return 0;
}
