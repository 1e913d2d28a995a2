// In-harness drop-in proof (test infrastructure): the reference's OWN program
//   GEMM/tensorcore_gemm_fused_epilogue/tensorcore_gemm_fused_epilogue.cu
// compiled UNMODIFIED, its two `wmma_gemm_fused_kernel<<<grid, block>>>` launches (:220, :231) redirected to
// gemm_fused_epilogue by interposing cudaLaunchKernel (see dropin_v12i.cu).  The program's own verification is dead
// code (`#if DO_VERIFY` tests a constexpr int, :29,:190,:250), so the check it would have printed is done here, at
// the second launch, with the reference's own functions on the reference's own buffers: cpu_gemm_fused_ref (:136-155)
// on the half-rounded operands the kernels are fed, and the reference kernel itself (a forwarded, real launch).
#include <dlfcn.h>

#include "../../include/fa_b200.h"

#define main reference_main
#include "GEMM/tensorcore_gemm_fused_epilogue/tensorcore_gemm_fused_epilogue.cu"
#undef main

static int g_redirected = 0, g_failed = 0;
using LaunchFn = cudaError_t (*)(const void*, dim3, dim3, void**, size_t, cudaStream_t);
static LaunchFn real_launch() {
  static LaunchFn fn = reinterpret_cast<LaunchFn>(dlsym(RTLD_NEXT, "cudaLaunchKernel"));
  return fn;
}

static void verify(const half* dA, const half* dB, const float* dBias, const float* dRes, const float* dC, int m, int n,
                   int k, dim3 grid, dim3 block, void** args) {
  std::vector<half> hA((size_t)m * k), hB((size_t)k * n);
  std::vector<float> A((size_t)m * k), B((size_t)k * n), bias(n), res((size_t)m * n), ours((size_t)m * n),
      ref((size_t)m * n), theirs((size_t)m * n);
  CHECK_CUDA(cudaDeviceSynchronize());
  CHECK_CUDA(cudaMemcpy(hA.data(), dA, hA.size() * 2, cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(hB.data(), dB, hB.size() * 2, cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(bias.data(), dBias, bias.size() * 4, cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(res.data(), dRes, res.size() * 4, cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(ours.data(), dC, ours.size() * 4, cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < A.size(); ++i) A[i] = __half2float(hA[i]);
  for (size_t i = 0; i < B.size(); ++i) B[i] = __half2float(hB[i]);
  cpu_gemm_fused_ref(A, B, bias, res, ref, m, n, k);
  // the reference's kernel on the same buffers (a real launch), into a scratch C
  float* dC2 = nullptr;
  CHECK_CUDA(cudaMalloc(&dC2, ours.size() * 4));
  void* args2[8] = {args[0], args[1], args[2], args[3], &dC2, args[5], args[6], args[7]};
  CHECK_CUDA(real_launch()(reinterpret_cast<const void*>(&wmma_gemm_fused_kernel), grid, block, args2, 0, nullptr));
  CHECK_CUDA(cudaDeviceSynchronize());
  CHECK_CUDA(cudaMemcpy(theirs.data(), dC2, theirs.size() * 4, cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaFree(dC2));
  double max_ref = 0, max_kernel = 0, num = 0, den = 0;
  for (size_t i = 0; i < ours.size(); ++i) {
    const double d = std::fabs((double)ours[i] - ref[i]);
    max_ref = d > max_ref ? d : max_ref;
    const double e = std::fabs((double)ours[i] - theirs[i]);
    max_kernel = e > max_kernel ? e : max_kernel;
    num += d * d;
    den += (double)ref[i] * ref[i];
  }
  printf("drop-in: Max abs diff vs CPU fused float ref: %.6e\n", max_ref);
  printf("drop-in: Relative L2 error vs CPU fused float ref: %.6e\n", std::sqrt(num / den));
  printf("drop-in: Max abs diff vs wmma_gemm_fused_kernel: %.6e\n", max_kernel);
}

extern "C" cudaError_t cudaLaunchKernel(const void* func, dim3 grid, dim3 block, void** args, size_t shmem,
                                         cudaStream_t stream) {
  if (func == reinterpret_cast<const void*>(&wmma_gemm_fused_kernel)) {
    // kernel parameter list (:53-58): A half [M,K], B half [K,N], bias [N], residual [M,N], C [M,N], M, N, K
    const half* A = *static_cast<const half**>(args[0]);
    const half* B = *static_cast<const half**>(args[1]);
    const float* bias = *static_cast<const float**>(args[2]);
    const float* res = *static_cast<const float**>(args[3]);
    float* C = *static_cast<float**>(args[4]);
    const int m = *static_cast<int*>(args[5]), n = *static_cast<int*>(args[6]), k = *static_cast<int*>(args[7]);
    const int rc = gemm_fused_epilogue(A, B, bias, res, C, m, n, k, stream);
    if (rc != FA_OK) {
      ++g_failed;
      fprintf(stderr, "gemm_fused_epilogue failed: %d (%s)\n", rc, fa_strerror(rc));
      return cudaErrorLaunchFailure;
    }
    if (++g_redirected == 2) verify(A, B, bias, res, C, m, n, k, grid, block, args);
    return cudaSuccess;
  }
  return real_launch()(func, grid, block, args, shmem, stream);
}

int main() {
  const int rc = reference_main();
  printf("drop-in: %d launches redirected to gemm_fused_epilogue, %d failed\n", g_redirected, g_failed);
  return rc != 0 || g_failed != 0 || g_redirected == 0;
}
