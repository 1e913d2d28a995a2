// In-harness drop-in proof (test infrastructure, never on the product path): the reference's OWN program
//   flashattn_warp_spc/flashattn_streaming_16x32_2warp_softmax_v12i.cu
// compiled UNMODIFIED -- its main() (:298-393) generates the mt19937(123) fixture, launches its kernel 210 times,
// runs its cpu_reference (:243-286) and prints "Relative L2 error" (:381) -- with the kernel launch redirected to this
// library: the `<<<grid, block, shmem>>>` statements at :346-347 and :358-359 compile to cudaLaunchKernel(stub, ...),
// which this translation unit interposes (the program links the shared CUDA runtime, so the definition below wins;
// every other launch is forwarded to the real one).  The maintainer's edit that this emulates is the one-line
// replacement shown in INTEGRATION.md.
#include <dlfcn.h>

#include "../../include/fa_b200.h"

#define main reference_main
#include "flashattn_warp_spc/flashattn_streaming_16x32_2warp_softmax_v12i.cu"
#undef main

static int g_redirected = 0, g_forwarded = 0, g_failed = 0;

extern "C" cudaError_t cudaLaunchKernel(const void* func, dim3 grid, dim3 block, void** args, size_t shmem,
                                         cudaStream_t stream) {
  using Fn = cudaError_t (*)(const void*, dim3, dim3, void**, size_t, cudaStream_t);
  static Fn real = reinterpret_cast<Fn>(dlsym(RTLD_NEXT, "cudaLaunchKernel"));
  if (func == reinterpret_cast<const void*>(&flashattn_2warp_nosmemW_online_softmax<16, 16, 4, 32>)) {
    // kernel parameter list (:60-70): Q [B,M,16], K, V [B,L,16] half, O [B,M,16] float, seq_len, m_total, scale
    const void* Q = *static_cast<const void**>(args[0]);
    const void* K = *static_cast<const void**>(args[1]);
    const void* V = *static_cast<const void**>(args[2]);
    float* O = *static_cast<float**>(args[3]);
    const int seq_len = *static_cast<int*>(args[4]), m_total = *static_cast<int*>(args[5]);
    const float scale = *static_cast<float*>(args[6]);
    const int rc = fa_forward_cross(Q, K, V, O, static_cast<int>(grid.x), 1, m_total, seq_len, 16, scale, stream);
    ++g_redirected;
    if (rc != FA_OK) {
      ++g_failed;
      fprintf(stderr, "fa_forward_cross failed: %d (%s)\n", rc, fa_strerror(rc));
      return cudaErrorLaunchFailure;
    }
    return cudaSuccess;
  }
  ++g_forwarded;
  return real(func, grid, block, args, shmem, stream);
}

int main() {
  const int rc = reference_main();
  printf("drop-in: %d launches redirected to fa_forward_cross, %d forwarded, %d failed\n", g_redirected, g_forwarded,
         g_failed);
  return rc != 0 || g_failed != 0 || g_redirected == 0;
}
