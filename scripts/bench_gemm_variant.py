"""Developer tool: time + check gemm_fused_epilogue from alternative builds of the library."""
import ctypes
import sys

import torch


def run(path):
    lib = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
    vp = ctypes.c_void_p
    lib.gemm_fused_epilogue.argtypes = [vp] * 5 + [ctypes.c_int] * 3 + [vp]
    res = []
    for (M, N, K, iters) in [(1000, 1096, 520, 3), (8192, 8192, 8192, 10), (4096, 4096, 4096, 20), (16384, 8192, 2048, 10)]:
        g = torch.Generator(device="cuda").manual_seed(0)
        a = (torch.rand((M, K), device="cuda", generator=g) * 2 - 1).half()
        b = (torch.rand((K, N), device="cuda", generator=g) * 2 - 1).half()
        bias = torch.rand((N,), device="cuda", generator=g)
        r = torch.rand((M, N), device="cuda", generator=g)
        c = torch.empty((M, N), device="cuda")
        f = lambda: lib.gemm_fused_epilogue(a.data_ptr(), b.data_ptr(), bias.data_ptr(), r.data_ptr(), c.data_ptr(), M, N, K, None)
        for _ in range(2):
            rc = f()
            assert rc == 0, rc
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            f()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        rows = slice(M - 200, M)
        ref = torch.relu(a[rows].float() @ b.float() + bias) + r[rows]
        err = (c[rows] - ref).abs().max().item()
        res.append(f"[{M}x{N}x{K}] {ms:.4f} ms {2.0 * M * N * K / ms / 1e9:7.1f} TF/s max_abs={err:.1e}")
    print(path.split("/")[-1], " | ".join(res), flush=True)


for p in sys.argv[1:]:
    run(p)
