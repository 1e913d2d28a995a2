"""Developer tool: sustained (power-capped) gemm_fused_epilogue 8192^3 throughput of alternative builds, interleaved
A B A B ... in 2 s blocks.  usage: python scripts/sustained_ab_gemm.py lib1.so lib2.so ..."""
import ctypes
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = []
for p in sys.argv[1:]:
    L = ctypes.CDLL(os.path.join(ROOT, p), mode=ctypes.RTLD_LOCAL)
    L.gemm_fused_epilogue.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int] * 3 + [ctypes.c_void_p]
    libs.append((os.path.basename(p), L))
M = 8192
a = (torch.rand((M, M), device="cuda") * 2 - 1).half()
b = (torch.rand((M, M), device="cuda") * 2 - 1).half()
bias, res, c = torch.rand((M,), device="cuda"), torch.rand((M, M), device="cuda"), torch.empty((M, M), device="cuda")
fl = 2.0 * M ** 3


def block(L, seconds):
    run = lambda: L.gemm_fused_epilogue(a.data_ptr(), b.data_ptr(), bias.data_ptr(), res.data_ptr(), c.data_ptr(), M, M, M, None)
    t_end = time.time() + seconds * 0.4
    while time.time() < t_end:
        for _ in range(40):
            run()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    e0.record()
    t_end = time.time() + seconds * 0.6
    while time.time() < t_end:
        for _ in range(40):
            run()
        n += 40
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return fl * n / (e0.elapsed_time(e1) * 1e-3) / 1e12


class _Cublas:  # yardstick on the same box, same harness: torch.matmul fp16 8192^3 (no epilogue, fp16 output)
    def gemm_fused_epilogue(self, *args):
        torch.matmul(a, b, out=c16)


c16 = torch.empty((M, M), device="cuda", dtype=torch.float16)
libs.append(("cuBLAS fp16 matmul (torch)", _Cublas()))
for name, L in libs:
    block(L, 1.0)
res_t = {name: [] for name, _ in libs}
for rnd in range(3):
    for name, L in libs:
        res_t[name].append(block(L, 2.0))
for name, vals in res_t.items():
    print(f"{name:40s} " + " ".join(f"{x:7.1f}" for x in vals) + f"   mean {sum(vals) / len(vals):7.1f} TFLOP/s")
