"""Developer tool: two gemm_fused_epilogue launches (8192^3) from a given build of the library (ncu target).
usage: python scripts/run_gemm_variant_once.py variants/libfa_x.so"""
import ctypes
import sys

import torch

L = ctypes.CDLL(sys.argv[1], mode=ctypes.RTLD_LOCAL)
L.gemm_fused_epilogue.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int] * 3 + [ctypes.c_void_p]
M = 8192
a = (torch.rand((M, M), device="cuda") * 2 - 1).half()
b = (torch.rand((M, M), device="cuda") * 2 - 1).half()
bias, res, c = torch.rand((M,), device="cuda"), torch.rand((M, M), device="cuda"), torch.empty((M, M), device="cuda")
for _ in range(2):
    assert L.gemm_fused_epilogue(a.data_ptr(), b.data_ptr(), bias.data_ptr(), res.data_ptr(), c.data_ptr(), M, M, M, None) == 0
torch.cuda.synchronize()
print("ok")
