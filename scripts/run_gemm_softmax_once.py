"""Developer tool: a few gemm_fused_epilogue (8192^3) and softmax_rows (16384 x 4096) launches (target for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flashattention_kernel_project_b200 as fa
M = 8192
a = (torch.rand((M, M), device="cuda") * 2 - 1).half()
b = (torch.rand((M, M), device="cuda") * 2 - 1).half()
bias, res, c = torch.rand((M,), device="cuda"), torch.rand((M, M), device="cuda"), torch.empty((M, M), device="cuda")
x = torch.randn((16384, 4096), device="cuda")
p = torch.empty_like(x)
for _ in range(3):
    fa.gemm_fused_epilogue(a, b, bias, res, out=c)
    fa.softmax_rows(x, out=p)
torch.cuda.synchronize()
print("ok")
