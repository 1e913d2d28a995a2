"""Smallest end-to-end case for compute-sanitizer (one tool per gpurun call):
   compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flashattention_kernel_project_b200 as fa
q, k, v = ((torch.rand((1, 2, 300, 128), device="cuda") * 2 - 1).half() for _ in range(3))
o = fa.fa_forward(q, k, v, 0.09)
q, k, v = ((torch.rand((1, 1, 130, 64), device="cuda") * 2 - 1).half() for _ in range(3))
o2 = fa.fa_forward_f32out(q, k, v, 0.125)
a = (torch.rand((200, 136), device="cuda") - 0.5).half()
b = (torch.rand((136, 264), device="cuda") - 0.5).half()
c = fa.gemm_fused_epilogue(a, b, torch.rand(264, device="cuda"), torch.rand((200, 264), device="cuda"))
p = fa.softmax_rows(torch.randn((37, 1000), device="cuda"))
p2 = fa.softmax_rows(torch.randn((3, 4097), device="cuda"))
torch.cuda.synchronize()
print("sanitize_small ok", float(o.float().abs().sum()), float(c.sum()), float(p.sum()))
