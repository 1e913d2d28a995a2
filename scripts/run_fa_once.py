"""Developer tool: a few fa_forward launches at config-3 size (target for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flashattention_kernel_project_b200 as fa
B, H, N, D = 8, 32, 4096, int(os.environ.get("FA_D", "128"))
q, k, v = ((torch.rand((B, H, N, D), device="cuda") * 2 - 1).half() for _ in range(3))
o = torch.empty_like(q)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    fa.fa_forward(q, k, v, D ** -0.5, out=o)
torch.cuda.synchronize()
print("ok")
