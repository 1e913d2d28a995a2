import sys, time, os
sys.path.insert(0, "/root/repo")
import torch
import flashattention_kernel_project_b200 as fa
B, H, N, D = 8, 32, 4096, 128
q, k, v = ((torch.rand((B, H, N, D)) * 2 - 1).half().pin_memory() for _ in range(3))
o = torch.empty_like(q).pin_memory()
for _ in range(2):
    fa.fa_forward_host(q, k, v, o, D ** -0.5)
t0 = time.perf_counter()
n = 8
for _ in range(n):
    fa.fa_forward_host(q, k, v, o, D ** -0.5)
dt = (time.perf_counter() - t0) / n
print(os.environ.get("FA_HOST_CHUNKS", "16"), f"{dt*1e3:.2f} ms", f"{4.0*B*H*N*N*D/dt/1e12:.1f} TFLOP/s", f"H2D-equivalent {3*q.numel()*2/dt/1e9:.1f} GB/s")
