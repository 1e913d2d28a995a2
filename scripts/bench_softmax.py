import sys, os
sys.path.insert(0, "/root/repo")
import torch
import flashattention_kernel_project_b200 as fa
def t(f, it=100):
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
for rows, n in [(16384, 4096), (65536, 1024), (131072, 512), (16384, 8192), (8192, 2048), (1 << 20, 128)]:
    x = torch.randn((rows, n), device="cuda"); p = torch.empty_like(x)
    ms = t(lambda: fa.softmax_rows(x, out=p))
    print(f"{rows}x{n}: {ms*1e3:.1f} us {2*x.numel()*4/ms/1e6:.0f} GB/s")
