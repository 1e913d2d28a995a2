"""Developer tool: time + check fa_forward from alternative builds of the library.
Usage: python scripts/bench_variant.py lib1.so [lib2.so ...]"""
import ctypes
import sys

import torch


def run(path):
    lib = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
    lib.fa_forward.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 4 + [ctypes.c_float, ctypes.c_void_p]
    res = []
    for (B, H, N, D, iters) in [(8, 32, 4096, 128, 30), (4, 16, 1024, 64, 200), (2, 8, 16384, 128, 10), (8, 32, 4096, 64, 20), (4, 37, 2048, 64, 50)]:
        g = torch.Generator(device="cuda").manual_seed(0)
        q, k, v = ((torch.rand((B, H, N, D), device="cuda", generator=g) * 2 - 1).half() for _ in range(3))
        o = torch.empty_like(q)
        f = lambda: lib.fa_forward(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, H, N, D, D ** -0.5, None)
        for _ in range(3):
            assert f() == 0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            f()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        qs, ks, vs = q[:1, :2].float(), k[:1, :2].float(), v[:1, :2].float()
        ref = torch.softmax(qs @ ks.transpose(-1, -2) * D ** -0.5, dim=-1) @ vs
        d = (o[:1, :2].float() - ref).double()
        rel = (d.pow(2).sum() / ref.double().pow(2).sum()).sqrt().item()
        res.append(f"[{B}x{H}x{N}x{D}] {ms:.4f} ms {4.0 * B * H * N * N * D / ms / 1e9:7.1f} TF/s rel_l2={rel:.2e}")
    # peaked input (normal std 3) exercises rescale + wide exponent range of the polynomial path
    B, H, N, D = 1, 2, 2048, 128
    g = torch.Generator(device="cuda").manual_seed(1)
    q, k, v = ((torch.randn((B, H, N, D), device="cuda", generator=g) * 3).half() for _ in range(3))
    o = torch.empty_like(q)
    assert lib.fa_forward(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, H, N, D, D ** -0.5, None) == 0
    torch.cuda.synchronize()
    ref = torch.softmax(q.float() @ k.float().transpose(-1, -2) * D ** -0.5, dim=-1) @ v.float()
    d = (o.float() - ref).double()
    res.append(f"peaked rel_l2={(d.pow(2).sum() / ref.double().pow(2).sum()).sqrt().item():.2e}")
    print(path.split("/")[-1], " | ".join(res), flush=True)


for p in sys.argv[1:]:
    run(p)
