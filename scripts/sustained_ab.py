"""Developer tool: sustained (power-capped) throughput of alternative builds, interleaved A B A B ... on one box so that
thermal drift hits them equally.  usage: python scripts/sustained_ab.py lib1.so lib2.so ... (paths relative to repo)"""
import ctypes
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = []
for p in sys.argv[1:]:
    L = ctypes.CDLL(os.path.join(ROOT, p), mode=ctypes.RTLD_LOCAL)
    L.fa_forward.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 4 + [ctypes.c_float, ctypes.c_void_p]
    libs.append((os.path.basename(p), L))
B, H, N, D = 8, 32, 4096, int(os.environ.get("FA_D", "128"))
q, k, v = ((torch.rand((B, H, N, D), device="cuda") * 2 - 1).half() for _ in range(3))
o = torch.empty_like(q)
fl = 4.0 * B * H * N * N * D


def block(L, seconds):
    run = lambda: L.fa_forward(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, H, N, D, D ** -0.5, None)
    t_end = time.time() + seconds * 0.4
    while time.time() < t_end:  # settle
        for _ in range(20):
            run()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    e0.record()
    t_end = time.time() + seconds * 0.6
    while time.time() < t_end:
        for _ in range(20):
            run()
        n += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return fl * n / (e0.elapsed_time(e1) * 1e-3) / 1e12


for name, L in libs:  # heat the part up
    block(L, 1.0)
res = {name: [] for name, _ in libs}
for rnd in range(4):
    for name, L in libs:
        res[name].append(block(L, 2.0))
for name, vals in res.items():
    print(f"{name:40s} " + " ".join(f"{x:7.1f}" for x in vals) + f"   mean {sum(vals) / len(vals):7.1f} TFLOP/s")
