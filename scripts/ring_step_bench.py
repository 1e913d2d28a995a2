"""Developer tool: the P partial-attention steps of BASELINE config 4 as one rank of P runs them (no exchange), on one
GPU: what the ring costs when the exchange is hidden.  python scripts/ring_step_bench.py [P ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flashattention_kernel_project_b200 as fa

for P in [int(x) for x in sys.argv[1:]] or [8, 4, 2]:
    B, H, N, D = 1, 32, 65536, 128
    n = N // P
    q, k, v = ((torch.rand((B, H, n, D), device="cuda") * 2 - 1).half() for _ in range(3))
    o_acc = torch.empty((B, H, n, D), dtype=torch.float32, device="cuda")
    m = torch.empty((B, H, n), dtype=torch.float32, device="cuda")
    l = torch.empty((B, H, n), dtype=torch.float32, device="cuda")
    out = torch.empty_like(q)
    o32 = torch.empty((B, H, n, D), dtype=torch.float32, device="cuda")

    def steps():
        for s in range(P):
            fa.fa_forward_partial(q, k, v, o_acc, m, l, D ** -0.5, s == 0, s == P - 1, out)

    def timeit(f, it):
        f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(it):
            f()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / it
    ms = timeit(steps, 5)
    ms_plain = timeit(lambda: fa.fa_forward(q, k, v, D ** -0.5, out=out), 10)
    ms_f32 = timeit(lambda: fa.fa_forward_f32out(q, k, v, D ** -0.5, out=o32), 10)
    print(f"P={P}: {P} partial steps {ms:.3f} ms ({ms / P * 1e3:.1f} us per step); one fa_forward of the same shape "
          f"{ms_plain * 1e3:.1f} us, fa_forward_f32out {ms_f32 * 1e3:.1f} us", flush=True)
