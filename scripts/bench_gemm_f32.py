"""Developer tool: gemm_fused_epilogue_f32 (fp32 operands, split tf32 pairs) next to the reference's fp32 kernel and
torch's fp32 / tf32 matmul.  python scripts/bench_gemm_f32.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flashattention_kernel_project_b200 as fa
from tests.oracle_lib import RefGpu, ref_available


def timeit(f, iters):
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for M in (1024, 4096, 8192):
    g = torch.Generator(device="cuda").manual_seed(42)
    a = torch.rand((M, M), device="cuda", generator=g) * 2 - 1
    b = torch.rand((M, M), device="cuda", generator=g) * 2 - 1
    bias = torch.rand((M,), device="cuda", generator=g) * 2 - 1
    res = torch.rand((M, M), device="cuda", generator=g) * 2 - 1
    c = torch.empty((M, M), device="cuda")
    ms = timeit(lambda: fa.gemm_fused_epilogue_f32(a, b, bias, res, out=c), 10 if M > 2048 else 50)
    ref64 = torch.relu(a[:64].double() @ b.double() + bias.double()) + res[:64].double()
    err = (c[:64].double() - ref64).abs().max().item()
    line = f"M=N=K={M}: ours {ms:.3f} ms {2.0 * M ** 3 / ms / 1e9:.1f} TFLOP/s (max abs err vs fp64, 64 rows: {err:.2e})"
    torch.backends.cuda.matmul.allow_tf32 = False
    ms32 = timeit(lambda: torch.matmul(a, b, out=c), 5)
    torch.backends.cuda.matmul.allow_tf32 = True
    mstf = timeit(lambda: torch.matmul(a, b, out=c), 10)
    errtf = (c[:64].double() - a[:64].double() @ b.double()).abs().max().item()
    torch.backends.cuda.matmul.allow_tf32 = False
    line += f" | torch fp32 {2.0 * M ** 3 / ms32 / 1e9:.1f} TFLOP/s, torch tf32 {2.0 * M ** 3 / mstf / 1e9:.1f} TFLOP/s (err {errtf:.2e})"
    if ref_available("gemm_fp32_fused") and M <= 4096:
        ref = RefGpu()
        msr = timeit(lambda: ref.gemm_fp32_fused(a, b, bias, res, c), 3)
        line += f" | reference gemm_tiled_fused_kernel {2.0 * M ** 3 / msr / 1e9:.1f} TFLOP/s"
    print(line, flush=True)
