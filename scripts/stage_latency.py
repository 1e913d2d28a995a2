"""Per-stage latency of the unfused attention pipeline next to the fused kernel, on the shapes of the reference's
study (FlashAttention/flashattn_forward_memory_bound/flashattn_stage_latency_breakdown.cu:212-424: N=512, D=64, BH=4)
and on BASELINE configs 2 and a slice of 3.  Usage (GPU box): python scripts/stage_latency.py [--json out.json]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flashattention_kernel_project_b200 as fa  # noqa: E402


def timed(fn, iters=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def study(BH, N, D):
    g = torch.Generator(device="cuda").manual_seed(1)
    q, k, v = ((torch.rand((1, BH, N, D), device="cuda", generator=g) * 2 - 1).half() for _ in range(3))
    q3, k3, v3 = (t.reshape(BH, N, D) for t in (q, k, v))
    S = torch.empty((BH, N, N), device="cuda")
    P = torch.empty((BH, N, N), device="cuda", dtype=torch.float16)
    O = torch.empty((BH, N, D), device="cuda")
    scale = D ** -0.5
    r = {"BH": BH, "N": N, "D": D}
    r["stage_a_qk_us"] = timed(lambda: fa.gemm_batched(q3, k3, trans_b=True, alpha=scale, out=S))
    r["stage_b_softmax_us"] = timed(lambda: fa.softmax_rows_f16out(S.view(BH * N, N), out=P.view(BH * N, N)))
    r["stage_c_pv_us"] = timed(lambda: fa.gemm_batched(P, v3, out=O))
    r["unfused_total_us"] = timed(lambda: fa.fa_forward_unfused(q, k, v, scale, workspace=(S, P)))
    r["fused_us"] = timed(lambda: fa.fa_forward_f32out(q, k, v, scale))
    flops = 4.0 * BH * N * N * D
    r["unfused_tflops"] = flops / r["unfused_total_us"] / 1e6
    r["fused_tflops"] = flops / r["fused_us"] / 1e6
    r["softmax_gbs"] = BH * N * N * 6 / r["stage_b_softmax_us"] / 1e3
    return r


def main():
    out = [study(4, 512, 64), study(64, 1024, 64), study(32, 4096, 128)]
    for r in out:
        print(json.dumps(r))
    if len(sys.argv) > 2 and sys.argv[1] == "--json":
        with open(sys.argv[2], "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
