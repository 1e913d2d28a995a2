"""Developer tool: repeat the stream-K cut (general plan, FA_TAIL_SPLIT mode 2) on shapes with full waves + a cut tail and
on sub-wave grids, many times, concurrently on two streams, and check that every run is bit-identical to the first
(the merge is done by whichever CTA arrives last: the result must not depend on which one that is)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flashattention_kernel_project_b200 as fa

lib = fa.load_library()
lib.fa_set_tail_split(2)
side = torch.cuda.Stream()
bad = 0
for (B, H, N, D, reps) in [(1, 40, 4096, 128, 30), (3, 37, 1000, 64, 60), (1, 1, 16384, 128, 30), (1, 5, 2304, 128, 60), (2, 3, 8192, 64, 30)]:
    g = torch.Generator(device="cuda").manual_seed(N)
    q, k, v = ((torch.randn((B, H, N, D), device="cuda", generator=g)).half() for _ in range(3))
    first = fa.fa_forward(q, k, v, D ** -0.5)
    first32 = fa.fa_forward_f32out(q, k, v, D ** -0.5)
    torch.cuda.synchronize()
    for r in range(reps):
        a = fa.fa_forward(q, k, v, D ** -0.5)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            b = fa.fa_forward_f32out(q, k, v, D ** -0.5)
        c = fa.fa_forward(q, k, v, D ** -0.5)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if not (torch.equal(a, first) and torch.equal(c, first) and torch.equal(b, first32)):
            bad += 1
    segs = lib.fa_plan_segments(B, H, N, N, 148, None, 0)
    print(f"{B}x{H}x{N}x{D}: {segs} segments for {B * H * ((N + 255) // 256)} items, {reps} x 3 launches on two streams, mismatching runs so far: {bad}", flush=True)
lib.fa_set_tail_split(-1)
print("OK" if bad == 0 else "FAILED")
sys.exit(1 if bad else 0)
