"""Developer tool: the stream-K cut of the last wave on / off, interleaved in ONE process (fa_set_tail_split), so that
box-to-box and warm-up drift cancel.  python scripts/tail_split_ab.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import flashattention_kernel_project_b200 as fa

lib = fa.load_library()
MODES = tuple(int(x) for x in os.environ.get("AB_MODES", "0,2,3").split(","))  # 0 whole, 2 K/V cut, 3 row cut
SHAPES = [(1, 32, 4096, 128), (2, 32, 4096, 128), (4, 32, 4096, 128), (8, 32, 4096, 128), (4, 16, 1024, 64),
          (1, 1, 4096, 128), (1, 32, 8192, 128), (8, 32, 4096, 64), (1, 8, 1024, 64), (1, 4, 16384, 128)]
if os.environ.get("AB_SHAPES"):  # e.g. AB_SHAPES='1,32,4096,128;3,32,4096,128'
    SHAPES = [tuple(int(x) for x in t.split(",")) for t in os.environ["AB_SHAPES"].split(";")]
REPS = int(os.environ.get("AB_REPS", "6"))
for (B, H, N, D) in SHAPES:
    g = torch.Generator(device="cuda").manual_seed(0)
    q, k, v = ((torch.rand((B, H, N, D), device="cuda", generator=g) * 2 - 1).half() for _ in range(3))
    o = torch.empty_like(q)
    f = lambda: fa.fa_forward(q, k, v, D ** -0.5, out=o)
    iters = min(400, max(10, int(20e9 / (B * H * N * N * D) * 40)))
    graphs, outs = {}, {}
    for mode in MODES:
        lib.fa_set_tail_split(mode)
        for _ in range(3):
            f()
        outs[mode] = o.clone()
        gr = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            f()
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(gr):
            for _ in range(iters):
                f()
        graphs[mode] = gr
    torch.cuda.synchronize()
    best = {m: 1e9 for m in MODES}
    for rep in range(REPS):
        for mode in MODES:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graphs[mode].replay()
            e1.record()
            torch.cuda.synchronize()
            best[mode] = min(best[mode], e0.elapsed_time(e1) / iters)
    lib.fa_set_tail_split(-1)
    nseg = lib.fa_plan_segments(B, H, N, N, 148, None, 0)
    items = B * H * ((N + 255) // 256)
    fl = 4.0 * B * H * N * N * D
    names = {0: "whole", 1: "policy", 2: "kv-cut", 3: "row-cut"}
    print(f"{B}x{H}x{N}x{D}: items {items} segments(default policy) {nseg} | " + " | ".join(
        f"{names[m]} {best[m] * 1e3:8.1f} us {fl / best[m] / 1e9:7.1f} TF/s" + (
            f" x{best[m] / best[MODES[0]]:.3f} bit-identical={bool((outs[m] == outs[MODES[0]]).all())}" if m != MODES[0] else "")
        for m in MODES), flush=True)
