"""Developer tool: the stream-K cut of the last wave on / off (FA_TAIL_SPLIT, FA_TAIL_COST are read once per process,
so every setting runs in its own subprocess).  python scripts/tail_split_ab.py"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = [(1, 32, 4096, 128), (2, 32, 4096, 128), (4, 32, 4096, 128), (8, 32, 4096, 128), (4, 16, 1024, 64),
          (1, 1, 4096, 128), (1, 32, 8192, 128), (8, 32, 4096, 64), (1, 8, 1024, 64)]

if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    import flashattention_kernel_project_b200 as fa
    out = {}
    for (B, H, N, D) in SHAPES:
        g = torch.Generator(device="cuda").manual_seed(0)
        q, k, v = ((torch.rand((B, H, N, D), device="cuda", generator=g) * 2 - 1).half() for _ in range(3))
        o = torch.empty_like(q)
        f = lambda: fa.fa_forward(q, k, v, D ** -0.5, out=o)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        iters = max(10, int(20e9 / (B * H * N * N * D) * 40))
        iters = min(iters, 400)
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                f()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / iters)
        # CUDA-graph time (host cost off the critical path)
        gr = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            f()
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(gr):
            for _ in range(10):
                f()
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        gms = e0.elapsed_time(e1) / 100
        ref = torch.softmax(q[:1, :1].float() @ k[:1, :1].float().transpose(-1, -2) * D ** -0.5, dim=-1) @ v[:1, :1].float()
        d = (o[:1, :1].float() - ref).double()
        rel = (d.pow(2).sum() / ref.double().pow(2).sum()).sqrt().item()
        out[f"{B}x{H}x{N}x{D}"] = {"ms": best, "graph_ms": gms, "tflops": 4.0 * B * H * N * N * D / min(best, gms) / 1e9, "rel_l2": rel}
    print(json.dumps(out))
    sys.exit(0)

res = {}
for name, env in [("off", {"FA_TAIL_SPLIT": "0"}), ("cost2", {}), ("cost1", {"FA_TAIL_COST": "1"}), ("cost3", {"FA_TAIL_COST": "3"}),
                  ("off_again", {"FA_TAIL_SPLIT": "0"})]:
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, __file__, "child"], env=e, capture_output=True, text=True)
    if r.returncode != 0:
        print(name, "FAILED", r.stderr[-2000:])
        continue
    res[name] = json.loads(r.stdout.strip().splitlines()[-1])
for shape in res.get("off", {}):
    print(shape, " | ".join(f"{n}: {res[n][shape]['ms'] * 1e3:8.1f} us g {res[n][shape]['graph_ms'] * 1e3:8.1f} us {res[n][shape]['tflops']:7.1f} TF/s rel {res[n][shape]['rel_l2']:.1e}" for n in res))
