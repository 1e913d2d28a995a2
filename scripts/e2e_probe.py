"""Developer probe: fa_forward_host (pinned host buffers, H2D + kernel + D2H pipelined over chunks of (b,h)) against the
bare-copy ceiling, for the full config 3 (256 (b,h)) and for the 32-(b,h) shard one rank owns at 8 GPUs, over the chunk
bound FA_HOST_CHUNKS (read once per process: one subprocess per setting).  python scripts/e2e_probe.py"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    import flashattention_kernel_project_b200 as fa
    from bench import link_probe
    N, D = 4096, 128
    for BH in (32, 256):
        q, k, v = ((torch.rand((1, BH, N, D)) * 2 - 1).half().pin_memory() for _ in range(3))
        o = torch.empty_like(q).pin_memory()
        for _ in range(2):
            fa.fa_forward_host(q, k, v, o, D ** -0.5)
        n = 8
        t0 = time.perf_counter()
        for _ in range(n):
            fa.fa_forward_host(q, k, v, o, D ** -0.5)
        dt = (time.perf_counter() - t0) / n
        link = link_probe(torch, torch.device("cuda", 0), 3 * q.numel() * 2, q.numel() * 2) * 1e-3
        print(f"chunks<={os.environ.get('FA_HOST_CHUNKS', 'default')} BH={BH}: {dt * 1e3:.2f} ms "
              f"{4.0 * BH * N * N * D / dt / 1e12:.1f} TFLOP/s, bare copies {link * 1e3:.2f} ms, link_frac {link / dt:.3f}", flush=True)
        fa.load_library().fa_host_release()
    sys.exit(0)
for chunks in ("1", "2", "4", "8", "16", "32"):
    env = dict(os.environ, FA_HOST_CHUNKS=chunks)
    r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
    print(r.stdout.strip() if r.returncode == 0 else "FAILED " + r.stderr[-500:], flush=True)
