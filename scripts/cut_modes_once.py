"""Developer tool: a few launches of fa_forward under each work-decomposition mode (fa_set_tail_split), for a per-kernel
ncu launch list:  ncu --metrics gpu__time_duration.sum --clock-control none --csv python scripts/cut_modes_once.py
The eager times it prints drift with the power state (the modes run one after the other); the interleaved comparison is
scripts/tail_split_ab.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flashattention_kernel_project_b200 as fa

lib = fa.load_library()
modes = tuple(int(x) for x in os.environ.get("AB_MODES", "0,1,2,3").split(","))
for (B, H, N, D) in [(1, 32, 4096, 128), (8, 32, 4096, 128)]:
    g = torch.Generator(device="cuda").manual_seed(0)
    q, k, v = ((torch.rand((B, H, N, D), device="cuda", generator=g) * 2 - 1).half() for _ in range(3))
    o = torch.empty_like(q)
    for mode in modes:
        lib.fa_set_tail_split(mode)
        for _ in range(3):
            fa.fa_forward(q, k, v, D ** -0.5, out=o)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fa.fa_forward(q, k, v, D ** -0.5, out=o)
        e1.record()
        torch.cuda.synchronize()
        print(f"{B}x{H}x{N}x{D} mode {mode}: eager {e0.elapsed_time(e1) * 100:.1f} us per launch", flush=True)
lib.fa_set_tail_split(-1)
