"""Developer tool: fa_forward_host on this rank's head shard of BASELINE config 3 over the chunk size of the host
pipeline (FA_HOST_CHUNK_MB, read per call), settings interleaved and repeated in ONE job so that the box's run-to-run
variation (the host side of this pool is bimodal) cancels.  Max over ranks per call, as in bench.py.
  python scripts/e2e_chunk_sweep.py                      (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 scripts/e2e_chunk_sweep.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import flashattention_kernel_project_b200 as fa
from flashattention_kernel_project_b200.sharding import shard_range

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
lo, hi = shard_range(256, world, rank)
N, D = 4096, 128
q, k, v = ((torch.rand((1, hi - lo, N, D)) * 2 - 1).half().pin_memory() for _ in range(3))
o = torch.empty_like(q).pin_memory()
settings = [int(x) for x in os.environ.get("SWEEP_MB", "32,16,8,4,2").split(",")]
times = {mb: [] for mb in settings}
flag = torch.zeros(1, device="cuda")
for rep in range(int(os.environ.get("SWEEP_REPS", "4")) + 1):
    for mb in settings:
        os.environ["FA_HOST_CHUNK_MB"] = str(mb)
        if world > 1:
            dist.all_reduce(flag)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(4):
            fa.fa_forward_host(q, k, v, o, D ** -0.5)
        dt = torch.tensor([(time.perf_counter() - t0) / 4], device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if rep > 0:  # the first round warms the scratch buffers up
            times[mb].append(dt.item() * 1e3)
if rank == 0:
    for mb in settings:
        t = sorted(times[mb])
        print(f"world {world} ({hi - lo} (b,h) per rank) copies >= {mb:2d} MB: ms per call min {t[0]:.2f} median {t[len(t) // 2]:.2f} max {t[-1]:.2f}"
              f" -> {4.0 * 256 * N * N * D / (t[len(t) // 2] * 1e-3) / 1e12:.1f} TFLOP/s whole job", flush=True)
if world > 1:
    dist.destroy_process_group()
