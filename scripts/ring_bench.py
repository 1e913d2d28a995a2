"""BASELINE config 4 (B=1, H=32, N=65536, D=128) as ring / context-parallel attention over the ranks of a
torchrun job:  python -m torch.distributed.run --nproc-per-node P --master-addr 127.0.0.1 scripts/ring_bench.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import flashattention_kernel_project_b200 as fa

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, H, N, D = 1, 32, 65536, 128
n = N // world
g = torch.Generator(device=dev).manual_seed(7 + rank)
q, k, v = ((torch.rand((B, H, n, D), device=dev, generator=g) * 2 - 1).half() for _ in range(3))
scale = D ** -0.5
comm = torch.cuda.Stream(device=dev)
mode = os.environ.get("RING_MODE", "nccl")  # pull: CUDA IPC peer mappings + copy engines; nccl: send/recv rank -> rank+1
if mode == "pull":
    peers = fa.PeerShards(k, v)
    run = lambda: fa.ring_attention_pull(q, peers, scale, copy_stream=comm)
else:
    ring = fa.new_ring_group()
    run = lambda: fa.ring_attention(q, k, v, scale, group=ring, comm_stream=comm)
for _ in range(2):
    out = run()
dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 3
e0.record()
for _ in range(iters):
    out = run()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
# compute-only lower bound on this rank: the same P partial steps without any exchange
o_acc = torch.empty((B, H, n, D), dtype=torch.float32, device=dev)
m = torch.empty((B, H, n), dtype=torch.float32, device=dev)
l = torch.empty((B, H, n), dtype=torch.float32, device=dev)
of = torch.empty_like(q)
torch.cuda.synchronize()
e0.record()
for _ in range(iters):
    for s in range(world):
        fa.fa_forward_partial(q, k, v, o_acc, m, l, scale, s == 0, s == world - 1, of)
e1.record()
torch.cuda.synchronize()
ms_c = e0.elapsed_time(e1) / iters
if rank == 0:
    fl = 4.0 * B * H * N * N * D
    print(json.dumps({"config": "attention fwd B=1,H=32,N=65536,D=128 ring/context-parallel", "exchange": mode, "n_gpus": world,
                      "ms": float(ms.item()), "tflops_total": fl / float(ms.item()) / 1e9,
                      "ms_compute_only_rank0": ms_c, "exchange_bytes_per_gpu_per_step": 2 * B * H * n * D * 2}))
if mode == "pull":
    peers.close()
dist.destroy_process_group()
