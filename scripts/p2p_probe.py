"""Developer probe (torchrun, 2+ ranks): how fast does a K/V shard travel between two B200s, alone and while the
attention kernel holds every SM?  (a) NCCL send/recv (SM kernels), (b) cudaMemcpyAsync issued on THIS GPU's stream from
a peer-mapped (CUDA IPC) buffer with peer access enabled (copy engines), then BASELINE config 4 as a ring both ways."""
import ctypes
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
cudart = ctypes.CDLL("libcudart.so.12")
cudart.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
x = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
y = torch.empty_like(x)
peer = (rank + 1) % world
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def xchg():
    ops = [dist.P2POp(dist.isend, x, peer), dist.P2POp(dist.irecv, y, (rank - 1) % world)]
    for r in dist.batch_isend_irecv(ops):
        r.wait()


for _ in range(3):
    xchg()
torch.cuda.synchronize()
dist.barrier()
e0.record()
for _ in range(5):
    xchg()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
if rank == 0:
    print(f"NCCL send/recv 256 MiB each way, alone: {ms:.3f} ms -> {x.numel() / ms / 1e6:.0f} GB/s per direction", flush=True)
peers = fa.PeerShards(x, x)
src_ptr = peers.ptrs((rank - 1) % world)[0]
side = torch.cuda.Stream(device=dev)


def pull(stream):
    rc = cudart.cudaMemcpyAsync(y.data_ptr(), src_ptr, x.numel(), 3, stream.cuda_stream)
    assert rc == 0, rc


for _ in range(3):
    pull(torch.cuda.current_stream())
torch.cuda.synchronize()
dist.barrier()
e0.record()
for _ in range(5):
    pull(torch.cuda.current_stream())
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
if rank == 0:
    print(f"copy-engine pull 256 MiB from the peer mapping, alone: {ms:.3f} ms -> {x.numel() / ms / 1e6:.0f} GB/s", flush=True)
q, k, v = ((torch.rand((1, 32, 16384, 128), device=dev) * 2 - 1).half() for _ in range(3))
o = torch.empty_like(q)
fa.fa_forward(q, k, v, 0.088, out=o)
torch.cuda.synchronize()
dist.barrier()
c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
fa.fa_forward(q, k, v, 0.088, out=o)
with torch.cuda.stream(side):
    c0.record()
    pull(side)
    c1.record()
e1.record()
torch.cuda.synchronize()
if rank == 0:
    print(f"attention kernel {e0.elapsed_time(e1):.2f} ms; pull issued next to it took {c0.elapsed_time(c1):.3f} ms "
          f"-> {x.numel() / c0.elapsed_time(c1) / 1e6:.0f} GB/s", flush=True)
dist.barrier()
peers.close()
del peers, x, y, q, k, v, o
# BASELINE config 4 as a ring, both exchanges
B, H, N, D = 1, 32, 65536, 128
n = N // world
q, k, v = ((torch.rand((B, H, n, D), device=dev) * 2 - 1).half() for _ in range(3))
ws = fa.RingWorkspace(q.shape, world, dev)
out = torch.empty_like(q)
peers = fa.PeerShards(k, v)
for name, kw in (("nccl", {}), ("pull", {"peers": peers})):
    run = lambda: fa.ring_forward(q, k, v, D ** -0.5, out=out, workspace=ws, **kw)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    dist.barrier()
    e0.record()
    for _ in range(5):
        run()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 5], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"config 4 ring over {world} GPUs, exchange = {name}: {t.item():.3f} ms", flush=True)
    dist.barrier()
peers.close()
dist.destroy_process_group()
