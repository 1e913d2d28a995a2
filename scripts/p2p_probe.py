"""Developer probe: NCCL send/recv bandwidth between 2 ranks, alone and while the attention kernel runs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import flashattention_kernel_project_b200 as fa
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
opts = dist.ProcessGroupNCCL.Options()
opts.is_high_priority_stream = os.environ.get("COMM_PRIO", "0") != "0"
dist.init_process_group("nccl", device_id=dev, pg_options=opts)
x = torch.empty(512 << 20, dtype=torch.uint8, device=dev); y = torch.empty_like(x)
peer = 1 - rank
def xchg():
    ops = [dist.P2POp(dist.isend, x, peer), dist.P2POp(dist.irecv, y, peer)]
    for r in dist.batch_isend_irecv(ops): r.wait()
for _ in range(3): xchg()
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); [xchg() for _ in range(5)]; e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
if rank == 0: print(f"send/recv 512 MiB each way alone: {ms:.2f} ms -> {x.numel() / ms / 1e6:.0f} GB/s per direction")
q, k, v = ((torch.rand((1, 32, 32768, 128), device=dev) * 2 - 1).half() for _ in range(3))
o = torch.empty_like(q)
comm = torch.cuda.Stream(device=dev, priority=int(os.environ.get('COMM_PRIO', '0')))
torch.cuda.synchronize(); dist.barrier()
e0.record()
fa.fa_forward(q, k, v, 0.088, out=o)
with torch.cuda.stream(comm):
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(); xchg(); c1.record()
e1.record(); torch.cuda.synchronize()
if rank == 0: print(f"attention kernel {e0.elapsed_time(e1):.2f} ms (incl. wait), concurrent exchange took {c0.elapsed_time(c1):.2f} ms")
dist.destroy_process_group()
