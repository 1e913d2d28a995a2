"""Multi-GPU plumbing for the attention path.

* (b,h) sharding (BASELINE config 3): every (b,h) pair is an independent unit in every reference
  kernel (grid.y = BH, e.g. FlashAttention/flashattn_forward_wmma/flashattn_forward_wmma_v4.cu:452),
  so ranks take contiguous slices of the flat bh index and never communicate.
* ring / context-parallel attention (BASELINE config 4): N is split across ranks; K/V shards
  rotate rank -> rank+1 with point-to-point send/recv while `fa_forward_partial` folds each shard
  into the carried (O_acc, m, l) with the reference's own tile-to-tile merge rule
  (Streaming_FlashAttention_Forward_Kernel/flashattn_streaming_16x16.cu:247-261).

Only the ring uses a collective; `torch.distributed` is plumbing (NCCL on GPUs, gloo in the CPU
tests, which exercise the schedule with an injected step function).
"""
import torch
import torch.distributed as dist


def shard_range(total, world_size, rank):
    """Contiguous, balanced [begin, end) slice of `total` units for `rank` (first ranks get the
    remainder), e.g. GPU g of G takes bh in [g*256/G, (g+1)*256/G) for config 3."""
    if world_size <= 0 or not (0 <= rank < world_size) or total < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(total, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def ring_schedule(world_size, rank):
    """Owner rank of the K/V shard this rank attends to at each ring step."""
    return [(rank - s) % world_size for s in range(world_size)]


def new_ring_group(ranks=None):
    """NCCL group for the K/V exchange with a HIGH-PRIORITY internal stream.  The attention kernel occupies
    every SM (one 227 KB CTA each), so NCCL's send/recv kernel only gets scheduled while attention runs if its
    stream outranks the compute stream: measured on 2 x B200, a 512 MiB exchange issued next to a 15 ms
    attention launch takes 1.5 ms with the high-priority stream and 14-19 ms (i.e. after the kernel) without."""
    opts = dist.ProcessGroupNCCL.Options()
    opts.is_high_priority_stream = True
    return dist.new_group(ranks=ranks, backend="nccl", pg_options=opts)


def ring_attention(q, k, v, scale, group=None, step_fn=None, comm_stream=None):
    """Context-parallel attention over the ranks of `group`.

    q, k, v: this rank's [B, H, N/P, D] fp16 shards.  Returns this rank's [B, H, N/P, D] output.
    On GPUs pass a group made by `new_ring_group()` so the exchange overlaps the attention kernel.
    K/V for step s+1 are received into the alternate buffer while step s computes (double
    buffering; on GPUs the exchange runs on `comm_stream`).  `step_fn` defaults to the CUDA entry
    point `fa_forward_partial`; the gloo CPU tests inject a torch fp32 implementation of the same
    contract to test the schedule and the merge without a GPU.
    """
    if step_fn is None:
        from .ops import fa_forward_partial as step_fn
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B, H, Nq, D = q.shape
    o_acc = torch.empty((B, H, Nq, D), dtype=torch.float32, device=q.device)
    m = torch.empty((B, H, Nq), dtype=torch.float32, device=q.device)
    l = torch.empty((B, H, Nq), dtype=torch.float32, device=q.device)
    out = torch.empty_like(q)
    if world == 1:
        step_fn(q, k, v, o_acc, m, l, scale, True, True, out)
        return out

    send_to = dist.get_global_rank(group, (rank + 1) % world) if group is not None else (rank + 1) % world
    recv_from = dist.get_global_rank(group, (rank - 1) % world) if group is not None else (rank - 1) % world
    # one fused [2, ...] buffer per ring slot so K and V travel in a single message
    cur = torch.stack((k, v)).contiguous()
    nxt = torch.empty_like(cur)
    use_cuda = q.is_cuda
    compute_stream = torch.cuda.current_stream(q.device) if use_cuda else None
    if use_cuda and comm_stream is None:
        comm_stream = torch.cuda.Stream(device=q.device)

    for s in range(world):
        reqs = []
        if s + 1 < world:
            ops = [dist.P2POp(dist.isend, cur, send_to, group), dist.P2POp(dist.irecv, nxt, recv_from, group)]
            if use_cuda:
                comm_stream.wait_stream(compute_stream)  # `cur` was produced / last read on the compute stream
                with torch.cuda.stream(comm_stream):
                    reqs = dist.batch_isend_irecv(ops)
            else:
                reqs = dist.batch_isend_irecv(ops)
        step_fn(q, cur[0], cur[1], o_acc, m, l, scale, s == 0, s == world - 1, out)
        for r in reqs:
            r.wait()
        if use_cuda and s + 1 < world:
            compute_stream.wait_stream(comm_stream)
        cur, nxt = nxt, cur
    return out
