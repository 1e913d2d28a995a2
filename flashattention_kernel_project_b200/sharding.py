"""Multi-GPU plumbing for the attention path.

* (b,h) sharding (BASELINE config 3): every (b,h) pair is an independent unit in every reference
  kernel (grid.y = BH, e.g. FlashAttention/flashattn_forward_wmma/flashattn_forward_wmma_v4.cu:452),
  so ranks take contiguous slices of the flat bh index and never communicate.
* ring / context-parallel attention (BASELINE config 4): N is split across ranks; K/V shards
  rotate rank -> rank+1 with point-to-point send/recv while `fa_forward_partial` folds each shard
  into the carried (O_acc, m, l) with the reference's own tile-to-tile merge rule
  (Streaming_FlashAttention_Forward_Kernel/flashattn_streaming_16x16.cu:247-261).

  `ring_forward` is the product path: the whole loop runs inside the library (C ABI `fa_ring_forward`: the
  partial-attention launches and grouped ncclSend / ncclRecv on a high-priority stream, tied together with events
  from C++; every rank sends its own read-only shard straight to the rank that needs it next, so all exchanges
  are in flight from the start and none waits for an earlier one; or, with `peers=`, `fa_ring_forward_pull`: no
  NCCL at all, the attention kernel's spare warps load the next shard from its owner's memory over NVLink).
  The Python-orchestrated `ring_attention` below is kept as the independent implementation it is tested against:
  it passes shards rank -> rank+1 with NCCL send/recv.  (Round 1's copy-engine variant is gone: it copied from
  mappings opened on the OWNER's device index -- what torch's tensor sharing produces -- which the driver stages at
  25 GB/s; opened in the consumer's context the same copy runs at 725 GB/s, and kernels can load from the mapping.)

Only the ring communicates; `torch.distributed` is plumbing (NCCL on GPUs, gloo in the CPU tests, which
exercise the schedule with an injected step function).
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


def _nccl_comm_ptr(group, device):
    """The ncclComm_t behind `group` on `device` (created by a first collective if the group is still lazy)."""
    pg = group if group is not None else dist.group.WORLD
    backend = pg._get_backend(torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device()))
    ptr = backend._comm_ptr()
    if not ptr:
        dist.barrier(group=group, device_ids=[device.index])
        ptr = backend._comm_ptr()
    if not ptr:
        raise RuntimeError("the process group has no NCCL communicator on this device")
    return int(ptr)


class RingWorkspace:
    """Device memory for `ring_forward`: the carried fp32 (O_acc, m, l) and `depth` receive buffers of one K and one
    V shard (depth = 0: world - 1 buffers, every exchange prefetched).  Reusable across calls of the same shape."""

    def __init__(self, shape, world, device, depth=0):
        from ._lib import load_library
        B, H, n, D = (int(x) for x in shape)
        self.shape, self.world, self.depth = (B, H, n, D), int(world), int(depth)
        self.nbytes = int(load_library().fa_ring_workspace_bytes(B, H, n, D, self.world, self.depth))
        if self.nbytes == 0:
            raise ValueError("bad ring shape")
        self.buffer = torch.empty(self.nbytes, dtype=torch.uint8, device=device)
        if self.buffer.data_ptr() % 256:
            raise RuntimeError("allocator returned memory that is not 256-byte aligned")


def ring_forward(q, k, v, scale=None, group=None, out=None, workspace=None, depth=0, stream=None, peers=None):
    """Context-parallel attention over the ranks of `group` through the C ABI (one call per rank, same shapes
    everywhere).  q, k, v: this rank's [B, H, N/P, D] fp16 shards; returns this rank's rows of O.
    Nothing is synchronised on the host: the call only enqueues on `stream` and on the library's exchange stream.

    peers=None: `fa_ring_forward`, the exchange is grouped ncclSend / ncclRecv on the group's communicator.
    peers=PeerShards(k, v): `fa_ring_forward_pull`, the exchange is fused into the attention kernel -- the two spare
    warps of every CTA load the next step's shard from its owner through the peer mappings (no extra launches, no
    SMs taken, no rank-to-rank synchronisation inside the call; the caller keeps K / V unchanged between the barriers
    PeerShards places at construction and close())."""
    import ctypes
    from ._lib import check, load_library
    if not (q.is_cuda and k.is_cuda and v.is_cuda):
        raise TypeError("ring_forward needs CUDA tensors (no CPU fallback exists)")
    if q.dtype != torch.float16 or k.dtype != torch.float16 or v.dtype != torch.float16:
        raise TypeError("Q, K, V must be float16")
    if q.dim() != 4 or k.shape != q.shape or v.shape != q.shape:
        raise ValueError("Q, K, V must be [B, H, N/P, D] shards of one shape")
    if not (q.is_contiguous() and k.is_contiguous() and v.is_contiguous()):
        raise ValueError("Q, K, V must be contiguous")
    B, H, n, D = (int(x) for x in q.shape)
    if scale is None:
        scale = 1.0 / float(D) ** 0.5
    initialised = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if initialised else 1
    rank = dist.get_rank(group) if initialised else 0
    if out is None:
        out = torch.empty_like(q)
    elif out.shape != q.shape or out.dtype != torch.float16 or out.device != q.device or not out.is_contiguous():
        raise ValueError("out must be a contiguous float16 tensor of Q's shape on Q's device")
    if workspace is None:
        workspace = RingWorkspace(q.shape, world, q.device, depth)
    elif workspace.shape != (B, H, n, D) or workspace.world != world or workspace.buffer.device != q.device:
        raise ValueError("workspace was made for another shape, world size or device")
    s = torch.cuda.current_stream(q.device) if stream is None else stream
    if peers is not None and world > 1:
        if peers.world != world or peers.rank != rank:
            raise ValueError("peers was made for another group")
        pk = (ctypes.c_void_p * world)(*[ctypes.c_void_p(peers.ptrs(r)[0]) for r in range(world)])
        pv = (ctypes.c_void_p * world)(*[ctypes.c_void_p(peers.ptrs(r)[1]) for r in range(world)])
        with torch.cuda.device(q.device):
            check(load_library().fa_ring_forward_pull(
                ctypes.c_void_p(q.data_ptr()), ctypes.c_void_p(k.data_ptr()), ctypes.c_void_p(v.data_ptr()),
                ctypes.c_void_p(out.data_ptr()), B, H, n, D, float(scale), rank, world, pk, pv,
                ctypes.c_void_p(workspace.buffer.data_ptr()), ctypes.c_size_t(workspace.nbytes),
                ctypes.c_void_p(s.cuda_stream)), "fa_ring_forward_pull")
        return out
    comm = _nccl_comm_ptr(group, q.device) if world > 1 else None
    with torch.cuda.device(q.device):
        check(load_library().fa_ring_forward(
            ctypes.c_void_p(q.data_ptr()), ctypes.c_void_p(k.data_ptr()), ctypes.c_void_p(v.data_ptr()),
            ctypes.c_void_p(out.data_ptr()), B, H, n, D, float(scale), rank, world,
            ctypes.c_void_p(comm) if comm else None, ctypes.c_void_p(workspace.buffer.data_ptr()),
            ctypes.c_size_t(workspace.nbytes), ctypes.c_void_p(s.cuda_stream)), "fa_ring_forward")
    return out


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


class PeerShards:
    """Every rank's K/V shard, mapped into this process over NVLink / NVSwitch (CUDA IPC through the library's
    fa_ipc_* entry points, opened in THIS GPU's context so that its kernels can load from them directly).

    Collective over `group`: all ranks construct it together with their own [B, H, N/P, D] fp16 shards, which must
    stay allocated and unmodified until `close()` (also collective).  `ptrs(r)` returns the (K, V) addresses of rank
    r's shard as seen from this process (its own tensors' addresses for r == rank)."""

    def __init__(self, k, v, group=None):
        import ctypes
        from ._lib import check, load_library
        if not (k.is_cuda and v.is_cuda and k.is_contiguous() and v.is_contiguous()):
            raise ValueError("K and V shards must be contiguous CUDA tensors")
        lib = load_library()
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._local = (k, v)
        self._device = k.device

        def export(t):
            handle = (ctypes.c_ubyte * 64)()
            offset = ctypes.c_ulonglong(0)
            check(lib.fa_ipc_export(ctypes.c_void_p(t.data_ptr()), handle, ctypes.byref(offset)), "fa_ipc_export")
            return bytes(handle), int(offset.value)
        # whatever produced k and v on this rank has finished before any peer can see the handles ...
        torch.cuda.synchronize(k.device)
        with torch.cuda.device(k.device):
            mine = (export(k), export(v))
        table = [None] * self.world
        dist.all_gather_object(table, mine, group=group)
        self._bases = {}  # handle bytes -> mapped base address (one mapping per allocation and process)
        self._ptrs = []
        with torch.cuda.device(k.device):
            for r, entry in enumerate(table):
                if r == self.rank:
                    self._ptrs.append((k.data_ptr(), v.data_ptr()))
                    continue
                pair = []
                for handle, offset in entry:
                    if handle not in self._bases:
                        base = ctypes.c_void_p()
                        buf = (ctypes.c_ubyte * 64).from_buffer_copy(handle)
                        check(lib.fa_ipc_open(buf, ctypes.byref(base)), "fa_ipc_open")
                        self._bases[handle] = int(base.value)
                    pair.append(self._bases[handle] + offset)
                self._ptrs.append(tuple(pair))
        # ... and no rank starts reading before every rank has mapped every shard (mirrors close())
        dist.barrier(group=group)

    def ptrs(self, r):
        return self._ptrs[r]

    def close(self):
        """No rank may free or overwrite its shard while another still reads it."""
        import ctypes
        from ._lib import check, load_library
        torch.cuda.synchronize(self._device)
        dist.barrier(group=self.group)
        with torch.cuda.device(self._device):
            for base in self._bases.values():
                check(load_library().fa_ipc_close(ctypes.c_void_p(base)), "fa_ipc_close")
        self._bases, self._ptrs = {}, []
        dist.barrier(group=self.group)
