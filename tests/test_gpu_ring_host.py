"""GPU tests for fa_forward_partial (the ring / context-parallel step, chained on one GPU exactly as
the ranks of a ring would see the shards) and for the host-buffer entry points."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(B, H, N, D, seed, std=1.0):
    rng = np.random.default_rng(seed)
    return tuple((rng.standard_normal((B, H, N, D)) * std).astype(np.float16) for _ in range(3))


@pytest.mark.parametrize("D", [64, 128])
@pytest.mark.parametrize("chunks", [[(0, 512)], [(0, 256), (256, 512)], [(0, 100), (100, 228), (228, 229), (229, 512)],
                                    [(256, 512), (0, 256)]])
def test_partial_chain_equals_full_attention(fa, oracle, D, chunks):
    B, H, N = 1, 3, 512
    q, k, v = _inputs(B, H, N, D, seed=D + len(chunks), std=1.5)
    scale = 1.0 / np.sqrt(D)
    tq, tk, tv = (torch.from_numpy(x).cuda() for x in (q, k, v))
    nq = 300  # this rank's query shard is ragged on purpose
    tq = tq[:, :, :nq].contiguous()
    o_acc = torch.full((B, H, nq, D), float("nan"), device="cuda")  # is_first must ignore the incoming state
    m = torch.full((B, H, nq), float("nan"), device="cuda")
    l = torch.full((B, H, nq), float("nan"), device="cuda")
    out = None
    for i, (b, e) in enumerate(chunks):
        out = fa.fa_forward_partial(tq, tk[:, :, b:e].contiguous(), tv[:, :, b:e].contiguous(), o_acc, m, l, scale,
                                    i == 0, i == len(chunks) - 1)
    torch.cuda.synchronize()
    ref = oracle.attention_f16(q[:, :, :nq].reshape(B * H, nq, D), k.reshape(B * H, N, D), v.reshape(B * H, N, D), scale)
    got = out.float().cpu().numpy().reshape(B * H, nq, D)
    # N(0, 1.5) inputs give |O| up to ~4: the absolute bound scales with it (see tests/test_gpu_attention._check)
    assert oracle.rel_l2(got, ref) <= 1e-3 and oracle.max_abs(got, ref) <= 2e-3 * max(1.0, float(np.abs(ref).max()))


def test_partial_state_is_the_online_softmax_state(fa, oracle):
    """After a non-final step, (O_acc / l, m) are the exact attention / row max over the shards seen."""
    B, H, N, D = 1, 2, 384, 64
    q, k, v = _inputs(B, H, N, D, seed=1)
    scale = 0.125
    tq, tk, tv = (torch.from_numpy(x).cuda() for x in (q, k, v))
    o_acc = torch.empty((B, H, N, D), device="cuda")
    m = torch.empty((B, H, N), device="cuda")
    l = torch.empty((B, H, N), device="cuda")
    fa.fa_forward_partial(tq, tk[:, :, :200].contiguous(), tv[:, :, :200].contiguous(), o_acc, m, l, scale, True, False)
    torch.cuda.synchronize()
    ref = oracle.attention_f16(q.reshape(B * H, N, D), k[:, :, :200].reshape(B * H, 200, D),
                               v[:, :, :200].reshape(B * H, 200, D), scale)
    got = (o_acc / l[..., None]).cpu().numpy().reshape(B * H, N, D)
    assert oracle.rel_l2(got, ref) <= 1e-3
    s = np.einsum("bhqd,bhkd->bhqk", q.astype(np.float32), k[:, :, :200].astype(np.float32)) * scale
    # m may lag the true max by the lazy-rescale slack (2^8 in the exp2 domain = 8 ln 2 natural units), never exceed it
    gap = s.max(axis=-1) - m.cpu().numpy()
    assert (gap >= -1e-3).all() and (gap <= 8 * np.log(2) + 1e-3).all()


def test_host_buffer_entry_points(fa, oracle):
    """The reference main() flow (H2D -> kernel -> D2H) behind one call, on pinned and pageable memory."""
    B, H, N, D = 2, 2, 300, 128
    q, k, v = _inputs(B, H, N, D, seed=4)
    scale = 1.0 / np.sqrt(D)
    ref = oracle.attention_f16(*(x.reshape(B * H, N, D) for x in (q, k, v)), scale)
    out = np.empty_like(q)
    fa.fa_forward_host(q, k, v, out, scale)
    assert oracle.rel_l2(out.astype(np.float32).reshape(B * H, N, D), ref) <= 1e-3
    pq, pk, pv = (torch.from_numpy(x).pin_memory() for x in (q, k, v))
    pout = torch.empty_like(pq).pin_memory()
    fa.fa_forward_host(pq, pk, pv, pout, scale)
    assert np.array_equal(pout.numpy(), out)

    rng = np.random.default_rng(0)
    a = rng.uniform(-1, 1, (256, 128)).astype(np.float16)
    b = rng.uniform(-1, 1, (128, 512)).astype(np.float16)
    bias = rng.uniform(-1, 1, (512,)).astype(np.float32)
    res = rng.uniform(-1, 1, (256, 512)).astype(np.float32)
    c = np.empty((256, 512), np.float32)
    fa.gemm_fused_epilogue_host(a, b, bias, res, c)
    assert oracle.max_abs(c, oracle.gemm_fused_f16(a, b, bias, res)) <= 2e-3
    fa.gemm_fused_epilogue_host(a, b, None, None, c)
    assert oracle.max_abs(c, oracle.gemm_fused_f16(a, b, None, None)) <= 2e-3

    x = (rng.standard_normal((100, 777)) * 3).astype(np.float32)
    p = np.empty_like(x)
    fa.softmax_rows_host(x, p)
    assert oracle.max_abs(p, oracle.softmax_rows(x)) <= 1e-6
    assert fa.launch_count() >= 5


def test_host_buffer_calls_are_chunk_pipelined_and_identical_to_the_device_path(fa):
    """Sizes large enough that the host-buffer entry points split the batch into several H2D / kernel / D2H chunks:
    the result must be bit-identical to one device-resident call (chunks are independent units)."""
    B, H, N, D = 2, 24, 2048, 128  # 25 MB per tensor -> several chunks; 48 (b,h) do not divide evenly by every count
    g = torch.Generator().manual_seed(3)
    q, k, v = ((torch.rand((B, H, N, D), generator=g) * 2 - 1).half().pin_memory() for _ in range(3))
    out = torch.empty_like(q).pin_memory()
    scale = D ** -0.5
    fa.fa_forward_host(q, k, v, out, scale)
    dev = fa.fa_forward(q.cuda(), k.cuda(), v.cuda(), scale)
    torch.cuda.synchronize()
    assert torch.equal(out, dev.cpu())
    # second call reuses the cached scratch and streams
    out2 = torch.empty_like(out)
    fa.fa_forward_host(q, k, v, out2, scale)
    assert torch.equal(out2, out)

    a = (torch.rand((2000, 512), generator=g) * 2 - 1).half()
    b = (torch.rand((512, 1536), generator=g) * 2 - 1).half()
    bias, res = torch.rand((1536,), generator=g), torch.rand((2000, 1536), generator=g)
    c = torch.empty((2000, 1536))
    fa.gemm_fused_epilogue_host(a.numpy(), b.numpy(), bias.numpy(), res.numpy(), c.numpy())
    cdev = fa.gemm_fused_epilogue(a.cuda(), b.cuda(), bias.cuda(), res.cuda())
    assert torch.equal(c, cdev.cpu())

    x = torch.randn((5000, 2048), generator=g) * 3
    p = torch.empty_like(x)
    fa.softmax_rows_host(x.numpy(), p.numpy())
    assert torch.equal(p, fa.softmax_rows(x.cuda()).cpu())


def test_plain_c_program_through_the_c_abi(tmp_path):
    """examples/drop_in.c: a gcc-built C99 program (no CUDA headers, no Python) calls fa_forward_host and checks a row."""
    from tests.test_abi import build_and_run_c_example
    r = build_and_run_c_example(tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "max abs error" in r.stdout


@pytest.mark.parametrize("world,shape", [(2, (1, 4, 1024, 128)), (4, (2, 3, 1024, 64)), (8, (1, 2, 4096, 128)), (3, (1, 2, 768, 64))])
def test_fused_ring_with_every_rank_driven_on_one_gpu(fa, oracle, world, shape):
    """fa_ring_forward_pull on ONE GPU: every rank's call is made in turn by this process, the "peer" pointers being the
    other ranks' shards on the same device (the header allows plain pointers when one process drives several GPUs;
    a call never waits for another rank).  This runs the complete product path of the context-parallel ring -- P
    launches per rank, the spare warps of every CTA loading the next step's K / V shard, the carried fp32 state through
    TMA, two alternating receive buffers -- where only one GPU is available, against the oracle on the whole sequence."""
    import ctypes
    B, H, N, D = shape
    n = N // world
    q, k, v = _inputs(B, H, N, D, seed=world + N, std=1.2)
    scale = 1.0 / np.sqrt(D)
    lib = fa.load_library()
    shards = lambda x: [torch.from_numpy(np.ascontiguousarray(x[:, :, r * n:(r + 1) * n])).cuda() for r in range(world)]
    qs, ks, vs = shards(q), shards(k), shards(v)
    nbytes = int(lib.fa_ring_workspace_bytes(B, H, n, D, world, 2))
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    peer_k = (ctypes.c_void_p * world)(*[ctypes.c_void_p(t.data_ptr()) for t in ks])
    peer_v = (ctypes.c_void_p * world)(*[ctypes.c_void_p(t.data_ptr()) for t in vs])
    outs = []
    for r in range(world):
        out = torch.empty_like(qs[r])
        ws.fill_(0xFF)  # stale workspace contents (NaN patterns) must not leak into the result
        rc = lib.fa_ring_forward_pull(ctypes.c_void_p(qs[r].data_ptr()), ctypes.c_void_p(ks[r].data_ptr()),
                                      ctypes.c_void_p(vs[r].data_ptr()), ctypes.c_void_p(out.data_ptr()), B, H, n, D,
                                      ctypes.c_float(scale), r, world, peer_k, peer_v, ctypes.c_void_p(ws.data_ptr()),
                                      ctypes.c_size_t(nbytes), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, rc
        outs.append(out)
    torch.cuda.synchronize()
    got = torch.cat(outs, dim=2).float().cpu().numpy().reshape(B * H, N, D)
    ref = oracle.attention_f16(q.reshape(B * H, N, D), k.reshape(B * H, N, D), v.reshape(B * H, N, D), scale)
    assert oracle.rel_l2(got, ref) <= 1e-3 and oracle.max_abs(got, ref) <= 2e-3 * max(1.0, float(np.abs(ref).max()))
    # the same steps through fa_forward_partial, no prefetch: bit-identical (same kernel calls, same shard order)
    r = world - 1
    o_acc = torch.empty((B, H, n, D), device="cuda")
    m, l = torch.empty((B, H, n), device="cuda"), torch.empty((B, H, n), device="cuda")
    for x in range(world):
        src = (r - x) % world
        plain = fa.fa_forward_partial(qs[r], ks[src], vs[src], o_acc, m, l, scale, x == 0, x == world - 1)
    torch.cuda.synchronize()
    assert torch.equal(plain, outs[r])
