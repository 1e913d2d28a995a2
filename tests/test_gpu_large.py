"""Maximum-size and concurrency cases: BASELINE config 4's sequence length on one GPU, tensors with more than 2^31
elements (every index computation must be 64-bit), and concurrent calls from several host threads on several streams
(re-entrant, no global mutable state on the call path)."""
import threading

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL_ABS, TOL_REL = 2e-3, 1e-3  # BASELINE.json north_star


def _uniform_half(shape, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.rand(shape, device="cuda", generator=g) * 2 - 1).half()


def test_config4_sequence_length_on_one_gpu(fa, oracle):
    """N = 65 536, D = 128 (two of config 4's 32 heads): sampled rows against the oracle, the 8-shard partial chain (the
    ring's arithmetic without the exchange) against the single launch, and invariance under a joint permutation of K/V rows."""
    B, H, N, D = 1, 2, 65536, 128
    q, k, v = (_uniform_half((B, H, N, D), s) for s in (1, 2, 3))
    scale = D ** -0.5
    out = fa.fa_forward(q, k, v, scale)
    torch.cuda.synchronize()
    rows = np.array([0, 255, 256, 32767, 65535, N + 1, N + 40000, 2 * N - 1], dtype=np.int64)
    ref = oracle.attention_f16_rows(*(t.reshape(B * H, N, D).cpu().numpy() for t in (q, k, v)), scale, rows)
    got = out.reshape(B * H * N, D)[torch.from_numpy(rows).cuda()].float().cpu().numpy()
    assert oracle.rel_l2(got, ref) <= TOL_REL and oracle.max_abs(got, ref) <= TOL_ABS
    # ring arithmetic: 8 K/V shards of 8192 keys folded one after the other
    o_acc = torch.empty((B, H, N, D), device="cuda")
    m, l = torch.empty((B, H, N), device="cuda"), torch.empty((B, H, N), device="cuda")
    final = None
    for s in range(8):
        ks, vs = (t[:, :, s * 8192:(s + 1) * 8192].contiguous() for t in (k, v))
        final = fa.fa_forward_partial(q, ks, vs, o_acc, m, l, scale, s == 0, s == 7)
    torch.cuda.synchronize()
    a, b = final.float().cpu().numpy(), out.float().cpu().numpy()
    assert oracle.rel_l2(a, b) <= TOL_REL and oracle.max_abs(a, b) <= TOL_ABS
    # softmax attention is a set function of the (key, value) pairs
    perm = torch.randperm(N, device="cuda", generator=torch.Generator(device="cuda").manual_seed(9))
    outp = fa.fa_forward(q, k[:, :, perm].contiguous(), v[:, :, perm].contiguous(), scale)
    torch.cuda.synchronize()
    assert oracle.rel_l2(outp.float().cpu().numpy(), b) <= TOL_REL


def test_attention_with_more_than_2_pow_31_elements_per_tensor(fa, oracle):
    """B*H*N*D = 2.18e9 elements (4.4 GB per fp16 tensor, 8.7 GB of fp32 output): the last (b,h) must equal the same
    (b,h) run alone, bit for bit, and match the oracle on sampled rows."""
    BH, N, D = 520, 32768, 128
    assert BH * N * D > 2 ** 31
    q, k, v = (torch.empty((1, BH, N, D), dtype=torch.float16, device="cuda") for _ in range(3))
    g = torch.Generator(device="cuda").manual_seed(4)
    for t in (q, k, v):  # fill in slabs: torch.rand of the whole tensor would need a 17 GB fp32 temporary
        for s in range(0, BH, 40):
            t[0, s:s + 40] = (torch.rand((min(40, BH - s), N, D), device="cuda", generator=g) * 2 - 1).half()
    scale = D ** -0.5
    out = fa.fa_forward_f32out(q, k, v, scale)
    torch.cuda.synchronize()
    for bh in (0, 300, BH - 1):
        alone = fa.fa_forward_f32out(q[:, bh:bh + 1].contiguous(), k[:, bh:bh + 1].contiguous(), v[:, bh:bh + 1].contiguous(), scale)
        torch.cuda.synchronize()
        assert torch.equal(alone[0, 0], out[0, bh]), f"(b,h) {bh} differs from its stand-alone result"
    rows = np.array([0, N - 1, 17], dtype=np.int64)
    last = BH - 1
    ref = oracle.attention_f16_rows(q[0, last:last + 1].cpu().numpy(), k[0, last:last + 1].cpu().numpy(),
                                    v[0, last:last + 1].cpu().numpy(), scale, rows)
    got = out[0, last][torch.from_numpy(rows).cuda()].cpu().numpy()
    assert oracle.rel_l2(got, ref) <= TOL_REL and oracle.max_abs(got, ref) <= TOL_ABS
    del q, k, v, out
    torch.cuda.empty_cache()


def test_gemm_and_softmax_with_more_than_2_pow_31_output_elements(fa, oracle):
    M, N, K = 65536, 36864, 64  # 2.42e9 output elements; K = 64 keeps the arithmetic small
    assert M * N > 2 ** 31
    a, b = _uniform_half((M, K), 5), _uniform_half((K, N), 6)
    bias = torch.rand((N,), device="cuda")
    c = fa.gemm_fused_epilogue(a, b, bias, None)
    torch.cuda.synchronize()
    ref = oracle.gemm_fused_f16(a[-64:].cpu().numpy(), b.cpu().numpy(), bias.cpu().numpy(), None)
    got = c[-64:].cpu().numpy()
    assert oracle.max_abs(got, ref) <= TOL_ABS and oracle.rel_l2(got, ref) <= TOL_REL
    ref0 = oracle.gemm_fused_f16(a[:8].cpu().numpy(), b.cpu().numpy(), bias.cpu().numpy(), None)
    assert oracle.max_abs(c[:8].cpu().numpy(), ref0) <= TOL_ABS
    # in-place row softmax over the same 9.7 GB buffer: rows sum to 1, first and last rows against the oracle
    first, lastrows = c[:4].clone(), c[-4:].clone()
    fa.softmax_rows(c, out=c)
    torch.cuda.synchronize()
    sums = c.sum(dim=1)
    assert float((sums - 1).abs().max()) <= 1e-4
    # the oracle sums a row's 36 864 exps sequentially in fp32 (as the reference does) and drifts by up to ~2e-4 relative at
    # this row length; the kernel sums as a tree.  So: loose against the oracle, tight against a float64 evaluation.
    for got, x in ((c[:4].cpu().numpy(), first.cpu().numpy()), (c[-4:].cpu().numpy(), lastrows.cpu().numpy())):
        assert np.allclose(got, oracle.softmax_rows(x), rtol=1e-3, atol=1e-10)
        e = np.exp(x.astype(np.float64) - x.astype(np.float64).max(axis=1, keepdims=True))
        assert np.allclose(got, e / e.sum(axis=1, keepdims=True), rtol=2e-5, atol=1e-12)
    del c
    torch.cuda.empty_cache()


def test_concurrent_calls_from_host_threads_on_their_own_streams(fa):
    """Four host threads, each with its own stream and its own shapes, interleave attention / GEMM / softmax calls; every
    result must equal the one computed serially (the library keeps no per-call state outside the caller's buffers)."""
    shapes = [(2, 4, 700, 128), (1, 8, 1024, 64), (3, 2, 333, 128), (1, 1, 4096, 64)]
    inputs = [tuple(_uniform_half(s, 10 + 3 * i + j) for j in range(3)) for i, s in enumerate(shapes)]
    ga, gb = _uniform_half((1000, 512), 30), _uniform_half((512, 1032), 31)
    sx = torch.randn((3000, 1000), device="cuda", generator=torch.Generator(device="cuda").manual_seed(32))
    serial = [fa.fa_forward(*qkv, qkv[0].shape[-1] ** -0.5) for qkv in inputs]
    serial_c, serial_p = fa.gemm_fused_epilogue(ga, gb), fa.softmax_rows(sx)
    torch.cuda.synchronize()
    results, errors = [None] * len(shapes), []

    def worker(i):
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                for _ in range(20):
                    o = fa.fa_forward(*inputs[i], inputs[i][0].shape[-1] ** -0.5)
                    c = fa.gemm_fused_epilogue(ga, gb)
                    p = fa.softmax_rows(sx)
                st.synchronize()
            results[i] = (o, c, p)
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(shapes))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for i, (o, c, p) in enumerate(results):
        assert torch.equal(o, serial[i]) and torch.equal(c, serial_c) and torch.equal(p, serial_p)
