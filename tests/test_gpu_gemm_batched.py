"""GPU parity tests for gemm_batched (batched plain GEMM, optional B^T and scale) through the C ABI:
against the oracle's restatements of cpu_batched_gemm_ref / qk_cpu_ref, the committed golden vectors,
the reference's own WMMA kernels on the same buffers, and the fused-GEMM entry."""
import numpy as np
import pytest
import torch

from tests.oracle_lib import RefGpu, ref_available

pytestmark = pytest.mark.gpu

TOL_ABS, TOL_REL = 2e-3, 1e-3  # BASELINE.json north_star


def _rand_half(shape, seed):
    return np.random.default_rng(seed).uniform(-1, 1, shape).astype(np.float16)


def _run(fa, a, b, **kw):
    c = fa.gemm_batched(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), **kw)
    torch.cuda.synchronize()
    return c.cpu().numpy()


def test_golden_vectors_from_the_reference(fa, oracle, golden):
    g = golden("gemm_batched.npz")  # operands are fp16-representable by construction
    got = _run(fa, g["a"].astype(np.float16), g["b"].astype(np.float16))
    assert oracle.max_abs(got, g["c"]) <= TOL_ABS and oracle.rel_l2(got, g["c"]) <= TOL_REL
    g = golden("qk_scores.npz")
    got = _run(fa, g["q"].astype(np.float16), g["k"].astype(np.float16), trans_b=True, alpha=float(g["scale"]))
    assert oracle.max_abs(got, g["s"]) <= TOL_ABS and oracle.rel_l2(got, g["s"]) <= TOL_REL


# (batch, M, N, K): the reference harness shape (8 x 512^3, gemm_4_9_tensorcore_batched.cu:30-33), ragged M,
# N / K that are multiples of 8 but not of a tile, a single k-block, more tiles than CTA pairs
@pytest.mark.parametrize("shape", [(8, 512, 512, 512), (3, 100, 72, 40), (1, 16, 16, 16), (5, 257, 264, 136),
                                   (2, 1024, 2304, 64), (37, 300, 520, 128)])
def test_batched_parity_with_oracle(fa, oracle, shape):
    batch, M, N, K = shape
    a, b = _rand_half((batch, M, K), 1 + M), _rand_half((batch, K, N), 2 + N)
    got = _run(fa, a, b)
    nb = min(batch, 3)  # the fp32 triple loop is the slow side: check the first, second and last batch entries
    sel = sorted({0, 1 % batch, batch - 1})[:nb]
    ref = oracle.gemm_batched(a[sel].astype(np.float32), b[sel].astype(np.float32))
    assert oracle.max_abs(got[sel], ref) <= TOL_ABS * max(1.0, float(np.abs(ref).max()) / 8)
    assert oracle.rel_l2(got[sel], ref) <= TOL_REL


# (BH, N, D): the reference harness (4, 512, 64; qk_wmma_5_1_2.cu:125-128), config-2-like, D = 128, ragged
@pytest.mark.parametrize("shape", [(4, 512, 64), (2, 1024, 64), (3, 384, 128), (2, 200, 16), (1, 72, 8)])
def test_scaled_qk_scores_parity_with_oracle(fa, oracle, shape):
    BH, N, D = shape
    q, k = _rand_half((BH, N, D), 3 + N), _rand_half((BH, N, D), 4 + D)
    scale = 1.0 / np.sqrt(D)
    got = _run(fa, q, k, trans_b=True, alpha=scale)
    ref = oracle.qk_scores(q.astype(np.float32), k.astype(np.float32), scale)
    assert oracle.max_abs(got, ref) <= TOL_ABS and oracle.rel_l2(got, ref) <= TOL_REL


def test_rectangular_b_transposed(fa, oracle):
    """C = alpha * A * B^T with N != M (cross-attention scores): checked through the Nk != N oracle form."""
    q, k = _rand_half((2, 130, 64), 5), _rand_half((2, 333 + 3, 64), 6)  # N = 336 (multiple of 8)
    got = _run(fa, q, k, trans_b=True, alpha=0.125)
    ref = oracle.qk_scores(q.astype(np.float32), k.astype(np.float32), 0.125)
    assert oracle.max_abs(got, ref) <= TOL_ABS and oracle.rel_l2(got, ref) <= TOL_REL


def test_transposed_and_plain_forms_agree_bit_for_bit(fa):
    """A * B computed from B [K,N] (MN-major operand) and from B^T [N,K] (K-major operand) feed the tensor core the
    same products in the same order: the results must be identical, not just close."""
    a, b = _rand_half((4, 300, 192), 7), _rand_half((4, 192, 264), 8)
    plain = _run(fa, a, b)
    trans = _run(fa, a, np.ascontiguousarray(b.transpose(0, 2, 1)), trans_b=True)
    assert np.array_equal(plain, trans)


def test_batch_of_one_equals_the_fused_entry_without_epilogue_terms(fa):
    """gemm_fused_epilogue with no bias / residual is ReLU(A*B): equal to max(gemm_batched, 0) bit for bit."""
    a, b = _rand_half((1, 500, 256), 9), _rand_half((1, 256, 328), 10)
    plain = _run(fa, a, b)[0]
    fused = fa.gemm_fused_epilogue(torch.from_numpy(a[0]).cuda(), torch.from_numpy(b[0]).cuda()).cpu().numpy()
    assert np.array_equal(np.maximum(plain, 0.0), fused)


def test_linearity_in_alpha_and_batch_independence(fa):
    a, b = _rand_half((6, 256, 128), 11), _rand_half((6, 128, 256), 12)
    one = _run(fa, a, b)
    assert np.array_equal(_run(fa, a, b, alpha=0.5), one * 0.5)  # power-of-two scale is exact
    # permuting the batch permutes the result; a batch entry does not see its neighbours
    perm = np.array([3, 0, 5, 1, 4, 2])
    assert np.array_equal(_run(fa, a[perm], b[perm]), one[perm])
    assert np.array_equal(_run(fa, a[2:3], b[2:3])[0], one[2])


def test_large_batched_case_by_properties(fa, oracle):
    """256 x (1024 x 1024 x 128): config-2-sized score matrices for all (b,h) of config 3; sampled rows vs the oracle."""
    BH, N, D = 256, 1024, 128
    g = torch.Generator(device="cuda").manual_seed(5)
    q = (torch.rand((BH, N, D), device="cuda", generator=g) * 2 - 1).half()
    k = (torch.rand((BH, N, D), device="cuda", generator=g) * 2 - 1).half()
    s = fa.gemm_batched(q, k, trans_b=True, alpha=D ** -0.5)
    torch.cuda.synchronize()
    for bh in (0, 100, 255):
        rows = [0, 1, 127, 128, 511, 1023]
        ref = oracle.qk_scores(q[bh:bh + 1, rows].float().cpu().numpy(), k[bh:bh + 1].float().cpu().numpy(), D ** -0.5)
        got = s[bh, rows].cpu().numpy()[None]
        assert oracle.max_abs(got, ref) <= TOL_ABS and oracle.rel_l2(got, ref) <= TOL_REL
    # symmetry: (Q K^T)^T == K Q^T, computed as a second launch with the operands swapped
    st = fa.gemm_batched(k[:8], q[:8], trans_b=True, alpha=D ** -0.5)
    assert torch.equal(st.transpose(1, 2), s[:8])


@pytest.mark.skipif(not (ref_available("gemm_batched") and ref_available("qk_wmma")), reason="oracle/_ref not built")
def test_against_the_reference_wmma_kernels_on_this_gpu(fa, oracle):
    r = RefGpu()
    a = torch.from_numpy(_rand_half((8, 512, 512), 13)).cuda()
    b = torch.from_numpy(_rand_half((8, 512, 512), 14)).cuda()
    theirs = torch.zeros((8, 512, 512), device="cuda")
    r.gemm_batched(a, b, theirs)
    ours = fa.gemm_batched(a, b)
    torch.cuda.synchronize()
    assert oracle.max_abs(ours.cpu().numpy(), theirs.cpu().numpy()) <= TOL_ABS
    assert oracle.rel_l2(ours.cpu().numpy(), theirs.cpu().numpy()) <= TOL_REL
    q = torch.from_numpy(_rand_half((4, 512, 64), 15)).cuda()
    k = torch.from_numpy(_rand_half((4, 512, 64), 16)).cuda()
    theirs = torch.zeros((4, 512, 512), device="cuda")
    r.qk_wmma(q, k, theirs, 0.125)
    ours = fa.gemm_batched(q, k, trans_b=True, alpha=0.125)
    torch.cuda.synchronize()
    assert oracle.max_abs(ours.cpu().numpy(), theirs.cpu().numpy()) <= TOL_ABS
    assert oracle.rel_l2(ours.cpu().numpy(), theirs.cpu().numpy()) <= TOL_REL


def test_rejected_calls(fa):
    from flashattention_kernel_project_b200 import FaError
    a = torch.zeros((2, 16, 16), dtype=torch.float16, device="cuda")
    with pytest.raises(ValueError):
        fa.gemm_batched(a, torch.zeros((3, 16, 16), dtype=torch.float16, device="cuda"))
    with pytest.raises(ValueError):
        fa.gemm_batched(a, torch.zeros((2, 24, 16), dtype=torch.float16, device="cuda"))
    with pytest.raises(FaError) as e:  # N % 8 != 0: row pitch of C is not a multiple of 16 B
        fa.gemm_batched(a, torch.zeros((2, 16, 12), dtype=torch.float16, device="cuda"))
    assert e.value.code == -3


def test_many_tiny_batch_entries(fa, oracle):
    """4096 entries of 16 x 24 x 16: one 256 x 256 output tile per entry, 4096 tiles walked by 74 persistent CTA pairs."""
    a, b = _rand_half((4096, 16, 16), 21), _rand_half((4096, 16, 24), 22)
    got = _run(fa, a, b)
    ref = np.einsum("zmk,zkn->zmn", a.astype(np.float32), b.astype(np.float32))
    assert oracle.max_abs(got, ref) <= TOL_ABS and oracle.rel_l2(got, ref) <= TOL_REL
    sel = [0, 73, 74, 2047, 4095]
    assert oracle.max_abs(got[sel], oracle.gemm_batched(a[sel].astype(np.float32), b[sel].astype(np.float32))) <= TOL_ABS


def test_strided_operands_need_no_copy(fa, oracle):
    """gemm_batched_strided: leading dimensions and batch strides go into the TMA descriptors.  Views of larger
    tensors -- one head of a [B, N, H, D] activation, a padded leading dimension, a column block of C -- give the
    same numbers as their packed copies, bit for bit (the kernel sees the same tiles in the same order).  The
    reference hard-codes the packed strides (gemm_4_9_tensorcore_batched.cu:58-60)."""
    g = torch.Generator(device="cuda").manual_seed(3)
    B, N, H, D = 3, 200, 4, 64
    x = (torch.rand((B, N, H, D), device="cuda", generator=g) * 2 - 1).half()
    y = (torch.rand((B, N, H, D), device="cuda", generator=g) * 2 - 1).half()
    for h in range(H):
        q, k = x[:, :, h, :], y[:, :, h, :]  # [B, N, D] views: row pitch H*D, batch stride N*H*D
        assert not q.is_contiguous()
        got = fa.gemm_batched(q, k, trans_b=True, alpha=0.125)
        want = fa.gemm_batched(q.contiguous(), k.contiguous(), trans_b=True, alpha=0.125)
        assert torch.equal(got, want)
    ref = oracle.qk_scores(x[:, :, 1, :].float().cpu().numpy(), y[:, :, 1, :].float().cpu().numpy(), 0.125)
    got = fa.gemm_batched(x[:, :, 1, :], y[:, :, 1, :], trans_b=True, alpha=0.125).cpu().numpy()
    assert oracle.max_abs(got, ref) <= TOL_ABS and oracle.rel_l2(got, ref) <= TOL_REL
    # padded leading dimensions of A, B ([K, N] form) and C; C written into a column block of a wider buffer
    batch, M, Nn, K = 2, 136, 72, 40
    a_big = (torch.rand((batch, M, K + 24), device="cuda", generator=g) * 2 - 1).half()
    b_big = (torch.rand((batch, K + 8, Nn + 16), device="cuda", generator=g) * 2 - 1).half()
    c_big = torch.full((batch, M, Nn + 56), -7.0, device="cuda")
    a, b, c = a_big[:, :, :K], b_big[:, 8:, 8:8 + Nn], c_big[:, :, 32:32 + Nn]
    fa.gemm_batched(a, b, out=c)
    want = fa.gemm_batched(a.contiguous(), b.contiguous())
    torch.cuda.synchronize()
    assert torch.equal(c, want)
    assert (c_big[:, :, :32] == -7.0).all() and (c_big[:, :, 32 + Nn:] == -7.0).all()  # nothing written outside the block


def test_strided_entry_validates_pitches(fa):
    import ctypes
    lib = fa.load_library()
    a = torch.zeros((2, 64, 64), device="cuda", dtype=torch.float16)
    c = torch.zeros((2, 64, 64), device="cuda")
    p = lambda x: ctypes.c_void_p(x.data_ptr())
    call = lambda lda=0, ldb=0, ldc=0, sa=0, sb=0, sc=0: lib.gemm_batched_strided(
        p(a), p(a), p(c), 2, 64, 64, 64, lda, ldb, ldc, sa, sb, sc, 0, 1.0, None)
    assert call() == 0
    assert call(lda=32) == -7      # a leading dimension shorter than the row
    assert call(ldc=66) == -3      # fp32 rows must stay 16-byte aligned
    assert call(lda=68) == -3      # fp16 rows must stay 16-byte aligned
    assert call(sa=-64) == -7
    torch.cuda.synchronize()
