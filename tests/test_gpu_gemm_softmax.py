"""GPU parity tests for gemm_fused_epilogue and softmax_rows through the C ABI."""
import numpy as np
import pytest
import torch

from tests.oracle_lib import RefGpu, ref_available

pytestmark = pytest.mark.gpu


def _gemm_inputs(M, N, K, seed=42):
    rng = np.random.default_rng(seed)
    a = rng.uniform(-1, 1, (M, K)).astype(np.float16)
    b = rng.uniform(-1, 1, (K, N)).astype(np.float16)
    return a, b, rng.uniform(-1, 1, (N,)).astype(np.float32), rng.uniform(-1, 1, (M, N)).astype(np.float32)


def _gemm(fa, a, b, bias, res):
    c = fa.gemm_fused_epilogue(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(),
                               None if bias is None else torch.from_numpy(bias).cuda(),
                               None if res is None else torch.from_numpy(res).cuda())
    torch.cuda.synchronize()
    return c.cpu().numpy()


def test_gemm_golden_vector_from_the_reference(fa, oracle, golden):
    """mt19937(42) fixture of tensorcore_gemm_fused_epilogue.cu:178-184 through cpu_gemm_fused_ref.  The
    reference feeds its CPU check UNROUNDED fp32 A, B while the GPU gets fp16 copies (:186-195), so this
    comparison carries input quantisation (~sqrt(K) * 2^-11); the tight check is the next test."""
    g = golden("gemm_fused.npz")
    got = _gemm(fa, g["a"].astype(np.float16), g["b"].astype(np.float16), g["bias"], g["residual"])
    assert oracle.max_abs(got, g["c"]) <= 2e-2
    tight = oracle.gemm_fused_f16(g["a"].astype(np.float16), g["b"].astype(np.float16), g["bias"], g["residual"])
    assert oracle.max_abs(got, tight) <= 2e-3 and oracle.rel_l2(got, tight) <= 1e-3


_GEMM_SHAPES = [(128, 256, 64), (16, 16, 16), (1024, 1024, 1024), (1000, 1096, 520), (129, 264, 72), (4096, 512, 2048)]
# nullable-operand combinations (tensorcore_gemm_fused_epilogue.cu:112,120) on the small shapes, full epilogue everywhere
_GEMM_CASES = [(s, b, r) for s in _GEMM_SHAPES for (b, r) in [(True, True), (False, False), (True, False), (False, True)]
               if (b and r) or (not b and not r) or s[0] <= 200]


@pytest.mark.parametrize("shape,with_bias,with_res", _GEMM_CASES)
def test_gemm_parity_with_oracle(fa, oracle, shape, with_bias, with_res):
    M, N, K = shape
    a, b, bias, res = _gemm_inputs(M, N, K, seed=M + K)
    bias, res = (bias if with_bias else None), (res if with_res else None)
    got = _gemm(fa, a, b, bias, res)
    rows = (0, M) if M <= 1024 else (M - 200, M - 72)  # sample a row block of the big case
    ref = oracle.gemm_fused_f16(a, b, bias, res, row_begin=rows[0], row_end=rows[1])
    sub = got[rows[0]:rows[1]]
    assert oracle.max_abs(sub, ref) <= 2e-3 and oracle.rel_l2(sub, ref) <= 1e-3


def test_gemm_config5a_sampled_rows_and_properties(fa, oracle):
    """BASELINE config 5a (M=N=K=8192): 64 sampled rows through the oracle + structural properties."""
    M = N = K = 8192
    g = torch.Generator(device="cuda").manual_seed(42)
    a = (torch.rand((M, K), device="cuda", generator=g) * 2 - 1).half()
    b = (torch.rand((K, N), device="cuda", generator=g) * 2 - 1).half()
    bias = torch.rand((N,), device="cuda", generator=g) * 2 - 1
    res = torch.rand((M, N), device="cuda", generator=g) * 2 - 1
    c = fa.gemm_fused_epilogue(a, b, bias, res)
    torch.cuda.synchronize()
    r0 = 4000
    ref = oracle.gemm_fused_f16(a.cpu().numpy(), b.cpu().numpy(), bias.cpu().numpy(), res.cpu().numpy(),
                                row_begin=r0, row_end=r0 + 64)
    got = c[r0:r0 + 64].cpu().numpy()
    # fp32 accumulation order differs (tensor-core tree vs sequential): error ~ sqrt(K) ulp of |acc| ~ 50
    assert oracle.rel_l2(got, ref) <= 1e-3 and oracle.max_abs(got, ref) <= 2e-3
    # epilogue identity: C - residual = ReLU(...) >= 0, and with residual = None it is exactly that
    c0 = fa.gemm_fused_epilogue(a, b, bias, None)
    assert (c0 >= 0).all()
    assert torch.equal(c0 + res, c)
    # zero A => C = ReLU(bias) + residual exactly
    cz = fa.gemm_fused_epilogue(torch.zeros_like(a[:256]), b, bias, res[:256].contiguous())
    assert torch.equal(cz, torch.relu(bias)[None, :] + res[:256])


@pytest.mark.skipif(not ref_available("gemm_fused"), reason="oracle/_ref not built")
def test_gemm_against_the_reference_kernel_on_the_same_gpu(fa, oracle):
    a, b, bias, res = (torch.from_numpy(x).cuda() for x in _gemm_inputs(512, 768, 1024))
    ours = fa.gemm_fused_epilogue(a, b, bias, res)
    theirs = torch.zeros_like(ours)
    RefGpu().gemm_fused(a, b, bias, res, theirs)
    torch.cuda.synchronize()
    assert oracle.max_abs(ours.cpu().numpy(), theirs.cpu().numpy()) <= 2e-3
    assert oracle.rel_l2(ours.cpu().numpy(), theirs.cpu().numpy()) <= 1e-3


# ------------------------------------- softmax -------------------------------------------
def test_softmax_golden_vector(fa, oracle, golden):
    g = golden("softmax_rows.npz")
    p = fa.softmax_rows(torch.from_numpy(g["x"]).cuda()).cpu().numpy()
    assert oracle.max_abs(p, g["p"]) <= 1e-6 and oracle.rel_l2(p, g["p"]) <= 1e-6


# the reference's three softmax programs: 4096x512 (naive, hybrid), 4096x32 (warp); plus ragged/unaligned/long rows
@pytest.mark.parametrize("rows,n", [(4096, 512), (4096, 32), (1, 1), (3, 5), (257, 1000), (64, 4096), (7, 4097),
                                    (5, 8192), (3, 20000), (1000, 132)])
def test_softmax_parity_with_oracle(fa, oracle, rows, n):
    rng = np.random.default_rng(rows + n)
    x = (rng.standard_normal((rows, n)) * 4).astype(np.float32)
    p = fa.softmax_rows(torch.from_numpy(x).cuda()).cpu().numpy()
    ref = oracle.softmax_rows(x)
    # tolerance: fp32 exp (2 ulp) + reduction order -- the oracle sums n terms sequentially in fp32 (relative
    # error up to ~sqrt(n) * 2^-24), the kernel reduces as a tree; both are far inside the north-star bounds
    assert oracle.max_abs(p, ref) <= 1e-4 and oracle.rel_l2(p, ref) <= 1e-4
    assert np.allclose(p.sum(axis=1), 1.0, atol=1e-5)


def test_softmax_reference_pattern_input_and_in_place(fa, oracle):
    """The reference harness' input ((i % 13) - 6) * 0.1 (softmax_naive_baseline.cu:148-151), in place."""
    rows, n = 512, 4096
    x = (((np.arange(rows * n) % 13) - 6) * 0.1).astype(np.float32).reshape(rows, n)
    t = torch.from_numpy(x).cuda()
    fa.softmax_rows(t, out=t)
    torch.cuda.synchronize()
    assert oracle.max_abs(t.cpu().numpy(), oracle.softmax_rows(x)) <= 1e-7


def test_softmax_config5b_full_size_properties(fa, oracle):
    """BASELINE config 5b (16384 x 4096): sums to one, shift invariance, monotone, sampled rows vs oracle."""
    rows, n = 16384, 4096
    g = torch.Generator(device="cuda").manual_seed(42)
    x = torch.rand((rows, n), device="cuda", generator=g) * 8 - 4
    p = fa.softmax_rows(x)
    torch.cuda.synchronize()
    assert (p.sum(dim=1) - 1).abs().max().item() <= 1e-5
    assert (p >= 0).all() and torch.isfinite(p).all()
    shifted = fa.softmax_rows(x + 1.0)  # softmax(x + c) == softmax(x) up to rounding of x + c
    assert (shifted - p).abs().max().item() <= 1e-6
    idx = torch.tensor([0, 1, 4095, 8191, 16383], device="cuda")
    ref = oracle.softmax_rows(x[idx].cpu().numpy())
    assert oracle.max_abs(p[idx].cpu().numpy(), ref) <= 1e-6
    # order preserving within a row
    xs, order = x[7].sort()
    assert (p[7][order].diff() >= 0).all()


@pytest.mark.skipif(not ref_available("softmax_naive"), reason="oracle/_ref not built")
def test_softmax_against_the_reference_kernel_on_the_same_gpu(fa, oracle):
    x = torch.randn((1024, 4096), device="cuda") * 3
    ours = fa.softmax_rows(x)
    theirs = torch.zeros_like(x)
    RefGpu().softmax_naive(x, theirs)
    torch.cuda.synchronize()
    assert oracle.max_abs(ours.cpu().numpy(), theirs.cpu().numpy()) <= 1e-6


@pytest.mark.skipif(not (ref_available("softmax_hybrid") and ref_available("softmax_warp")), reason="oracle/_ref not built")
@pytest.mark.parametrize("rows,n,kind", [(4096, 512, "hybrid"), (1000, 384, "hybrid"), (333, 100, "hybrid"),
                                         (4096, 32, "warp"), (1001, 17, "warp"), (64, 1, "warp")])
def test_softmax_against_the_reference_sibling_kernels(fa, oracle, rows, n, kind):
    """softmax_rows also stands in for softmax_block_hybrid_kernel (N <= 512, `__expf`;
    Softmax micro kernel/softmax_block_hybrid/softmax_block_hybrid.cu:74-169) and softmax_warp_shfl_kernel (N <= 32;
    ../softmax_warp_shfl_baseline/softmax_warp_shfl_baseline.cu:69-115): same buffers, both kernels."""
    x = torch.randn((rows, n), device="cuda", generator=torch.Generator(device="cuda").manual_seed(rows + n)) * 3
    ours = fa.softmax_rows(x)
    theirs = torch.zeros_like(x)
    r = RefGpu()
    (r.softmax_block_hybrid if kind == "hybrid" else r.softmax_warp_shfl)(x, theirs)
    torch.cuda.synchronize()
    a, b = ours.cpu().numpy(), theirs.cpu().numpy()
    assert oracle.max_abs(a, b) <= (2e-6 if kind == "hybrid" else 1e-6)  # `__expf` in the hybrid kernel: 2 ulp of exp
    assert oracle.max_abs(a, oracle.softmax_rows(x.cpu().numpy())) <= 4e-6  # sequential vs tree row sums


def test_gemm_residual_in_place(fa, oracle):
    """C may alias the residual, as in the reference kernel (read residual[idx], then write C[idx],
    tensorcore_gemm_fused_epilogue.cu:119-126): the epilogue reads a block's residual before it stores that block."""
    M, N, K = 1000, 1096, 264
    a, b, bias, res = _gemm_inputs(M, N, K, seed=77)
    ref = oracle.gemm_fused_f16(a, b, bias, res)
    c = torch.from_numpy(res.copy()).cuda()
    out = fa.gemm_fused_epilogue(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(bias).cuda(), c, out=c)
    torch.cuda.synchronize()
    assert out.data_ptr() == c.data_ptr()
    got = c.cpu().numpy()
    assert oracle.max_abs(got, ref) <= 2e-3 and oracle.rel_l2(got, ref) <= 1e-3
    separate = _gemm(fa, a, b, bias, res)
    assert np.array_equal(got, separate)


@pytest.mark.parametrize("act", ["identity", "relu", "gelu", "silu"])
@pytest.mark.parametrize("shape", [(129, 264, 72), (512, 512, 256)])
def test_gemm_epilogue_activations(fa, oracle, act, shape):
    """gemm_fused_epilogue_act: act between bias and residual.  "relu" is the reference's epilogue
    (tensorcore_gemm_fused_epilogue.cu:109-126) and must equal gemm_fused_epilogue bit for bit; identity / GELU /
    SiLU are checked against the oracle's restatement of the same order of operations."""
    M, N, K = shape
    a, b, bias, res = _gemm_inputs(M, N, K, seed=7 + M)
    ta, tb, tbias, tres = (torch.from_numpy(x).cuda() for x in (a, b, bias, res))
    got = fa.gemm_fused_epilogue(ta, tb, tbias, tres, act=act)
    torch.cuda.synchronize()
    ref = oracle.gemm_fused_f16(a, b, bias, res, act=act)
    assert oracle.max_abs(got.cpu().numpy(), ref) <= 2e-3 and oracle.rel_l2(got.cpu().numpy(), ref) <= 1e-3
    if act == "relu":
        assert torch.equal(got, fa.gemm_fused_epilogue(ta, tb, tbias, tres))
    # without bias and residual the activation is applied to the bare product
    got2 = fa.gemm_fused_epilogue(ta, tb, None, None, act=act).cpu().numpy()
    ref2 = oracle.gemm_fused_f16(a, b, None, None, act=act)
    assert oracle.max_abs(got2, ref2) <= 2e-3 and oracle.rel_l2(got2, ref2) <= 1e-3


def test_gemm_rejects_unknown_activation(fa):
    import ctypes
    lib = fa.load_library()
    t = torch.zeros((128, 128), device="cuda", dtype=torch.float16)
    c = torch.zeros((128, 128), device="cuda")
    p = lambda x: ctypes.c_void_p(x.data_ptr())
    assert lib.gemm_fused_epilogue_act(p(t), p(t), None, None, p(c), 128, 128, 128, 9, None) == -7


@pytest.mark.parametrize("shape", [(1024, 1024, 1024), (129, 264, 72), (512, 260, 2052), (16, 16, 16)])
def test_fp32_operand_gemm_parity(fa, oracle, shape):
    """gemm_fused_epilogue_f32 (the reference's gemm_tiled_fused_kernel, GEMM/gemm_fused_epilogue/gemm_fused_epilogue.cu:
    44-111; its harness shape is 1024^3 with mt19937(42) uniform(-1,1)): fp32 operands on the tensor cores as split
    tf32 pairs.  Tolerance: the north-star's max-abs <= 2e-3 and rel-L2 <= 1e-3 against the fp32 CPU reference --
    measured 1.4e-4 at 1024^3, where a PLAIN tf32 product is 1.3e-2 off: the products are fp32-exact, what remains is
    the tensor core's truncating fp32 accumulation (K / 8 accumulate steps on the large term; the cross terms have
    their own accumulator), so the second assertion demands at least 10x less error than tf32-rounded operands give."""
    M, N, K = shape
    rng = np.random.default_rng(42)
    a, b = rng.uniform(-1, 1, (M, K)).astype(np.float32), rng.uniform(-1, 1, (K, N)).astype(np.float32)
    bias, res = rng.uniform(-1, 1, N).astype(np.float32), rng.uniform(-1, 1, (M, N)).astype(np.float32)
    ta, tb, tbias, tres = (torch.from_numpy(x).cuda() for x in (a, b, bias, res))
    got = fa.gemm_fused_epilogue_f32(ta, tb, tbias, tres)
    torch.cuda.synchronize()
    ref = oracle.gemm_fused(a, b, bias, res)
    assert oracle.max_abs(got.cpu().numpy(), ref) <= 1e-3 and oracle.rel_l2(got.cpu().numpy(), ref) <= 5e-5
    # nullable operands and another activation; compare with what tf32-rounded operands would give
    got2 = fa.gemm_fused_epilogue_f32(ta, tb, None, None, act="identity").cpu().numpy().astype(np.float64)
    exact = a.astype(np.float64) @ b.astype(np.float64)
    tf32 = lambda x: ((x.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32).astype(np.float64)
    err_tf32 = np.abs(tf32(a) @ tf32(b) - exact).max()
    assert np.abs(got2 - exact).max() <= max(err_tf32 / 10, 2e-5), (np.abs(got2 - exact).max(), err_tf32)
    if ref_available("gemm_fp32_fused") and M % 32 == 0:
        theirs = torch.zeros_like(got)
        RefGpu().gemm_fp32_fused(ta, tb, tbias, tres, theirs)
        torch.cuda.synchronize()
        assert (got - theirs).abs().max().item() <= 2e-3
