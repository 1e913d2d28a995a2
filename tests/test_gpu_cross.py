"""GPU parity tests for fa_forward_cross: the D = 16 programs BASELINE.json names (v12i, v7_5) at their native
shapes -- against the golden vectors their own CPU oracles produced, and against their own kernels run on the
same GPU -- and Nq != Nkv at D = 64 / 128 through the tcgen05 kernel."""
import numpy as np
import pytest
import torch

from tests.oracle_lib import RefCpu, RefGpu, ref_available

pytestmark = pytest.mark.gpu


def _cross(fa, q, k, v, scale):
    out = fa.fa_forward_cross(torch.from_numpy(q).cuda(), torch.from_numpy(k).cuda(), torch.from_numpy(v).cuda(), scale)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def test_v12i_native_golden(fa, oracle, golden):
    g = golden("attn_v12i_native.npz")  # B=8, M=16, L=128, KDIM=DV=16, scale 1/4
    q, k, v = (g[x].view(np.float16) for x in ("q", "k", "v"))
    got = _cross(fa, q[None], k[None], v[None], float(g["scale"]))[0]
    assert oracle.rel_l2(got, g["o"]) <= 1e-5 and oracle.max_abs(got, g["o"]) <= 1e-5  # fp32 path: far inside 1e-3 / 2e-3


def test_v7_5_native_golden(fa, oracle, golden):
    g = golden("attn_v7_5_native.npz")  # B=32, single query, S=128, D=16
    q, k, v = (g[x].view(np.float16) for x in ("q", "k", "v"))
    got = _cross(fa, q.reshape(1, -1, 1, 16), k[None], v[None], float(g["scale"]))
    assert oracle.rel_l2(got.reshape(-1, 16), g["o"]) <= 1e-5 and oracle.max_abs(got.reshape(-1, 16), g["o"]) <= 1e-5


@pytest.mark.skipif(not (ref_available("v12i") and ref_available("v7_5")), reason="oracle/_ref not built")
def test_against_the_named_reference_kernels_at_their_native_shapes(fa, oracle):
    """flashattn_2warp_nosmemW_online_softmax<16,16,4,32> (B=1024, M=16, L=128) and
    flashattn_streaming_16x16_kernel_mw_v7_5 (B=1024, S=128), compiled unmodified, same device buffers."""
    cpu, ref = RefCpu(), RefGpu()
    q, k, v = cpu.v12i_fill_uniform(123, (1024, 16, 16), (1024, 128, 16), (1024, 128, 16))  # the reference's fixture
    tq, tk, tv = (torch.from_numpy(x).cuda() for x in (q, k, v))
    theirs = torch.zeros((1024, 16, 16), device="cuda")
    ref.v12i_kernel(tq, tk, tv, theirs, 0.25)
    ours = fa.fa_forward_cross(tq[None], tk[None], tv[None], 0.25)[0]
    torch.cuda.synchronize()
    assert oracle.rel_l2(ours.cpu().numpy(), theirs.cpu().numpy()) <= 1e-3
    assert oracle.max_abs(ours.cpu().numpy(), theirs.cpu().numpy()) <= 2e-3
    full = oracle.attention_f16(q, k, v, 0.25)
    assert oracle.rel_l2(ours.cpu().numpy(), full) <= 1e-5

    q, k, v = cpu.v7_5_fill(0, 1024, 128)
    tq, tk, tv = (torch.from_numpy(x).cuda() for x in (q, k, v))
    theirs = torch.zeros((1024, 16), device="cuda")
    ref.v7_5_kernel(tq, tk, tv, theirs)
    ours = fa.fa_forward_cross(tq.reshape(1, 1024, 1, 16), tk[None], tv[None], 0.25).reshape(1024, 16)
    torch.cuda.synchronize()
    assert oracle.rel_l2(ours.cpu().numpy(), theirs.cpu().numpy()) <= 1e-3
    assert oracle.max_abs(ours.cpu().numpy(), theirs.cpu().numpy()) <= 2e-3


@pytest.mark.parametrize("shape", [(1, 3, 1, 1, 16), (2, 2, 5, 33, 16), (1, 4, 16, 1000, 16), (1, 2, 200, 7, 16),
                                   (1, 2, 100, 300, 64), (2, 1, 300, 129, 128), (1, 1, 1, 4096, 128), (1, 2, 513, 64, 64)])
def test_cross_shapes_against_oracle(fa, oracle, shape):
    B, H, Nq, Nkv, D = shape
    rng = np.random.default_rng(Nq * 7 + Nkv)
    q = rng.uniform(-1, 1, (B, H, Nq, D)).astype(np.float16)
    k = rng.uniform(-1, 1, (B, H, Nkv, D)).astype(np.float16)
    v = rng.uniform(-1, 1, (B, H, Nkv, D)).astype(np.float16)
    scale = 1.0 / np.sqrt(D)
    got = _cross(fa, q, k, v, scale).reshape(B * H, Nq, D)
    ref = oracle.attention_f16(q.reshape(B * H, Nq, D), k.reshape(B * H, Nkv, D), v.reshape(B * H, Nkv, D), scale)
    assert np.isfinite(got).all()
    assert oracle.rel_l2(got, ref) <= 1e-3 and oracle.max_abs(got, ref) <= 2e-3


def test_cross_rejects_unsupported_head_dim(fa):
    q = torch.zeros((1, 1, 8, 32), dtype=torch.float16, device="cuda")
    with pytest.raises(fa.FaError) as e:
        fa.fa_forward_cross(q, q, q, 0.1)
    assert e.value.code == -2


@pytest.mark.skipif(not (ref_available("mw_v10") and ref_available("mw_v11")), reason="oracle/_ref not built")
@pytest.mark.parametrize("ver", ["v10", "v11"])
def test_against_the_wmma_streaming_kernels_v10_v11_at_their_native_shape(fa, oracle, ver):
    """flashattn_streaming_16x16_kernel_mw_v10 / _v11 (QK^T on WMMA, two warps, cp.async ping-pong;
    flashattn_warp_spc/flashattn_streaming_16x16_mw_v10.cu:104-281, ..._v11.cu:101-283) at B = 1024, M = 16, L = 128,
    D = 16, normal(0,1) inputs as in their harness, run on the same buffers as fa_forward_cross; both against the
    streaming family's host oracle (`+1e-6` normaliser)."""
    B, L = 1024, 128
    rng = np.random.default_rng(42)
    q = rng.standard_normal((B, 16, 16)).astype(np.float16)
    k = rng.standard_normal((B, L, 16)).astype(np.float16)   # row per key = the kernels' K_T
    v = rng.standard_normal((B, L, 16)).astype(np.float16)
    tq, tk, tv = (torch.from_numpy(x).cuda() for x in (q, k, v))
    theirs = torch.zeros((B, 16, 16), device="cuda")
    RefGpu().streaming_mw(ver, tq, tk, tv, theirs, 0.25)
    ours = fa.fa_forward_cross(tq[None], tk[None], tv[None], 0.25)
    torch.cuda.synchronize()
    ref = oracle.streaming_ref_f16(q, np.ascontiguousarray(k.transpose(0, 2, 1)), v, 0.25)
    a, b = ours[0].cpu().numpy(), theirs.cpu().numpy()
    assert oracle.max_abs(a, ref) <= 2e-3 and oracle.rel_l2(a, ref) <= 1e-3
    assert oracle.max_abs(b, ref) <= 2e-3 and oracle.rel_l2(b, ref) <= 1e-3
    assert oracle.max_abs(a, b) <= 2e-3 and oracle.rel_l2(a, b) <= 1e-3
