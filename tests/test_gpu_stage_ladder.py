"""Stage ladder: the fused kernel's carried online-softmax state is walked ONE 128-key tile per call through
fa_forward_partial and checked after every tile against the stage oracle (the reference's tile-to-tile merge rule,
Streaming_FlashAttention_Forward_Kernel/flashattn_streaming_16x16.cu:170-261), and the streaming family's own
kernel is run beside fa_forward_cross at its native shape.  Counterpart of the reference's stage-by-stage debug ladder
(FlashAttention/flashattn_fused_softmax_tensorcore_16x16/flashattn_{qkt_wmma,qkt_softmax,softmax_warp,pv_wmma}_debug_16x16.cu):
QK^T and softmax stages are covered by tests/test_gpu_gemm_batched.py and tests/test_gpu_unfused.py."""
import numpy as np
import pytest
import torch

from tests.oracle_lib import RefGpu, ref_available

pytestmark = pytest.mark.gpu

LN2 = float(np.log(2.0))


@pytest.mark.parametrize("D,dist", [(128, "uniform"), (64, "normal"), (128, "ramp")])
def test_state_after_every_key_tile_matches_stage_oracle(fa, oracle, D, dist):
    B, H, Nq, L, tile = 1, 2, 200, 1000, 128  # ragged query tile, ragged last key tile (1000 = 7 * 128 + 104)
    rng = np.random.default_rng(5 + D)
    if dist == "uniform":
        q, k, v = (rng.uniform(-1, 1, (B * H, n, D)).astype(np.float16) for n in (Nq, L, L))
    else:
        q, k, v = (rng.standard_normal((B * H, n, D)).astype(np.float16) for n in (Nq, L, L))
    if dist == "ramp":  # keys grow along the sequence: every tile raises the row max, every merge rescales
        k = (k.astype(np.float32) * 0.2 + np.linspace(0, 2.5, L, dtype=np.float32)[None, :, None] * np.sign(q.astype(np.float32).mean(axis=1, keepdims=True))).astype(np.float16)
    scale = D ** -0.5
    m_ref, l_ref, y_ref = oracle.streaming_state_f16(q, k, v, tile, scale)
    tq = torch.from_numpy(q).cuda().view(B, H, Nq, D)
    tk = torch.from_numpy(k).cuda().view(B, H, L, D)
    tv = torch.from_numpy(v).cuda().view(B, H, L, D)
    o_acc = torch.empty((B, H, Nq, D), device="cuda")
    m = torch.empty((B, H, Nq), device="cuda")
    l = torch.empty((B, H, Nq), device="cuda")
    T = (L + tile - 1) // tile
    final = None
    for t in range(T):
        ks = tk[:, :, t * tile:(t + 1) * tile].contiguous()
        vs = tv[:, :, t * tile:(t + 1) * tile].contiguous()
        final = fa.fa_forward_partial(tq, ks, vs, o_acc, m, l, scale, t == 0, t == T - 1)
        torch.cuda.synchronize()
        if t == T - 1:
            break  # the closing step emits the normalised output instead of a state (include/fa_b200.h)
        gm, gl, gy = (x.cpu().numpy().reshape((B * H, Nq) + x.shape[3:]) for x in (m, l, o_acc))
        # invariants of the state (the kernel may keep a stale max within the lazy-rescale slack, so m, l, y are
        # compared through what they define): normalised output and log-sum-exp
        assert oracle.rel_l2(gy / gl[..., None], y_ref[t] / l_ref[t][..., None]) <= 1e-3, f"tile {t}: O"
        lse, lse_ref = gm + np.log(gl), m_ref[t] + np.log(l_ref[t])
        assert np.abs(lse - lse_ref).max() <= 2e-3, f"tile {t}: log-sum-exp"
        gap = m_ref[t] - gm  # true running max minus carried max
        assert (gap >= -1e-3).all() and (gap <= 8 * LN2 + 1e-3).all(), f"tile {t}: carried max"
    got = final.float().cpu().numpy().reshape(B * H, Nq, D)
    last = y_ref[-1] / l_ref[-1][..., None]
    assert oracle.rel_l2(got, last) <= 1e-3
    ref = oracle.attention_f16(q, k, v, scale)
    assert oracle.rel_l2(got, ref) <= 1e-3 and oracle.max_abs(got, ref) <= 2e-3 * max(1.0, float(np.abs(ref).max()))


def test_streaming_family_golden_and_native_shape(fa, oracle, golden):
    """fa_forward_cross at the streaming family's native shape (M = 16 queries, D = 16, L = 128, K stored transposed)
    against its host oracle's golden vector."""
    g = golden("attn_streaming16.npz")
    q, kt, v = (g[n].view(np.float16) for n in ("q", "kt", "v"))
    k = np.ascontiguousarray(kt.transpose(0, 2, 1))  # our boundary takes K as [.., L, D] rows like v12i / WMMA v1-v5
    out = fa.fa_forward_cross(torch.from_numpy(q).cuda()[None], torch.from_numpy(k).cuda()[None],
                              torch.from_numpy(v).cuda()[None], float(g["scale"]))
    torch.cuda.synchronize()
    got = out[0].cpu().numpy()
    assert oracle.max_abs(got, g["o"]) <= 2e-3 and oracle.rel_l2(got, g["o"]) <= 1e-3


@pytest.mark.skipif(not ref_available("streaming16"), reason="oracle/_ref not built")
def test_reference_streaming_kernel_beside_ours_and_the_stage_oracle(fa, oracle):
    """The reference's own streaming kernel (WMMA QK^T, tile-to-tile merge, N_TILE = 16) on the same buffers: pins the
    stage oracle's final snapshot to device code that applies the merge rule 7 times per row (L = 128)."""
    r = RefGpu()
    B, L = 256, 128
    rng = np.random.default_rng(42)
    q = rng.standard_normal((B, 16, 16)).astype(np.float16)
    kt = rng.standard_normal((B, 16, L)).astype(np.float16)
    v = rng.standard_normal((B, L, 16)).astype(np.float16)
    k = np.ascontiguousarray(kt.transpose(0, 2, 1))
    theirs = torch.zeros((B, 16, 16), device="cuda")
    r.streaming16(torch.from_numpy(q).cuda(), torch.from_numpy(kt).cuda(), torch.from_numpy(v).cuda(), theirs, 0.25)
    ours = fa.fa_forward_cross(torch.from_numpy(q).cuda()[None], torch.from_numpy(k).cuda()[None],
                               torch.from_numpy(v).cuda()[None], 0.25)
    torch.cuda.synchronize()
    theirs, ours = theirs.cpu().numpy(), ours[0].cpu().numpy()
    assert oracle.max_abs(ours, theirs) <= 2e-3 and oracle.rel_l2(ours, theirs) <= 1e-3
    m, l, y = oracle.streaming_state_f16(q, k, v, 16, 0.25)
    stage = y[-1] / (l[-1][..., None] + np.float32(1e-6))  # the kernel's final normalisation (:266-271)
    assert oracle.max_abs(stage, theirs) <= 1e-4 and oracle.rel_l2(stage, theirs) <= 1e-4  # __expf / butterfly order only
