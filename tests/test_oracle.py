"""CPU tests of the CHECKER: the C restatement (oracle/fa_oracle.c) against the committed golden
vectors (produced by the reference's own compiled functions, tests/golden/make_golden.py), against
oracle/_ref when it is present, and against independent numpy float64 math."""
import numpy as np
import pytest

from tests.oracle_lib import RefCpu, ref_available


def _np_attention(q, k, v, scale):
    s = np.einsum("bmk,btk->bmt", q.astype(np.float64), k.astype(np.float64)) * scale
    s -= s.max(axis=-1, keepdims=True)
    p = np.exp(s)
    p /= p.sum(axis=-1, keepdims=True)
    return np.einsum("bmt,btd->bmd", p, v.astype(np.float64))


@pytest.mark.parametrize("name", ["attn_cfg1_v12i.npz", "attn_v12i_native.npz", "attn_ragged_d128.npz"])
def test_attention_f16_matches_reference_golden_bit_exact(oracle, golden, name):
    g = golden(name)
    out = oracle.attention_f16(g["q"], g["k"], g["v"], float(g["scale"]))
    assert np.array_equal(out, g["o"]), "oracle differs from the reference's cpu_reference output"
    # the threaded path is the same arithmetic
    assert np.array_equal(oracle.attention_f16(g["q"], g["k"], g["v"], float(g["scale"]), nthreads=1), g["o"])


def test_single_query_golden_from_v7_5(oracle, golden):
    """v7_5's cpu_ref_attention (one query per batch element, -1e30 sentinel) is the same arithmetic as
    v12i's cpu_reference with M_TOTAL = 1: the restatement reproduces its output bit for bit."""
    g = golden("attn_v7_5_native.npz")
    q = g["q"].reshape(-1, 1, 16)
    out = oracle.attention_f16(q, g["k"], g["v"], float(g["scale"]))
    assert np.array_equal(out.reshape(-1, 16), g["o"])


def test_attention_rows_sampling_equals_full(oracle, golden):
    g = golden("attn_ragged_d128.npz")
    rows = np.array([0, 1, 127, 128, 199, 200, 333, 399], dtype=np.int64)
    out = oracle.attention_f16_rows(g["q"], g["k"], g["v"], float(g["scale"]), rows)
    assert np.array_equal(out, g["o"].reshape(-1, g["o"].shape[-1])[rows])


def test_attention_f32_oracles_match_golden(oracle, golden):
    g = golden("attn_f32_oracles.npz")
    assert np.array_equal(oracle.attention_f32_dacc(g["q"], g["k"], g["v"]), g["o_542"])
    assert np.array_equal(oracle.attention_f32(g["q"], g["k"], g["v"], float(g["scale_block"])), g["o_block"])


def test_oracles_agree_with_each_other_and_float64(oracle, golden):
    # cross-check prescribed by SURVEY 8(c): v12i cpu_reference vs 5.4.2 (double accumulate) on config 1
    g = golden("attn_cfg1_v12i.npz")
    q, k, v = (oracle.half_to_float(g[n]).reshape(1, 128, 64) for n in ("q", "k", "v"))
    d = oracle.attention_f32_dacc(q, k, v)  # scale = 1/sqrt(64) = 0.125, same as the fixture
    assert oracle.rel_l2(g["o"], d) < 2e-6
    ref64 = _np_attention(q, k, v, 0.125)
    assert oracle.rel_l2(g["o"], ref64.astype(np.float32)) < 2e-6
    assert oracle.max_abs(g["o"], ref64.astype(np.float32)) < 1e-6


def test_gemm_fused_matches_golden_and_nullable_operands(oracle, golden):
    g = golden("gemm_fused.npz")
    assert np.array_equal(oracle.gemm_fused(g["a"], g["b"], g["bias"], g["residual"]), g["c"])
    a, b = g["a"], g["b"]
    plain = oracle.gemm_fused(a, b, None, None)
    ref = np.maximum(a.astype(np.float64) @ b.astype(np.float64), 0)
    assert np.allclose(plain, ref, atol=1e-4)
    # epilogue order: bias -> ReLU -> residual (tensorcore_gemm_fused_epilogue.cu:109-126)
    full = oracle.gemm_fused(a, b, g["bias"], g["residual"])
    ref = np.maximum(a.astype(np.float64) @ b.astype(np.float64) + g["bias"], 0) + g["residual"]
    assert np.allclose(full, ref, atol=1e-4)
    # row-block sampling leaves other rows untouched and equals the full result on its rows
    part = oracle.gemm_fused(a, b, g["bias"], g["residual"], row_begin=8, row_end=24)
    assert np.array_equal(part[8:24], full[8:24]) and not part[:8].any() and not part[24:].any()
    # fp16-operand variant == fp32 variant on half-rounded operands
    ah, bh = a.astype(np.float16), b.astype(np.float16)
    f16 = oracle.gemm_fused_f16(ah, bh, g["bias"], g["residual"], row_begin=8, row_end=24)
    f32 = oracle.gemm_fused(ah.astype(np.float32), bh.astype(np.float32), g["bias"], g["residual"])
    assert np.array_equal(f16, f32[8:24])


def test_gemm_batched_and_qk_scores_match_golden(oracle, golden):
    g = golden("gemm_batched.npz")
    c = oracle.gemm_batched(g["a"], g["b"])
    assert np.array_equal(c, g["c"])
    assert np.allclose(c, np.einsum("bmk,bkn->bmn", g["a"].astype(np.float64), g["b"].astype(np.float64)), atol=1e-4)
    # one batch entry of the batched GEMM == the fused GEMM's contraction before its epilogue on non-negative results
    g = golden("qk_scores.npz")
    s = oracle.qk_scores(g["q"], g["k"], float(g["scale"]))
    assert np.array_equal(s, g["s"])
    assert np.allclose(s, 0.125 * np.einsum("bik,bjk->bij", g["q"].astype(np.float64), g["k"].astype(np.float64)), atol=1e-6)
    # Nk != N is accepted by the restatement (the reference fixes Nk = N)
    s2 = oracle.qk_scores(g["q"], g["k"][:, :17], float(g["scale"]))
    assert np.array_equal(s2, s[:, :, :17])


def test_streaming_family_oracle_matches_golden(oracle, golden):
    g = golden("attn_streaming16.npz")
    q, kt, v = (g[n].view(np.float16) for n in ("q", "kt", "v"))
    o = oracle.streaming_ref_f16(q, kt, v, float(g["scale"]))
    assert np.array_equal(o, g["o"])
    # the `+1e-6` normaliser and the materialised P differ from cpu_reference (no EPS) far below the tolerance
    k = np.ascontiguousarray(kt.transpose(0, 2, 1))
    plain = oracle.attention_f16(q, k, v, float(g["scale"]))
    assert oracle.max_abs(o, plain) < 2e-6 and oracle.rel_l2(o, plain) < 2e-6


@pytest.mark.parametrize("tile", [16, 128, 50])
def test_stage_oracle_streaming_state_invariants(oracle, golden, tile):
    """(m, l, y) after every key tile: m is the running row max, y/l the attention over the prefix, and the snapshots at
    common prefix lengths do not depend on the tile size."""
    g = golden("attn_ragged_d128.npz")  # [2, 200, 128], scale 1/sqrt(128)
    q, k, v = (g[n].view(np.float16) for n in ("q", "k", "v"))
    scale = float(g["scale"])
    m, l, y = oracle.streaming_state_f16(q, k, v, tile, scale)
    T = (200 + tile - 1) // tile
    assert m.shape == (T, 2, 200) and y.shape == (T, 2, 200, 128)
    s = oracle.qk_scores(q.astype(np.float32), k.astype(np.float32), 1.0)  # double accumulate, then one rounding
    for t in range(T):
        end = min(200, (t + 1) * tile)
        prefix = oracle.attention_f16(q, k[:, :end], v[:, :end], scale)
        assert oracle.rel_l2(y[t] / l[t][..., None], prefix) < 5e-6
        assert np.allclose(m[t], (s[:, :, :end] * np.float32(scale)).max(axis=-1), rtol=0, atol=2e-6)
        assert (l[t] >= 1.0 - 1e-6).all()  # the max element contributes exp(0) = 1
    # last snapshot == full attention == the committed golden output of the reference's cpu_reference
    assert oracle.rel_l2(y[-1] / l[-1][..., None], g["o"]) < 5e-6
    if tile != 50:
        m50, l50, y50 = oracle.streaming_state_f16(q, k, v, 200, scale)  # a single tile: no merge at all
        assert np.array_equal(m[-1], m50[0])
        assert np.allclose(l[-1], l50[0], rtol=1e-5) and np.allclose(y[-1], y50[0], rtol=1e-4, atol=1e-5)


def test_softmax_rows_matches_golden(oracle, golden):
    g = golden("softmax_rows.npz")
    p = oracle.softmax_rows(g["x"])
    assert np.array_equal(p, g["p"])
    assert np.allclose(p.sum(axis=1), 1.0, atol=1e-5)


def test_half_conversions_are_exact(oracle):
    allh = np.arange(65536, dtype=np.uint16)
    f = oracle.half_to_float(allh)
    want = allh.view(np.float16).astype(np.float32)
    assert np.array_equal(f.view(np.uint32), want.view(np.uint32))
    rng = np.random.default_rng(0)
    x = rng.standard_normal(100000).astype(np.float32) * rng.choice([1e-8, 1e-5, 1e-3, 1, 100, 6e4], 100000).astype(np.float32)
    x = np.concatenate([x, f[np.isfinite(f)], np.array([65519.9, 65520.0, 2.98e-8, 2.9802322e-8, 0.0, -0.0], np.float32)])
    with np.errstate(over="ignore"):
        assert np.array_equal(oracle.float_to_half(x), x.astype(np.float16).view(np.uint16))


def test_metrics(oracle):
    a = np.array([1.0, 2.0, 3.0], np.float32)
    b = np.array([1.0, 2.5, 2.0], np.float32)
    assert oracle.max_abs(a, b) == pytest.approx(1.0)
    assert oracle.rel_l2(a, b) == pytest.approx(np.sqrt(1.25 / 11.25))
    assert oracle.rel_l2(a, a) == 0.0


@pytest.mark.skipif(not all(ref_available(n) for n in ("v12i", "fa542", "softmax_block", "gemm_fused")),
                    reason="oracle/_ref not built (reference tree absent)")
def test_oracle_bit_exact_against_compiled_reference(oracle):
    """Fresh seeded inputs (not the committed fixtures) through the reference's own functions."""
    r = RefCpu()
    for seed, shp in [(1, (3, 40, 64)), (2, (2, 130, 128)), (3, (5, 16, 16))]:
        q, k, v = r.v12i_fill_uniform(seed, shp, shp, shp)
        sc = 1.0 / np.sqrt(shp[-1])
        a, b = r.v12i_cpu_reference(q, k, v, sc), oracle.attention_f16(q, k, v, sc)
        assert np.array_equal(a, b)
        assert r.v12i_rel_l2(a, b) == oracle.rel_l2(a, b) == 0.0
    # Nq != Nkv and KDIM != DV are runtime-generic in cpu_reference (:248)
    q, _, _ = r.v12i_fill_uniform(4, (2, 7, 32), (1,), (1,))
    _, k, v = r.v12i_fill_uniform(5, (1,), (2, 50, 32), (2, 50, 48))
    assert np.array_equal(r.v12i_cpu_reference(q, k, v, 0.2), oracle.attention_f16(q, k, v, 0.2))
    rng = np.random.default_rng(9)
    qf, kf, vf = (rng.standard_normal((2, 33, 24)).astype(np.float32) for _ in range(3))
    assert np.array_equal(r.fa542_cpu_ref(qf, kf, vf), oracle.attention_f32_dacc(qf, kf, vf))
    assert np.array_equal(r.softmax_block_flashattn_ref(qf, kf, vf, 0.7), oracle.attention_f32(qf, kf, vf, 0.7))
    a, b, bias, res = r.gemm_fused_fill(43, 32, 48, 64)
    assert np.array_equal(r.gemm_fused_cpu(a, b, bias, res), oracle.gemm_fused(a, b, bias, res))
    if r.st16 is not None:
        q, kt, v = r.streaming16_fill(7, 5, 64)
        assert np.array_equal(r.streaming16_cpu(q, kt, v, 0.25), oracle.streaming_ref_f16(q, kt, v, 0.25))
    if r.bgemm is not None and r.qk is not None:
        ab, bb = rng.standard_normal((2, 19, 33)).astype(np.float32), rng.standard_normal((2, 33, 21)).astype(np.float32)
        assert np.array_equal(r.gemm_batched_cpu(ab, bb), oracle.gemm_batched(ab, bb))
        qq, kk = rng.standard_normal((3, 20, 24)).astype(np.float32), rng.standard_normal((3, 20, 24)).astype(np.float32)
        assert np.array_equal(r.qk_cpu(qq, kk, 0.2041), oracle.qk_scores(qq, kk, 0.2041))


def test_fp32_fused_gemm_oracle_matches_the_reference_fp32_file(oracle):
    """fa_oracle_gemm_fused also restates cpu_gemm_fused_ref of the fp32 sibling program
    (GEMM/gemm_fused_epilogue/gemm_fused_epilogue.cu:116-136): bit-exact against that function compiled unmodified."""
    if not ref_available("gemm_fp32_fused"):
        pytest.skip("oracle/_ref/libref_gemm_fp32_fused.so not built")
    ref = RefCpu()
    rng = np.random.default_rng(42)
    M, N, K = 70, 96, 132
    a, b = rng.uniform(-1, 1, (M, K)).astype(np.float32), rng.uniform(-1, 1, (K, N)).astype(np.float32)
    bias, res = rng.uniform(-1, 1, N).astype(np.float32), rng.uniform(-1, 1, (M, N)).astype(np.float32)
    assert np.array_equal(oracle.gemm_fused(a, b, bias, res), ref.gemm_fp32_fused_cpu(a, b, bias, res))
