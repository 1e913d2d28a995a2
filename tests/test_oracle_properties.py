"""Property tests of the CPU oracle (hypothesis): the algebraic facts the GPU parity tests lean on at sizes where the
oracle cannot be run in full -- attention is a set function of the (key, value) pairs, linear in V, invariant to a
common shift of a row's scores; the streaming (tile-merge) state defines the same output for every tile size; softmax
rows are probability vectors; the fused GEMM epilogue composes as bias -> ReLU -> residual."""
import numpy as np
from hypothesis import given, settings, strategies as st

settings.register_profile("oracle", max_examples=25, deadline=None)
settings.load_profile("oracle")


def _half(rng, shape, scale=1.0):
    return (rng.standard_normal(shape) * scale).astype(np.float16)


@given(seed=st.integers(0, 2 ** 31 - 1), nq=st.integers(1, 9), nk=st.integers(1, 70), d=st.sampled_from([8, 16, 64]),
       scale=st.floats(0.05, 1.0))
def test_attention_is_a_set_function_of_key_value_pairs_and_linear_in_v(oracle, seed, nq, nk, d, scale):
    rng = np.random.default_rng(seed)
    q, k, v = _half(rng, (2, nq, d)), _half(rng, (2, nk, d)), _half(rng, (2, nk, d))
    o = oracle.attention_f16(q, k, v, scale)
    perm = rng.permutation(nk)
    op = oracle.attention_f16(q, np.ascontiguousarray(k[:, perm]), np.ascontiguousarray(v[:, perm]), scale)
    assert np.allclose(o, op, rtol=2e-5, atol=2e-6)  # only the fp32 summation order changes
    v2 = _half(rng, (2, nk, d))
    lin = oracle.attention_f16(q, k, (v.astype(np.float32) * 0.5 + v2.astype(np.float32) * 0.25).astype(np.float16), scale)
    ref = 0.5 * o + 0.25 * oracle.attention_f16(q, k, v2, scale)
    assert np.allclose(lin, ref, rtol=0, atol=4e-3)  # the combined V is re-rounded to fp16
    # every output row is a convex combination of V rows
    assert (o <= v.astype(np.float32).max(axis=1, keepdims=True) + 1e-5).all()
    assert (o >= v.astype(np.float32).min(axis=1, keepdims=True) - 1e-5).all()


@given(seed=st.integers(0, 2 ** 31 - 1), nk=st.integers(2, 150), tile=st.sampled_from([1, 7, 16, 32, 128]))
def test_streaming_state_defines_the_same_output_for_every_tile_size(oracle, seed, nk, tile):
    rng = np.random.default_rng(seed)
    q, k, v = _half(rng, (1, 5, 16)), _half(rng, (1, nk, 16)), _half(rng, (1, nk, 16))
    m, l, y = oracle.streaming_state_f16(q, k, v, tile, 0.25)
    full = oracle.attention_f16(q, k, v, 0.25)
    assert np.allclose(y[-1] / l[-1][..., None], full, rtol=2e-5, atol=2e-6)
    assert (np.diff(m, axis=0) >= 0).all()  # the running max never decreases
    assert (l > 0).all()


@given(seed=st.integers(0, 2 ** 31 - 1), rows=st.integers(1, 12), n=st.integers(1, 300), spread=st.floats(0.1, 30.0))
def test_softmax_rows_are_probability_vectors_and_shift_invariant(oracle, seed, rows, n, spread):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((rows, n)) * spread).astype(np.float32)
    p = oracle.softmax_rows(x)
    assert (p >= 0).all() and np.allclose(p.sum(axis=1), 1.0, atol=1e-5)
    shifted = oracle.softmax_rows(x + np.float32(3.0))
    assert np.allclose(p, shifted, rtol=1e-4, atol=1e-7)  # exp(x - max) is shift-invariant up to the rounding of x + 3


@given(seed=st.integers(0, 2 ** 31 - 1), m=st.integers(1, 20), n=st.integers(1, 24), k=st.integers(1, 40))
def test_gemm_epilogue_composes_as_bias_relu_residual(oracle, seed, m, n, k):
    rng = np.random.default_rng(seed)
    a, b = rng.standard_normal((m, k)).astype(np.float32), rng.standard_normal((k, n)).astype(np.float32)
    bias, res = rng.standard_normal(n).astype(np.float32), rng.standard_normal((m, n)).astype(np.float32)
    full = oracle.gemm_fused(a, b, bias, res)
    no_res = oracle.gemm_fused(a, b, bias, None)
    assert np.array_equal(full, no_res + res)          # the residual is added after the ReLU, in fp32
    assert (no_res >= 0).all()                          # ReLU comes after the bias
    plain = oracle.gemm_batched(a[None], b[None])[0]    # same k-ascending fp32 contraction
    assert np.array_equal(no_res, np.maximum(plain + bias, 0))
