"""Regenerates tests/golden/*.npz from the reference's OWN code (oracle/_ref/libref_*.so, i.e. the
unmodified reference sources compiled through oracle/ref_wrappers/).  Run in the build container,
where /root/reference is mounted:   python tests/golden/make_golden.py
Inputs come from the reference's seeded generators (mt19937, uniform(-1,1)) where it has one.
The fixtures are what the oracle and the CUDA path are compared against on boxes that have neither
the reference tree nor oracle/_ref.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.oracle_lib import RefCpu  # noqa: E402


def main():
    r = RefCpu()
    # BASELINE config 1: B=1, H=1, N=128, D=64, mt19937(123) uniform(-1,1), order Q,K,V, scale 1/8
    # (v12i.cu:317-320 generator, :243-286 oracle)
    q, k, v = r.v12i_fill_uniform(123, (1, 128, 64), (1, 128, 64), (1, 128, 64))
    o = r.v12i_cpu_reference(q, k, v, 0.125)
    np.savez_compressed(os.path.join(HERE, "attn_cfg1_v12i.npz"), q=q.view(np.uint16), k=k.view(np.uint16),
                        v=v.view(np.uint16), o=o, scale=np.float32(0.125))
    # v12i's native shape, reduced batch: B=8, M=16, L=128, KDIM=DV=16, scale 1/4 (v12i.cu:299-307)
    q, k, v = r.v12i_fill_uniform(123, (8, 16, 16), (8, 128, 16), (8, 128, 16))
    o = r.v12i_cpu_reference(q, k, v, 0.25)
    np.savez_compressed(os.path.join(HERE, "attn_v12i_native.npz"), q=q.view(np.uint16), k=k.view(np.uint16),
                        v=v.view(np.uint16), o=o, scale=np.float32(0.25))
    # v7_5's native shape, reduced batch: one query per batch element, S=128, D=16, mt19937(0), scale 1/4
    # (flashattn_warp_spc_v2/..v7_5..cu:250-283 oracle, :297-303 generator)
    q, k, v = r.v7_5_fill(0, 32, 128)
    np.savez_compressed(os.path.join(HERE, "attn_v7_5_native.npz"), q=q.view(np.uint16), k=k.view(np.uint16),
                        v=v.view(np.uint16), o=r.v7_5_cpu_ref(q, k, v), scale=np.float32(0.25))
    # ragged D=128 case (N=200: one full and one partial 128-key tile), 2 (b,h) pairs
    q, k, v = r.v12i_fill_uniform(7, (2, 200, 128), (2, 200, 128), (2, 200, 128))
    sc = np.float32(1.0 / np.sqrt(128.0))
    o = r.v12i_cpu_reference(q, k, v, sc)
    np.savez_compressed(os.path.join(HERE, "attn_ragged_d128.npz"), q=q.view(np.uint16), k=k.view(np.uint16),
                        v=v.view(np.uint16), o=o, scale=sc)
    # fp32 oracles: 5.4.2 (double accumulate) and flashattn_ref (fp32), BH=2, N=64, D=32
    rng = np.random.default_rng(542)
    qf, kf, vf = (rng.standard_normal((2, 64, 32)).astype(np.float32) for _ in range(3))
    np.savez_compressed(os.path.join(HERE, "attn_f32_oracles.npz"), q=qf, k=kf, v=vf,
                        o_542=r.fa542_cpu_ref(qf, kf, vf), o_block=r.softmax_block_flashattn_ref(qf, kf, vf, 0.3),
                        scale_block=np.float32(0.3))
    # fused GEMM: mt19937(42), uniform(-1,1), order A,B,bias,residual (tensorcore_gemm_fused_epilogue.cu:178-184)
    a, b, bias, res = r.gemm_fused_fill(42, 48, 64, 80)
    np.savez_compressed(os.path.join(HERE, "gemm_fused.npz"), a=a, b=b, bias=bias, residual=res,
                        c=r.gemm_fused_cpu(a, b, bias, res))
    # batched plain GEMM (gemm_4_9_tensorcore_batched.cu:94-121) and scaled QK^T (qk_wmma_5_1_2.cu:101-122) on
    # half-rounded operands widened to fp32, as both harnesses feed their CPU reference
    rng2 = np.random.default_rng(49)
    ab = rng2.uniform(-1, 1, (3, 32, 48)).astype(np.float16).astype(np.float32)
    bb = rng2.uniform(-1, 1, (3, 48, 40)).astype(np.float16).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "gemm_batched.npz"), a=ab, b=bb, c=r.gemm_batched_cpu(ab, bb))
    qq = rng2.uniform(-1, 1, (2, 48, 64)).astype(np.float16).astype(np.float32)
    kk = rng2.uniform(-1, 1, (2, 48, 64)).astype(np.float16).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "qk_scores.npz"), q=qq, k=kk, s=r.qk_cpu(qq, kk, 0.125), scale=np.float32(0.125))
    # the streaming family's fixture and oracle: mt19937(42) normal(0,1), order Q,K,V, K stored [B,16,L], `+1e-6`
    # normaliser, scale 1/4 (Streaming_FlashAttention_Forward_Kernel/flashattn_streaming_16x16.cu:280-348,365-380); B reduced
    q, kt, v = r.streaming16_fill(42, 24, 128)
    np.savez_compressed(os.path.join(HERE, "attn_streaming16.npz"), q=q.view(np.uint16), kt=kt.view(np.uint16),
                        v=v.view(np.uint16), o=r.streaming16_cpu(q, kt, v, 0.25), scale=np.float32(0.25))
    # row softmax through flashattn_ref with K = V = I, scale = 1 (its softmax step, :242-252)
    n = 48
    x = (rng.standard_normal((1, n, n)) * 3).astype(np.float32)
    eye = np.eye(n, dtype=np.float32)[None]
    np.savez_compressed(os.path.join(HERE, "softmax_rows.npz"), x=x[0], p=r.softmax_block_flashattn_ref(x, eye, eye, 1.0)[0])
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
