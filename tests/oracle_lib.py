"""ctypes access to the CHECKER: oracle/libfa_oracle.so (C restatement of the reference's host
oracles) and, when present, oracle/_ref/libref_*.so (the reference's own functions and kernels,
compiled unmodified).  Test infrastructure only -- nothing under flashattention_kernel_project_b200/
imports this."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u16p = np.ctypeslib.ndpointer(dtype=np.uint16, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
c_int, c_float, c_size_t, c_double, c_ll = ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_double, ctypes.c_longlong


def _as_u16(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.float16:
        return a.view(np.uint16)
    assert a.dtype == np.uint16
    return a


def _opt_f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32).ctypes.data_as(ctypes.c_void_p)


class Oracle:
    """The C restatement (oracle/fa_oracle.c)."""

    def __init__(self):
        path = os.path.join(ORACLE_DIR, "libfa_oracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle"], check=True)
        L = ctypes.CDLL(path)
        L.fa_oracle_attention_f16.argtypes = [_u16p, _u16p, _u16p, _f32p] + [c_int] * 5 + [c_float, c_int]
        L.fa_oracle_attention_f16_rows.argtypes = [_u16p, _u16p, _u16p, _f32p, _i64p] + [c_int] * 6 + [c_float, c_int]
        L.fa_oracle_attention_f32_dacc.argtypes = [_f32p] * 4 + [c_int] * 3
        L.fa_oracle_attention_f32.argtypes = [_f32p] * 4 + [c_int] * 3 + [c_float]
        L.fa_oracle_softmax_rows.argtypes = [_f32p, _f32p, c_ll, c_int, c_int]
        L.fa_oracle_gemm_fused.argtypes = [_f32p, _f32p, ctypes.c_void_p, ctypes.c_void_p, _f32p] + [c_int] * 6
        L.fa_oracle_gemm_fused_f16.argtypes = [_u16p, _u16p, ctypes.c_void_p, ctypes.c_void_p, _f32p] + [c_int] * 6
        L.fa_oracle_gemm_fused_f16_act.argtypes = [_u16p, _u16p, ctypes.c_void_p, ctypes.c_void_p, _f32p] + [c_int] * 7
        L.fa_oracle_gemm_batched.argtypes = [_f32p, _f32p, _f32p] + [c_int] * 5
        L.fa_oracle_qk_scores.argtypes = [_f32p, _f32p, _f32p] + [c_int] * 4 + [c_float, c_int]
        L.fa_oracle_streaming_ref_f16.argtypes = [_u16p, _u16p, _u16p, _f32p] + [c_int] * 5 + [c_float]
        L.fa_oracle_streaming_state_f16.argtypes = [_u16p, _u16p, _u16p, _f32p, _f32p, _f32p] + [c_int] * 6 + [c_float]
        L.fa_oracle_rel_l2.argtypes = [_f32p, _f32p, c_size_t]
        L.fa_oracle_rel_l2.restype = c_double
        L.fa_oracle_max_abs.argtypes = [_f32p, _f32p, c_size_t]
        L.fa_oracle_max_abs.restype = c_double
        L.fa_oracle_half_to_float.argtypes = [_u16p, _f32p, c_size_t]
        L.fa_oracle_float_to_half.argtypes = [_f32p, _u16p, c_size_t]
        L.fa_oracle_num_threads.restype = c_int
        self.L = L

    def num_threads(self):
        return int(self.L.fa_oracle_num_threads())

    def attention_f16(self, q, k, v, scale, nthreads=0):
        """q [B, M, KDIM], k [B, L, KDIM], v [B, L, DV] fp16 -> fp32 [B, M, DV] (v12i cpu_reference)."""
        B, M, KD = q.shape
        Lk, DV = k.shape[1], v.shape[2]
        out = np.empty((B, M, DV), dtype=np.float32)
        self.L.fa_oracle_attention_f16(_as_u16(q), _as_u16(k), _as_u16(v), out, B, M, Lk, KD, DV, float(scale), nthreads)
        return out

    def attention_f16_rows(self, q, k, v, scale, rows, nthreads=0):
        """Sampled rows (flat ids b*M + m) -> fp32 [len(rows), DV]."""
        B, M, KD = q.shape
        Lk, DV = k.shape[1], v.shape[2]
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        out = np.empty((len(rows), DV), dtype=np.float32)
        self.L.fa_oracle_attention_f16_rows(_as_u16(q), _as_u16(k), _as_u16(v), out, rows, len(rows), B, M, Lk, KD, DV,
                                            float(scale), nthreads)
        return out

    def attention_f32_dacc(self, q, k, v):
        BH, N, D = q.shape
        out = np.empty((BH, N, D), dtype=np.float32)
        self.L.fa_oracle_attention_f32_dacc(*(np.ascontiguousarray(x, dtype=np.float32) for x in (q, k, v)), out, BH, N, D)
        return out

    def attention_f32(self, q, k, v, scale):
        BH, N, D = q.shape
        out = np.empty((BH, N, D), dtype=np.float32)
        self.L.fa_oracle_attention_f32(*(np.ascontiguousarray(x, dtype=np.float32) for x in (q, k, v)), out, N, D, BH,
                                       float(scale))
        return out

    def softmax_rows(self, x, nthreads=0):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        self.L.fa_oracle_softmax_rows(x, out, x.shape[0], x.shape[1], nthreads)
        return out

    def gemm_fused(self, a, b, bias, residual, row_begin=0, row_end=None, nthreads=0):
        """fp32 operands, full [M, N] result buffer; only rows [row_begin, row_end) are computed.  Restates both
        cpu_gemm_fused_ref functions of the reference (tensor-core file :136-155, fp32 file :116-136: same code)."""
        a = np.ascontiguousarray(a, dtype=np.float32)
        b = np.ascontiguousarray(b, dtype=np.float32)
        M, K = a.shape
        N = b.shape[1]
        row_end = M if row_end is None else row_end
        out = np.zeros((M, N), dtype=np.float32)
        self.L.fa_oracle_gemm_fused(a, b, _opt_f32(bias), _opt_f32(residual), out, M, N, K, row_begin, row_end, nthreads)
        return out

    def gemm_fused_f16(self, a, b, bias, residual, row_begin=0, row_end=None, nthreads=0, act="relu"):
        """fp16 operands (widened to fp32, as the GPU is fed); returns rows [row_begin, row_end).
        act: "relu" (the reference's epilogue) | "identity" | "gelu" | "silu" (extensions, parity-unpinned)."""
        M, K = a.shape
        N = b.shape[1]
        row_end = M if row_end is None else row_end
        out = np.empty((row_end - row_begin, N), dtype=np.float32)
        self.L.fa_oracle_gemm_fused_f16_act(_as_u16(a), _as_u16(b), _opt_f32(bias), _opt_f32(residual), out, M, N, K,
                                            row_begin, row_end, {"identity": 0, "relu": 1, "gelu": 2, "silu": 3}[act],
                                            nthreads)
        return out

    def gemm_batched(self, a, b, nthreads=0):
        """fp32 [batch,M,K] x [batch,K,N] -> [batch,M,N], fp32 accumulate (cpu_batched_gemm_ref)."""
        a = np.ascontiguousarray(a, dtype=np.float32)
        b = np.ascontiguousarray(b, dtype=np.float32)
        batch, M, K = a.shape
        N = b.shape[2]
        out = np.empty((batch, M, N), dtype=np.float32)
        self.L.fa_oracle_gemm_batched(a, b, out, batch, M, N, K, nthreads)
        return out

    def qk_scores(self, q, k, scale, nthreads=0):
        """fp32 q [BH,N,D], k [BH,Nk,D] -> scale * q k^T [BH,N,Nk], double accumulate (qk_cpu_ref)."""
        q = np.ascontiguousarray(q, dtype=np.float32)
        k = np.ascontiguousarray(k, dtype=np.float32)
        BH, N, D = q.shape
        Nk = k.shape[1]
        out = np.empty((BH, N, Nk), dtype=np.float32)
        self.L.fa_oracle_qk_scores(q, k, out, BH, N, Nk, D, float(scale), nthreads)
        return out

    def streaming_ref_f16(self, q, kt, v, scale):
        """q [B,M,Kd], kt [B,Kd,L] (transposed K, as the streaming family stores it), v [B,L,Dv] fp16 -> fp32 [B,M,Dv];
        `+1e-6` normaliser (flashattn_streaming_cpu_ref)."""
        B, M, Kd = q.shape
        L, Dv = v.shape[1], v.shape[2]
        out = np.empty((B, M, Dv), dtype=np.float32)
        self.L.fa_oracle_streaming_ref_f16(_as_u16(q), _as_u16(kt), _as_u16(v), out, B, M, L, Kd, Dv, float(scale))
        return out

    def streaming_state_f16(self, q, k, v, tile, scale):
        """Online-softmax state after every key tile: (m, l) [T,B,M], y [T,B,M,Dv]; q [B,M,D], k [B,L,D], v [B,L,Dv]."""
        B, M, D = q.shape
        L, Dv = k.shape[1], v.shape[2]
        T = (L + tile - 1) // tile
        m, l = np.empty((T, B, M), np.float32), np.empty((T, B, M), np.float32)
        y = np.empty((T, B, M, Dv), np.float32)
        self.L.fa_oracle_streaming_state_f16(_as_u16(q), _as_u16(k), _as_u16(v), m, l, y, B, M, L, D, Dv, tile, float(scale))
        return m, l, y

    def rel_l2(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.float32).ravel()
        b = np.ascontiguousarray(b, dtype=np.float32).ravel()
        return float(self.L.fa_oracle_rel_l2(a, b, a.size))

    def max_abs(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.float32).ravel()
        b = np.ascontiguousarray(b, dtype=np.float32).ravel()
        return float(self.L.fa_oracle_max_abs(a, b, a.size))

    def half_to_float(self, h):
        h = _as_u16(h).ravel()
        out = np.empty(h.size, dtype=np.float32)
        self.L.fa_oracle_half_to_float(h, out, h.size)
        return out

    def float_to_half(self, f):
        f = np.ascontiguousarray(f, dtype=np.float32).ravel()
        out = np.empty(f.size, dtype=np.uint16)
        self.L.fa_oracle_float_to_half(f, out, f.size)
        return out


def ref_available(name):
    return os.path.exists(os.path.join(REF_DIR, f"libref_{name}.so"))


class RefLib:
    """The reference's own code (oracle/_ref/libref_<name>.so), one library per reference file."""

    def __init__(self, name):
        path = os.path.join(REF_DIR, f"libref_{name}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.L = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
        self.name = name


class RefCpu:
    """Host-side functions of the reference (no GPU needed)."""

    def __init__(self):
        self.v12i = RefLib("v12i").L
        self.v12i.ref_v12i_cpu_reference.argtypes = [_u16p, _u16p, _u16p, _f32p] + [c_int] * 5 + [c_float]
        self.v12i.ref_v12i_rel_l2.argtypes = [_f32p, _f32p, c_size_t]
        self.v12i.ref_v12i_rel_l2.restype = c_double
        self.v12i.ref_v12i_fill_uniform.argtypes = [ctypes.c_uint, _u16p, c_size_t, _u16p, c_size_t, _u16p, c_size_t]
        self.v75 = RefLib("v7_5").L
        self.v75.ref_v7_5_cpu_ref.argtypes = [_u16p, _u16p, _u16p, _f32p, c_int, c_int]
        self.v75.ref_v7_5_fill.argtypes = [ctypes.c_uint, _u16p, c_size_t, _u16p, c_size_t, _u16p, c_size_t]
        self.fa542 = RefLib("fa542").L
        self.fa542.ref_fa542_cpu_ref.argtypes = [_f32p] * 4 + [c_int] * 3
        self.sb = RefLib("softmax_block").L
        self.sb.ref_softmax_block_flashattn_ref.argtypes = [_f32p] * 4 + [c_int] * 3 + [c_float]
        self.gemm = RefLib("gemm_fused").L
        self.gemm.ref_gemm_fused_cpu.argtypes = [_f32p] * 5 + [c_int] * 3
        self.gemm.ref_gemm_fused_fill.argtypes = [ctypes.c_uint] + [_f32p, c_size_t] * 4
        self.gemm32 = RefLib("gemm_fp32_fused").L if ref_available("gemm_fp32_fused") else None
        if self.gemm32 is not None:
            self.gemm32.ref_gemm_fp32_fused_cpu.argtypes = [_f32p] * 5 + [c_int] * 3
        # added after the first fixtures were cut: absent in an older prebuilt oracle/_ref
        self.bgemm = RefLib("gemm_batched").L if ref_available("gemm_batched") else None
        if self.bgemm is not None:
            self.bgemm.ref_gemm_batched_cpu.argtypes = [_f32p] * 3 + [c_int] * 4
        self.qk = RefLib("qk_wmma").L if ref_available("qk_wmma") else None
        if self.qk is not None:
            self.qk.ref_qk_cpu.argtypes = [_f32p] * 3 + [c_int] * 3 + [c_float]
        self.st16 = RefLib("streaming16").L if ref_available("streaming16") else None
        if self.st16 is not None:
            self.st16.ref_streaming16_cpu.argtypes = [_u16p, _u16p, _u16p, _f32p, c_int, c_int, c_float]
            self.st16.ref_streaming16_fill.argtypes = [ctypes.c_uint, _u16p, c_size_t, _u16p, c_size_t, _u16p, c_size_t]

    def v12i_cpu_reference(self, q, k, v, scale):
        B, M, KD = q.shape
        Lk, DV = k.shape[1], v.shape[2]
        out = np.empty((B, M, DV), dtype=np.float32)
        self.v12i.ref_v12i_cpu_reference(_as_u16(q), _as_u16(k), _as_u16(v), out, B, M, Lk, KD, DV, float(scale))
        return out

    def v12i_fill_uniform(self, seed, q_shape, k_shape, v_shape):
        q, k, v = (np.empty(s, dtype=np.uint16) for s in (q_shape, k_shape, v_shape))
        self.v12i.ref_v12i_fill_uniform(seed, q, q.size, k, k.size, v, v.size)
        return q.view(np.float16), k.view(np.float16), v.view(np.float16)

    def v12i_rel_l2(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.float32).ravel()
        b = np.ascontiguousarray(b, dtype=np.float32).ravel()
        return float(self.v12i.ref_v12i_rel_l2(a, b, a.size))

    def v7_5_cpu_ref(self, q, k, v):
        """q [B, 16], k, v [B, S, 16] fp16 -> fp32 [B, 16] (single query per batch element, scale 1/4)."""
        B, S = q.shape[0], k.shape[1]
        out = np.empty((B, 16), dtype=np.float32)
        self.v75.ref_v7_5_cpu_ref(_as_u16(q), _as_u16(k), _as_u16(v), out, B, S)
        return out

    def v7_5_fill(self, seed, B, S):
        q, k, v = np.empty((B, 16), np.uint16), np.empty((B, S, 16), np.uint16), np.empty((B, S, 16), np.uint16)
        self.v75.ref_v7_5_fill(seed, q, q.size, k, k.size, v, v.size)
        return q.view(np.float16), k.view(np.float16), v.view(np.float16)

    def fa542_cpu_ref(self, q, k, v):
        BH, N, D = q.shape
        out = np.empty((BH, N, D), dtype=np.float32)
        self.fa542.ref_fa542_cpu_ref(*(np.ascontiguousarray(x, dtype=np.float32) for x in (q, k, v)), out, BH, N, D)
        return out

    def softmax_block_flashattn_ref(self, q, k, v, scale):
        BH, N, D = q.shape
        out = np.empty((BH, N, D), dtype=np.float32)
        self.sb.ref_softmax_block_flashattn_ref(*(np.ascontiguousarray(x, dtype=np.float32) for x in (q, k, v)), out,
                                                N, D, BH, float(scale))
        return out

    def gemm_fp32_fused_cpu(self, a, b, bias, residual):
        """cpu_gemm_fused_ref of the fp32 sibling (GEMM/gemm_fused_epilogue/gemm_fused_epilogue.cu:116-136)."""
        M, K = a.shape
        N = b.shape[1]
        out = np.empty((M, N), dtype=np.float32)
        self.gemm32.ref_gemm_fp32_fused_cpu(*(np.ascontiguousarray(x, dtype=np.float32) for x in (a, b, bias, residual)),
                                            out, M, N, K)
        return out

    def gemm_fused_cpu(self, a, b, bias, residual):
        M, K = a.shape
        N = b.shape[1]
        out = np.empty((M, N), dtype=np.float32)
        self.gemm.ref_gemm_fused_cpu(*(np.ascontiguousarray(x, dtype=np.float32) for x in (a, b, bias, residual)), out,
                                     M, N, K)
        return out

    def gemm_batched_cpu(self, a, b):
        batch, M, K = a.shape
        N = b.shape[2]
        out = np.empty((batch, M, N), dtype=np.float32)
        self.bgemm.ref_gemm_batched_cpu(np.ascontiguousarray(a, dtype=np.float32), np.ascontiguousarray(b, dtype=np.float32),
                                        out, batch, M, N, K)
        return out

    def qk_cpu(self, q, k, scale):
        BH, N, D = q.shape
        out = np.empty((BH, N, N), dtype=np.float32)
        self.qk.ref_qk_cpu(np.ascontiguousarray(q, dtype=np.float32), np.ascontiguousarray(k, dtype=np.float32), out,
                           BH, N, D, float(scale))
        return out

    def streaming16_cpu(self, q, kt, v, scale):
        """q [B,16,16], kt [B,16,L], v [B,L,16] fp16 -> fp32 [B,16,16]."""
        B, L = q.shape[0], v.shape[1]
        out = np.empty((B, 16, 16), dtype=np.float32)
        self.st16.ref_streaming16_cpu(_as_u16(q), _as_u16(kt), _as_u16(v), out, B, L, float(scale))
        return out

    def streaming16_fill(self, seed, B, L):
        q, kt, v = np.empty((B, 16, 16), np.uint16), np.empty((B, 16, L), np.uint16), np.empty((B, L, 16), np.uint16)
        self.st16.ref_streaming16_fill(seed, q, q.size, kt, kt.size, v, v.size)
        return q.view(np.float16), kt.view(np.float16), v.view(np.float16)

    def gemm_fused_fill(self, seed, M, N, K):
        a, b = np.empty((M, K), np.float32), np.empty((K, N), np.float32)
        bias, res = np.empty((N,), np.float32), np.empty((M, N), np.float32)
        self.gemm.ref_gemm_fused_fill(seed, a, a.size, b, b.size, bias, bias.size, res, res.size)
        return a, b, bias, res


class RefGpu:
    """Launchers for the reference's own CUDA kernels (device pointers as ints)."""

    def __init__(self):
        vp = ctypes.c_void_p
        self.v5 = RefLib("wmma_v5").L
        self.v5.ref_wmma_v5_launch.argtypes = [vp, vp, vp, vp, c_int, c_int, c_int, c_float, vp]
        self.v4 = RefLib("wmma_v4").L
        self.v4.ref_wmma_v4_launch.argtypes = [vp, vp, vp, vp, c_int, c_int, c_int, c_float, vp]
        self.v12i = RefLib("v12i").L
        self.v12i.ref_v12i_launch.argtypes = [vp, vp, vp, vp, c_int, c_int, c_int, c_float, vp]
        self.v75 = RefLib("v7_5").L
        self.v75.ref_v7_5_launch.argtypes = [vp, vp, vp, vp, c_int, c_int, c_float, vp]
        self.gemm = RefLib("gemm_fused").L
        self.gemm.ref_gemm_fused_launch.argtypes = [vp, vp, vp, vp, vp, c_int, c_int, c_int, vp]
        self.gemm32 = RefLib("gemm_fp32_fused").L if ref_available("gemm_fp32_fused") else None
        if self.gemm32 is not None:
            self.gemm32.ref_gemm_fp32_fused_launch.argtypes = [vp, vp, vp, vp, vp, c_int, c_int, c_int, vp]
        self.softmax = RefLib("softmax_naive").L
        self.softmax.ref_softmax_naive_launch.argtypes = [vp, vp, c_int, c_int, vp]
        self.bgemm = RefLib("gemm_batched").L
        self.bgemm.ref_gemm_batched_launch.argtypes = [vp, vp, vp, c_int, c_int, c_int, c_int, vp]
        self.qk = RefLib("qk_wmma").L
        self.qk.ref_qk_launch.argtypes = [vp, vp, vp, c_int, c_int, c_int, c_float, vp]
        self.st16 = RefLib("streaming16").L
        self.st16.ref_streaming16_launch.argtypes = [vp, vp, vp, vp, c_int, c_int, c_float, vp]
        self.mw = {}
        for ver in ("v10", "v11"):
            L = RefLib("mw_" + ver).L
            getattr(L, f"ref_mw_{ver}_launch").argtypes = [vp, vp, vp, vp, c_int, c_int, c_float, vp]
            self.mw[ver] = getattr(L, f"ref_mw_{ver}_launch")
        self.sm_hybrid = RefLib("softmax_hybrid").L
        self.sm_hybrid.ref_softmax_hybrid_launch.argtypes = [vp, vp, c_int, c_int, vp]
        self.sm_warp = RefLib("softmax_warp").L
        self.sm_warp.ref_softmax_warp_launch.argtypes = [vp, vp, c_int, c_int, vp]

    @staticmethod
    def _chk(rc, what):
        if rc != 0:
            raise RuntimeError(f"reference launcher {what} failed with cudaError {rc}")

    def wmma_v5(self, q, k, v, out, scale, stream=0):
        BH, N, D = q.shape[0] * q.shape[1], q.shape[2], q.shape[3]
        self._chk(self.v5.ref_wmma_v5_launch(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), BH, N, D,
                                             float(scale), stream), "wmma_v5")

    def wmma_v4(self, q, k, v, out, scale, stream=0):
        BH, N, D = q.shape[0] * q.shape[1], q.shape[2], q.shape[3]
        self._chk(self.v4.ref_wmma_v4_launch(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), BH, N, D,
                                             float(scale), stream), "wmma_v4")

    def v12i_kernel(self, q, k, v, out, scale, stream=0):
        B, M, L = q.shape[0], q.shape[1], k.shape[1]
        self._chk(self.v12i.ref_v12i_launch(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, M, L,
                                            float(scale), stream), "v12i")

    def v7_5_kernel(self, q, k, v, out, stream=0):
        B, S = q.shape[0], k.shape[1]
        self._chk(self.v75.ref_v7_5_launch(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, S, 0.25, stream),
                  "v7_5")

    def gemm_fused(self, a, b, bias, residual, out, stream=0):
        M, K = a.shape
        N = b.shape[1]
        self._chk(self.gemm.ref_gemm_fused_launch(a.data_ptr(), b.data_ptr(), bias.data_ptr() if bias is not None else None,
                                                  residual.data_ptr() if residual is not None else None, out.data_ptr(),
                                                  M, N, K, stream), "gemm_fused")

    def gemm_fp32_fused(self, a, b, bias, residual, out, stream=0):
        """gemm_tiled_fused_kernel (fp32 operands) on device tensors."""
        M, K = a.shape
        N = b.shape[1]
        self._chk(self.gemm32.ref_gemm_fp32_fused_launch(a.data_ptr(), b.data_ptr(), bias.data_ptr() if bias is not None else None,
                                                         residual.data_ptr() if residual is not None else None,
                                                         out.data_ptr(), M, N, K, stream), "gemm_tiled_fused_kernel")

    def softmax_naive(self, x, out, stream=0):
        self._chk(self.softmax.ref_softmax_naive_launch(x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], stream),
                  "softmax_naive")

    def gemm_batched(self, a, b, out, stream=0):
        batch, M, K = a.shape
        N = b.shape[2]
        self._chk(self.bgemm.ref_gemm_batched_launch(a.data_ptr(), b.data_ptr(), out.data_ptr(), batch, M, N, K, stream),
                  "gemm_batched")

    def qk_wmma(self, q, k, out, scale, stream=0):
        BH, N, D = q.shape
        self._chk(self.qk.ref_qk_launch(q.data_ptr(), k.data_ptr(), out.data_ptr(), BH, N, D, float(scale), stream),
                  "qk_wmma")

    def streaming16(self, q, kt, v, out, scale, stream=0):
        """q [B,16,16], kt [B,16,L] (transposed K), v [B,L,16] fp16 device tensors -> out fp32 [B,16,16]."""
        self._chk(self.st16.ref_streaming16_launch(q.data_ptr(), kt.data_ptr(), v.data_ptr(), out.data_ptr(), q.shape[0],
                                                   v.shape[1], float(scale), stream), "streaming16")

    def softmax_block_hybrid(self, x, out, stream=0):
        """fp32 [rows, N <= 512] device tensors (softmax_block_hybrid_kernel, `__expf`)."""
        self._chk(self.sm_hybrid.ref_softmax_hybrid_launch(x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], stream),
                  "softmax_block_hybrid")

    def softmax_warp_shfl(self, x, out, stream=0):
        """fp32 [rows, N <= 32] device tensors (softmax_warp_shfl_kernel)."""
        self._chk(self.sm_warp.ref_softmax_warp_launch(x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], stream),
                  "softmax_warp_shfl")

    def streaming_mw(self, ver, q, k, v, out, scale, stream=0):
        """flashattn_streaming_16x16_kernel_mw_v10 / _v11: q [B,16,16], k [B,L,16] (row per key), v [B,L,16] fp16 device
        tensors -> out fp32 [B,16,16]."""
        self._chk(self.mw[ver](q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), q.shape[0], k.shape[1],
                               float(scale), stream), "mw_" + ver)
