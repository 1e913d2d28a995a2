"""Host-side mirror of the reference's operator surface, one function per C-ABI entry point.

Tensors are torch CUDA tensors (torch is plumbing here: device memory and streams); the *_host
variants take numpy arrays / pinned torch CPU tensors and run the reference's main() flow
(H2D -> kernel -> D2H) inside the library.  Argument meaning follows the reference kernels:
flat (B*H) batches of row-major [N, D] matrices, explicit `scale`, epilogue order
bias -> ReLU -> residual with nullable bias / residual.
"""
import ctypes
import math

import numpy as np
import torch

from ._lib import check, load_library


class _on(object):
    """`with _on(device):` -- make `device` current for the call, at no cost when it already is (torch.cuda.device
    is a ~5 us context manager; BASELINE config 1 is 4 us of GPU time)."""
    __slots__ = ("ctx",)

    def __init__(self, device):
        self.ctx = None if device.index is None or device.index == torch.cuda.current_device() else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


def _stream_ptr(stream=None):
    s = torch.cuda.current_stream() if stream is None else stream
    return ctypes.c_void_p(s.cuda_stream)


def _dev_ptr(t, dtype, name, shape=None, device=None):
    """Device address of `t` after checking type, dtype, layout and -- when given -- the exact shape and the device
    (an undersized or foreign buffer would be written past its end by a TMA / float4 store)."""
    if t is None:
        return None
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor (no CPU fallback exists)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous (row-major)")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name} must have shape {tuple(shape)}, got {tuple(t.shape)}")
    if device is not None and t.device != device:
        raise ValueError(f"{name} must be on {device}, got {t.device}")
    return ctypes.c_void_p(t.data_ptr())


def _bhnd(q):
    if not isinstance(q, torch.Tensor) or not q.is_cuda:
        raise TypeError("expected a CUDA tensor (no CPU fallback exists)")
    if q.dim() != 4:
        raise ValueError("expected a [B, H, N, D] tensor")
    return tuple(int(x) for x in q.shape)


def _require_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor (no CPU fallback exists)")


def fa_forward(q, k, v, scale=None, out=None, stream=None):
    """O = softmax(scale * Q K^T) V, fp16 [B,H,N,D] in and out (include/fa_b200.h: fa_forward)."""
    B, H, N, D = _bhnd(q)
    if tuple(k.shape) != (B, H, N, D) or tuple(v.shape) != (B, H, N, D):
        raise ValueError("Q, K, V must share the shape [B, H, N, D]")
    if scale is None:
        scale = 1.0 / float(D) ** 0.5  # the reference's default (v12i.cu:307)
    if out is None:
        out = torch.empty_like(q)
    lib = load_library()
    with _on(q.device):
        dev, shp = q.device, (B, H, N, D)
        check(lib.fa_forward(_dev_ptr(q, torch.float16, "Q"), _dev_ptr(k, torch.float16, "K", shp, dev),
                             _dev_ptr(v, torch.float16, "V", shp, dev), _dev_ptr(out, torch.float16, "O", shp, dev),
                             B, H, N, D, float(scale), _stream_ptr(stream)), "fa_forward")
    return out


def fa_forward_f32out(q, k, v, scale=None, out=None, stream=None):
    """Same with the reference kernels' fp32 output (flashattn_forward_wmma.cu:49-58)."""
    B, H, N, D = _bhnd(q)
    if tuple(k.shape) != (B, H, N, D) or tuple(v.shape) != (B, H, N, D):
        raise ValueError("Q, K, V must share the shape [B, H, N, D]")
    if scale is None:
        scale = 1.0 / float(D) ** 0.5
    if out is None:
        out = torch.empty((B, H, N, D), dtype=torch.float32, device=q.device)
    lib = load_library()
    with _on(q.device):
        dev, shp = q.device, (B, H, N, D)
        check(lib.fa_forward_f32out(_dev_ptr(q, torch.float16, "Q"), _dev_ptr(k, torch.float16, "K", shp, dev),
                                    _dev_ptr(v, torch.float16, "V", shp, dev), _dev_ptr(out, torch.float32, "O", shp, dev),
                                    B, H, N, D, float(scale), _stream_ptr(stream)), "fa_forward_f32out")
    return out


def fa_forward_cross(q, k, v, scale=None, out=None, stream=None):
    """Attention with Nq != Nkv and fp32 output; D in {16, 64, 128}.  D = 16 is the native shape of the
    reference's v12i (Q [B,M,16], K,V [B,L,16]) and v7_5 (one query per batch element) programs."""
    B, H, Nq, D = _bhnd(q)
    Bk, Hk, Nkv, Dk = _bhnd(k)
    if (Bk, Hk, Dk) != (B, H, D) or tuple(v.shape) != tuple(k.shape):
        raise ValueError("K, V must be [B, H, Nkv, D] matching Q's B, H, D")
    if scale is None:
        scale = 1.0 / float(D) ** 0.5
    if out is None:
        out = torch.empty((B, H, Nq, D), dtype=torch.float32, device=q.device)
    lib = load_library()
    with _on(q.device):
        dev, kshp = q.device, (B, H, Nkv, D)
        check(lib.fa_forward_cross(_dev_ptr(q, torch.float16, "Q"), _dev_ptr(k, torch.float16, "K", kshp, dev),
                                   _dev_ptr(v, torch.float16, "V", kshp, dev),
                                   _dev_ptr(out, torch.float32, "O", (B, H, Nq, D), dev),
                                   B, H, Nq, Nkv, D, float(scale), _stream_ptr(stream)), "fa_forward_cross")
    return out


def fa_forward_partial(q, k, v, o_acc, m, l, scale, is_first, is_last, o_final=None, stream=None):
    """One ring step: fold attention of Q against one K/V shard into the carried (O_acc, m, l)."""
    B, H, Nq, D = _bhnd(q)
    Bk, Hk, Nkv, Dk = _bhnd(k)
    if (Bk, Hk, Dk) != (B, H, D) or tuple(v.shape) != tuple(k.shape):
        raise ValueError("K, V must be [B, H, Nkv, D] matching Q's B, H, D")
    if is_last and o_final is None:
        o_final = torch.empty_like(q)
    lib = load_library()
    with _on(q.device):
        dev, kshp = q.device, (B, H, Nkv, D)
        check(lib.fa_forward_partial(_dev_ptr(q, torch.float16, "Q"), _dev_ptr(k, torch.float16, "K", kshp, dev),
                                     _dev_ptr(v, torch.float16, "V", kshp, dev),
                                     _dev_ptr(o_acc, torch.float32, "O_acc", (B, H, Nq, D), dev),
                                     _dev_ptr(m, torch.float32, "m", (B, H, Nq), dev),
                                     _dev_ptr(l, torch.float32, "l", (B, H, Nq), dev),
                                     B, H, Nq, Nkv, D, float(scale), int(bool(is_first)), int(bool(is_last)),
                                     _dev_ptr(o_final, torch.float16, "O_final", (B, H, Nq, D), dev) if is_last else None,
                                     _stream_ptr(stream)), "fa_forward_partial")
    return o_final


ACTIVATIONS = {"identity": 0, "relu": 1, "gelu": 2, "silu": 3}  # FA_ACT_* of include/fa_b200.h


def gemm_fused_epilogue(a, b, bias=None, residual=None, out=None, stream=None, act="relu"):
    """C = act(A @ B + bias) + residual; A fp16 [M,K], B fp16 [K,N], fp32 bias/residual/C.
    act = "relu" is the reference's contract (tensorcore_gemm_fused_epilogue.cu:109-126); "identity", "gelu" (exact,
    erf form) and "silu" go through gemm_fused_epilogue_act."""
    if act not in ACTIVATIONS:
        raise ValueError(f"act must be one of {sorted(ACTIVATIONS)}")
    _require_cuda(a, "A")
    _require_cuda(b, "B")
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != b.shape[0]:
        raise ValueError("A must be [M, K] and B [K, N]")
    M, K = int(a.shape[0]), int(a.shape[1])
    N = int(b.shape[1])
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    if bias is not None and tuple(bias.shape) != (N,):
        raise ValueError("bias must be [N]")
    if residual is not None and tuple(residual.shape) != (M, N):
        raise ValueError("residual must be [M, N]")
    lib = load_library()
    with _on(a.device):
        dev = a.device
        ptrs = (_dev_ptr(a, torch.float16, "A"), _dev_ptr(b, torch.float16, "B", (K, N), dev),
                _dev_ptr(bias, torch.float32, "bias", (N,), dev),
                _dev_ptr(residual, torch.float32, "residual", (M, N), dev),
                _dev_ptr(out, torch.float32, "C", (M, N), dev))
        if act == "relu":
            check(lib.gemm_fused_epilogue(*ptrs, M, N, K, _stream_ptr(stream)), "gemm_fused_epilogue")
        else:
            check(lib.gemm_fused_epilogue_act(*ptrs, M, N, K, ACTIVATIONS[act], _stream_ptr(stream)),
                  "gemm_fused_epilogue_act")
    return out


def gemm_fused_epilogue_f32(a, b, bias=None, residual=None, out=None, stream=None, act="relu"):
    """C = act(A @ B + bias) + residual with fp32 operands: A [M,K], B [K,N] (gemm_tiled_fused_kernel,
    GEMM/gemm_fused_epilogue/gemm_fused_epilogue.cu:44-111), on the tensor cores as split tf32 pairs (3xTF32)."""
    _require_cuda(a, "A")
    _require_cuda(b, "B")
    if act not in ACTIVATIONS:
        raise ValueError(f"act must be one of {sorted(ACTIVATIONS)}")
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != b.shape[0]:
        raise ValueError("A must be [M, K] and B [K, N]")
    M, K = int(a.shape[0]), int(a.shape[1])
    N = int(b.shape[1])
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    lib = load_library()
    with _on(a.device):
        dev = a.device
        check(lib.gemm_fused_epilogue_f32(_dev_ptr(a, torch.float32, "A"), _dev_ptr(b, torch.float32, "B", (K, N), dev),
                                          _dev_ptr(bias, torch.float32, "bias", (N,), dev),
                                          _dev_ptr(residual, torch.float32, "residual", (M, N), dev),
                                          _dev_ptr(out, torch.float32, "C", (M, N), dev), M, N, K, ACTIVATIONS[act],
                                          _stream_ptr(stream)), "gemm_fused_epilogue_f32")
    return out


def _strided3(t, dtype, name, device=None):
    """(pointer, row pitch, batch stride) of a [batch, rows, cols] tensor whose last dimension is contiguous: a dense
    tensor or a view of one (a head of a [B,N,H,D] tensor, a sub-matrix, a padded leading dimension)."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor (no CPU fallback exists)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if device is not None and t.device != device:
        raise ValueError(f"{name} must be on {device}, got {t.device}")
    if t.dim() != 3 or (t.shape[2] > 1 and t.stride(2) != 1):
        raise ValueError(f"{name} must be [batch, rows, cols] with a contiguous last dimension")
    ld = int(t.stride(1)) if t.shape[1] > 1 else int(t.shape[2])
    sb = int(t.stride(0)) if t.shape[0] > 1 else ld * int(t.shape[1])
    if ld < t.shape[2] or sb <= 0:
        raise ValueError(f"{name}: overlapping or broadcast strides are not supported")
    return ctypes.c_void_p(t.data_ptr()), ld, sb


def gemm_batched(a, b, trans_b=False, alpha=1.0, out=None, stream=None):
    """C[z] = alpha * A[z] @ op(B[z]); A fp16 [batch,M,K]; B fp16 [batch,K,N], or [batch,N,K] with trans_b; C fp32.
    Dense tensors go through gemm_batched, views with a contiguous last dimension (row pitch / batch stride different
    from the packed ones) through gemm_batched_strided -- same kernel, the pitches go into the TMA descriptors.

    trans_b=False, alpha=1: wmma_batched_gemm_kernel (gemm_4_9_tensorcore_batched.cu:42-46);
    trans_b=True, alpha=scale: QK_wmma_kernel (qk_wmma_5_1_2.cu:32-38)."""
    _require_cuda(a, "A")
    _require_cuda(b, "B")
    if a.dim() != 3 or b.dim() != 3 or a.shape[0] != b.shape[0]:
        raise ValueError("A must be [batch, M, K] and B [batch, K, N] (or [batch, N, K] with trans_b)")
    batch, M, K = (int(v) for v in a.shape)
    N = int(b.shape[1] if trans_b else b.shape[2])
    if int(b.shape[2] if trans_b else b.shape[1]) != K:
        raise ValueError("inner extents of A and B differ")
    if out is None:
        out = torch.empty((batch, M, N), dtype=torch.float32, device=a.device)
    elif tuple(out.shape) != (batch, M, N):
        raise ValueError("C must be [batch, M, N]")
    lib = load_library()
    with _on(a.device):
        if a.is_contiguous() and b.is_contiguous() and out.is_contiguous():
            check(lib.gemm_batched(_dev_ptr(a, torch.float16, "A"), _dev_ptr(b, torch.float16, "B", None, a.device),
                                   _dev_ptr(out, torch.float32, "C", (batch, M, N), a.device), batch, M, N, K,
                                   1 if trans_b else 0, float(alpha), _stream_ptr(stream)), "gemm_batched")
        else:
            pa, lda, sa = _strided3(a, torch.float16, "A")
            pb, ldb, sb = _strided3(b, torch.float16, "B", a.device)
            pc, ldc, sc = _strided3(out, torch.float32, "C", a.device)
            check(lib.gemm_batched_strided(pa, pb, pc, batch, M, N, K, lda, ldb, ldc, sa, sb, sc, 1 if trans_b else 0,
                                           float(alpha), _stream_ptr(stream)), "gemm_batched_strided")
    return out


def softmax_rows(x, out=None, stream=None):
    """Row softmax of an fp32 [rows, N] matrix (softmax_naive_baseline.cu:56-61)."""
    _require_cuda(x, "X")
    if x.dim() != 2:
        raise ValueError("expected a [rows, N] tensor")
    rows, N = int(x.shape[0]), int(x.shape[1])
    if out is None:
        out = torch.empty_like(x)
    lib = load_library()
    with _on(x.device):
        check(lib.softmax_rows(_dev_ptr(x, torch.float32, "X"), _dev_ptr(out, torch.float32, "P", (rows, N), x.device),
                               rows, N, _stream_ptr(stream)), "softmax_rows")
    return out


def softmax_rows_f16out(x, out=None, stream=None):
    """Row softmax with fp16 output (feeds the P*V GEMM of the unfused pipeline)."""
    _require_cuda(x, "X")
    if x.dim() != 2:
        raise ValueError("expected a [rows, N] tensor")
    rows, N = int(x.shape[0]), int(x.shape[1])
    if out is None:
        out = torch.empty((rows, N), dtype=torch.float16, device=x.device)
    lib = load_library()
    with _on(x.device):
        check(lib.softmax_rows_f16out(_dev_ptr(x, torch.float32, "X"),
                                      _dev_ptr(out, torch.float16, "P", (rows, N), x.device), rows, N,
                                      _stream_ptr(stream)), "softmax_rows_f16out")
    return out


def fa_forward_unfused(q, k, v, scale=None, workspace=None, stream=None):
    """Attention through materialised scores (QK^T GEMM -> row softmax -> PV GEMM); fp32 [B,H,N,D].

    workspace: optional (S fp32 [B*H,N,N], P fp16 [B*H,N,N]) pair to reuse; returned as the second value."""
    _require_cuda(q, "Q")
    if q.dim() != 4 or k.shape != q.shape or v.shape != q.shape:
        raise ValueError("Q, K, V must be [B, H, N, D] of one shape")
    B, H, N, D = (int(x) for x in q.shape)
    if scale is None:
        scale = 1.0 / math.sqrt(D)
    out = torch.empty((B, H, N, D), dtype=torch.float32, device=q.device)
    if workspace is None:
        workspace = (torch.empty((B * H, N, N), dtype=torch.float32, device=q.device),
                     torch.empty((B * H, N, N), dtype=torch.float16, device=q.device))
    S, P = workspace
    if tuple(S.shape) != (B * H, N, N) or tuple(P.shape) != (B * H, N, N):
        raise ValueError("workspace must be (fp32 [B*H,N,N], fp16 [B*H,N,N])")
    lib = load_library()
    with _on(q.device):
        check(lib.fa_forward_unfused(_dev_ptr(q, torch.float16, "Q"), _dev_ptr(k, torch.float16, "K"),
                                     _dev_ptr(v, torch.float16, "V"), _dev_ptr(out, torch.float32, "O"),
                                     _dev_ptr(S, torch.float32, "S"), _dev_ptr(P, torch.float16, "P"), B, H, N, D,
                                     float(scale), _stream_ptr(stream)), "fa_forward_unfused")
    return out, workspace


# ----------------------------- host-buffer variants ------------------------------------
_NP_TO_TORCH = {np.dtype(np.float16): torch.float16, np.dtype(np.float32): torch.float32}


def _host_ptr(a, dtype, name, shape=None):
    """Host address of `a` (numpy array or torch CPU tensor) after checking dtype, layout and the exact shape: the
    library copies extents derived from the FIRST argument's shape, so a short or mistyped buffer would be read or
    written past its end inside cudaMemcpyAsync."""
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        if a.is_cuda:
            raise TypeError(f"{name}: the *_host entry points take host memory")
        if a.dtype != _NP_TO_TORCH[np.dtype(dtype)]:
            raise TypeError(f"{name} must be {_NP_TO_TORCH[np.dtype(dtype)]}, got {a.dtype}")
        if not a.is_contiguous():
            raise ValueError(f"{name} must be contiguous")
        ptr = a.data_ptr()
    elif isinstance(a, np.ndarray):
        if a.dtype != np.dtype(dtype) or not a.flags["C_CONTIGUOUS"]:
            raise TypeError(f"{name} must be a C-contiguous numpy array of {np.dtype(dtype)}")
        ptr = a.ctypes.data
    else:
        raise TypeError(f"{name} must be a numpy array or a torch CPU tensor")
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"{name} must have shape {tuple(shape)}, got {tuple(a.shape)}")
    return ctypes.c_void_p(ptr)


def fa_forward_host(q, k, v, out, scale=None):
    """fa_forward on HOST buffers (fp16 [B,H,N,D]); copies are part of the call."""
    if getattr(q, "ndim", 0) != 4:
        raise ValueError("Q must be [B, H, N, D]")
    B, H, N, D = (int(x) for x in q.shape)
    shp = (B, H, N, D)
    if scale is None:
        scale = 1.0 / float(D) ** 0.5
    lib = load_library()
    check(lib.fa_forward_host(_host_ptr(q, np.float16, "Q", shp), _host_ptr(k, np.float16, "K", shp),
                              _host_ptr(v, np.float16, "V", shp), _host_ptr(out, np.float16, "O", shp),
                              B, H, N, D, float(scale)), "fa_forward_host")
    return out


def gemm_fused_epilogue_host(a, b, bias, residual, out):
    if getattr(a, "ndim", 0) != 2 or getattr(b, "ndim", 0) != 2 or a.shape[1] != b.shape[0]:
        raise ValueError("A must be [M, K] and B [K, N]")
    M, K = (int(x) for x in a.shape)
    N = int(b.shape[1])
    lib = load_library()
    check(lib.gemm_fused_epilogue_host(_host_ptr(a, np.float16, "A", (M, K)), _host_ptr(b, np.float16, "B", (K, N)),
                                       _host_ptr(bias, np.float32, "bias", (N,)),
                                       _host_ptr(residual, np.float32, "residual", (M, N)),
                                       _host_ptr(out, np.float32, "C", (M, N)), M, N, K), "gemm_fused_epilogue_host")
    return out


def softmax_rows_host(x, out):
    if getattr(x, "ndim", 0) != 2:
        raise ValueError("X must be [rows, N]")
    rows, N = (int(v) for v in x.shape)
    lib = load_library()
    check(lib.softmax_rows_host(_host_ptr(x, np.float32, "X", (rows, N)), _host_ptr(out, np.float32, "P", (rows, N)),
                                rows, N), "softmax_rows_host")
    return out


def launch_count():
    """Kernels launched by the library in this process."""
    return int(load_library().fa_launch_count())


def last_debug():
    """The hang-diagnostic record (list of ints), empty if none was written."""
    buf = (ctypes.c_uint * 8)()
    n = load_library().fa_last_debug(buf, 8)
    words = [int(buf[i]) for i in range(n)]
    return words if words and words[0] == 0xDEAD0001 else []
