"""B200-native (sm_100a) FlashAttention-forward engine behind the operator surface of
jeehun98/FlashAttention_Kernel_Project: attention forward, GEMM + fused epilogue, row softmax.

The product is the C-ABI shared library ``libfa_b200.so`` (declared in ``include/fa_b200.h``);
this package is the thin host-side mirror used by the tests and the bench.  There is no CPU
fallback: importing works anywhere, but every compute call fails loudly unless the CUDA library
is built and an sm_100 GPU is present.
"""
from ._lib import FaError, lib_path, load_library  # noqa: F401
from .ops import (  # noqa: F401
    fa_forward,
    fa_forward_f32out,
    fa_forward_cross,
    fa_forward_partial,
    fa_forward_host,
    fa_forward_unfused,
    gemm_batched,
    gemm_fused_epilogue,
    gemm_fused_epilogue_f32,
    gemm_fused_epilogue_host,
    softmax_rows,
    softmax_rows_f16out,
    softmax_rows_host,
    launch_count,
    last_debug,
)
from .sharding import PeerShards, RingWorkspace, new_ring_group, ring_forward, ring_attention, ring_schedule, shard_range  # noqa: F401

__all__ = [
    "FaError", "lib_path", "load_library", "fa_forward", "fa_forward_f32out", "fa_forward_cross", "fa_forward_partial",
    "fa_forward_host", "gemm_batched", "fa_forward_unfused", "softmax_rows_f16out", "gemm_fused_epilogue", "gemm_fused_epilogue_f32", "gemm_fused_epilogue_host", "softmax_rows",
    "softmax_rows_host", "launch_count", "last_debug", "shard_range", "ring_schedule", "ring_attention", "PeerShards", "new_ring_group", "ring_forward", "RingWorkspace",
]
