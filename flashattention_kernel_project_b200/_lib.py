"""ctypes binding of libfa_b200.so -- exactly the symbols include/fa_b200.h declares."""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libfa_b200.so"
_lock = threading.Lock()
_lib = None

c_void_p, c_int, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
c_float_p = ctypes.c_void_p  # device / host addresses are passed as integers

# name -> (restype, argtypes); kept in sync with include/fa_b200.h (tests/test_abi.py checks it)
SIGNATURES = {
    "fa_version": (c_int, []),
    "fa_init": (c_int, []),
    "fa_strerror": (ctypes.c_char_p, [c_int]),
    "fa_forward": (c_int, [c_void_p] * 4 + [c_int] * 4 + [c_float, c_void_p]),
    "fa_forward_f32out": (c_int, [c_void_p] * 4 + [c_int] * 4 + [c_float, c_void_p]),
    "fa_forward_cross": (c_int, [c_void_p] * 4 + [c_int] * 5 + [c_float, c_void_p]),
    "fa_forward_partial": (c_int, [c_void_p] * 6 + [c_int] * 5 + [c_float, c_int, c_int, c_void_p, c_void_p]),
    "fa_ring_workspace_bytes": (ctypes.c_size_t, [c_int] * 6),
    "fa_ring_forward": (c_int, [c_void_p] * 4 + [c_int] * 4 + [c_float, c_int, c_int, c_void_p, c_void_p, ctypes.c_size_t, c_void_p]),
    "fa_enable_peer_access": (c_int, [c_int]),
    "fa_ipc_export": (c_int, [c_void_p, ctypes.POINTER(ctypes.c_ubyte), ctypes.POINTER(ctypes.c_ulonglong)]),
    "fa_ipc_open": (c_int, [ctypes.POINTER(ctypes.c_ubyte), ctypes.POINTER(c_void_p)]),
    "fa_ipc_close": (c_int, [c_void_p]),
    "fa_ring_forward_pull": (c_int, [c_void_p] * 4 + [c_int] * 4 + [c_float, c_int, c_int, ctypes.POINTER(c_void_p),
                                     ctypes.POINTER(c_void_p), c_void_p, ctypes.c_size_t, c_void_p]),
    "gemm_fused_epilogue": (c_int, [c_void_p] * 5 + [c_int] * 3 + [c_void_p]),
    "gemm_fused_epilogue_act": (c_int, [c_void_p] * 5 + [c_int] * 4 + [c_void_p]),
    "gemm_fused_epilogue_f32": (c_int, [c_void_p] * 5 + [c_int] * 4 + [c_void_p]),
    "gemm_batched_strided": (c_int, [c_void_p] * 3 + [c_int] * 4 + [ctypes.c_longlong] * 6 + [c_int, c_float, c_void_p]),
    "gemm_batched": (c_int, [c_void_p] * 3 + [c_int] * 5 + [c_float, c_void_p]),
    "softmax_rows": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "softmax_rows_f16out": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "fa_forward_unfused": (c_int, [c_void_p] * 6 + [c_int] * 4 + [c_float, c_void_p]),
    "fa_forward_host": (c_int, [c_void_p] * 4 + [c_int] * 4 + [c_float]),
    "gemm_fused_epilogue_host": (c_int, [c_void_p] * 5 + [c_int] * 3),
    "softmax_rows_host": (c_int, [c_void_p, c_void_p, c_int, c_int]),
    "fa_host_release": (c_int, []),
    "fa_last_debug": (c_int, [ctypes.POINTER(ctypes.c_uint), c_int]),
    "fa_launch_count": (ctypes.c_ulonglong, []),
    "fa_plan_segments": (c_int, [c_int] * 5 + [ctypes.POINTER(c_int), c_int]),
    "fa_set_tail_split": (c_int, [c_int]),
    "fa_map_cache_stats": (c_int, [ctypes.POINTER(ctypes.c_ulonglong)] * 2),
}


class FaError(RuntimeError):
    """A non-zero return code from the C ABI (negative: FA_ERR_*, positive: cudaError_t)."""

    def __init__(self, code, what, message):
        super().__init__(f"{what} failed with code {code}: {message}")
        self.code = code


def lib_path():
    return os.path.join(_HERE, _LIB_NAME)


def load_library():
    """Load the CUDA library.  Raises (never falls back) when it has not been built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = lib_path()
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  This package has no CPU fallback.")
        lib = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
        return lib


def check(code, what):
    if code != 0:
        msg = load_library().fa_strerror(code)
        raise FaError(code, what, msg.decode() if msg else "?")
