"""CPU tests of the drop-in boundary: the shared library loads, exports exactly what
include/fa_b200.h declares, validates arguments, and refuses to compute without a B200 (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import flashattention_kernel_project_b200 as fa
from flashattention_kernel_project_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(fa.lib_path()):
        import __graft_entry__ as g
        g.build_cuda()
    return fa.load_library()


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "fa_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.findall(r"\b([a-z_0-9]+)\s*\(", re.sub(r"#.*", "", text))


def test_header_and_binding_agree(lib):
    declared = set(_declared_functions())
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/fa_b200.h but not exported"


def test_no_torch_types_in_c_abi():
    text = open(os.path.join(ROOT, "include", "fa_b200.h")).read()
    assert "torch" not in text and "at::" not in text and 'extern "C"' in text


def test_version_and_strerror(lib):
    assert lib.fa_version() >= 100
    assert b"64 or 128" in lib.fa_strerror(-2)
    assert b"no CPU fallback" in lib.fa_strerror(-5)


def test_argument_validation_happens_before_any_device_work(lib):
    p = ctypes.c_void_p(0x1000)
    assert lib.fa_forward(p, p, p, p, 1, 1, 128, 96, 0.1, None) == -2      # bad head dim
    assert lib.fa_forward(p, p, p, p, 0, 1, 128, 64, 0.1, None) == -1      # bad shape
    assert lib.fa_forward(None, p, p, p, 1, 1, 128, 64, 0.1, None) == -4   # null pointer
    assert lib.fa_forward(ctypes.c_void_p(0x1004), p, p, p, 1, 1, 128, 64, 0.1, None) == -3  # misaligned
    assert lib.fa_forward(p, p, p, p, 1, 1, 128, 64, float("nan"), None) == -7
    assert lib.gemm_fused_epilogue(p, p, None, None, p, 128, 100, 64, None) == -3  # N % 8 != 0
    assert lib.gemm_fused_epilogue(p, p, None, None, p, 128, 128, 0, None) == -1
    assert lib.softmax_rows(p, None, 4, 4, None) == -4
    assert lib.fa_forward_partial(p, p, p, p, None, p, 1, 1, 128, 128, 64, 0.1, 1, 0, None, None) == -4


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour on a box without a GPU")
def test_compute_calls_fail_loudly_without_a_gpu(lib):
    p = ctypes.c_void_p(0x1000)
    assert lib.fa_forward(p, p, p, p, 1, 1, 128, 64, 0.1, None) == -5
    assert lib.softmax_rows(p, p, 4, 4, None) == -5
    assert lib.gemm_fused_epilogue(p, p, None, None, p, 128, 128, 64, None) == -5
    x = np.zeros((4, 8), np.float32)
    with pytest.raises(fa.FaError) as e:
        fa.softmax_rows_host(x, np.empty_like(x))
    assert e.value.code == -5


def test_python_mirror_rejects_cpu_tensors():
    with pytest.raises(TypeError):
        fa.softmax_rows(torch.zeros(4, 8))
    with pytest.raises(TypeError):
        fa.fa_forward(torch.zeros(1, 1, 128, 64, dtype=torch.float16), torch.zeros(1, 1, 128, 64, dtype=torch.float16),
                      torch.zeros(1, 1, 128, 64, dtype=torch.float16))
    with pytest.raises(TypeError):
        fa.gemm_fused_epilogue(torch.zeros(16, 16, dtype=torch.float16), torch.zeros(16, 16, dtype=torch.float16))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "flashattention_kernel_project_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "__init__.py" and "oracle" not in text, (dirpath, f)


def build_and_run_c_example(tmp_path):
    """gcc-compiles examples/drop_in.c against include/fa_b200.h + libfa_b200.so and runs it."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cc = shutil.which("gcc")
    if cc is None:
        pytest.skip("gcc not available")
    subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c",
                    os.path.join(root, "include", "fa_b200.h")], check=True)
    libdir = os.path.join(root, "flashattention_kernel_project_b200")
    exe = str(tmp_path / "drop_in")
    subprocess.run([cc, "-std=c99", "-Wall", "-I" + os.path.join(root, "include"), os.path.join(root, "examples", "drop_in.c"),
                    "-o", exe, "-L" + libdir, "-lfa_b200", "-lm", "-Wl,-rpath," + libdir], check=True)
    return subprocess.run([exe], capture_output=True, text=True, timeout=120)


def test_header_is_plain_c_and_a_c_program_links_against_the_library(tmp_path):
    """The boundary is C, not C++: the header compiles as C99 with -pedantic, and examples/drop_in.c (gcc, no CUDA
    headers) links against libfa_b200.so.  Without an sm_100 GPU the program must report FA_ERR_NO_DEVICE (exit 2);
    tests/test_gpu_ring_host.py runs the same program on the GPU box, where it checks an output row itself."""
    r = build_and_run_c_example(tmp_path)
    assert "fa_version" in r.stdout
    if not torch.cuda.is_available():
        assert r.returncode == 2 and "-5" in r.stdout, r.stdout  # FA_ERR_NO_DEVICE, reported through fa_strerror


def test_host_wrappers_validate_every_buffer_before_any_copy():
    """The *_host entry points size their copies from the FIRST argument: a short, mistyped or non-contiguous K, V,
    bias, residual or output would be read or written past its end inside cudaMemcpyAsync.  The mirror rejects
    them on the host (numpy arrays and torch CPU tensors alike) before the library is called."""
    import numpy as np
    import pytest
    import torch
    import flashattention_kernel_project_b200 as fa
    a, b = np.zeros((256, 64), np.float16), np.zeros((64, 128), np.float16)
    good = np.zeros((256, 128), np.float32)
    for bad in (np.zeros((256, 128), np.float16), np.zeros((255, 128), np.float32),
                torch.zeros((256, 128), dtype=torch.float16), np.zeros((128, 256), np.float32).T):
        with pytest.raises((TypeError, ValueError)):
            fa.gemm_fused_epilogue_host(a, b, None, None, bad)
    with pytest.raises(ValueError):
        fa.gemm_fused_epilogue_host(a, b, np.zeros(127, np.float32), None, good)
    with pytest.raises(ValueError):
        fa.gemm_fused_epilogue_host(a, b, None, np.zeros((256, 64), np.float32), good)
    q = np.zeros((1, 2, 128, 64), np.float16)
    with pytest.raises(ValueError):
        fa.fa_forward_host(q, q, q[:, :, :64].copy(), q.copy())
    with pytest.raises(TypeError):
        fa.fa_forward_host(q, q, q, np.zeros(q.shape, np.float32))
    with pytest.raises(TypeError):
        fa.fa_forward_host(torch.zeros(q.shape, dtype=torch.float32), q, q, q.copy())
    x = np.zeros((8, 128), np.float32)
    with pytest.raises(ValueError):
        fa.softmax_rows_host(x, np.zeros((8, 64), np.float32))


def test_ring_entry_validates_without_a_device():
    """fa_ring_forward / fa_ring_workspace_bytes: argument checks happen before any device or NCCL work."""
    import flashattention_kernel_project_b200 as fa
    lib = fa.load_library()
    full = lib.fa_ring_workspace_bytes(1, 32, 8192, 128, 8, 0)
    state = lib.fa_ring_workspace_bytes(1, 32, 8192, 128, 1, 0)
    shard = 32 * 8192 * 128 * 2
    assert full == state + 7 * 2 * shard  # carried fp32 (O_acc, m, l) + world - 1 receive buffers of K and V
    assert lib.fa_ring_workspace_bytes(1, 32, 8192, 128, 8, 2) == state + 2 * 2 * shard
    assert lib.fa_ring_workspace_bytes(0, 32, 8192, 128, 8, 0) == 0
    p = ctypes.c_void_p(256)
    args = lambda **kw: [kw.get("q", p), p, p, p, 1, 2, 128, kw.get("D", 128), 0.1, kw.get("rank", 0), kw.get("world", 2),
                         kw.get("comm", p), kw.get("ws", p), kw.get("bytes", 1 << 30), None]
    assert lib.fa_ring_forward(*args(D=32)) == -2
    assert lib.fa_ring_forward(*args(rank=2)) == -7
    assert lib.fa_ring_forward(*args(q=None)) == -4
    assert lib.fa_ring_forward(*args(comm=None)) == -4
    assert lib.fa_ring_forward(*args(ws=ctypes.c_void_p(8))) == -3
    assert lib.fa_ring_forward(*args(bytes=1024)) == -7  # workspace too small for the state and one receive buffer
