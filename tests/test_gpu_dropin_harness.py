"""In-harness drop-in proof (BASELINE north-star: "in the same .cu harness"): the reference's OWN programs, main()
included, compiled unmodified with their `<<<>>>` launches redirected to the C ABI (oracle/harness/*.cu interpose
cudaLaunchKernel; built by `make -C oracle ref` where the reference tree is mounted, the executables travel in
oracle/_ref/).  The assertions read what the harness itself prints."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(name):
    exe = os.path.join(ROOT, "oracle", "_ref", name)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (needs the reference tree at build time)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def test_v12i_program_with_its_launch_redirected(fa):
    """flashattn_streaming_16x32_2warp_softmax_v12i.cu: mt19937(123) fixture, 10 + 200 launches, cpu_reference, and
    the program's own "Relative L2 error" line (:381) -- computed on fa_forward_cross's output."""
    out = _run("harness_v12i")
    rel = float(re.search(r"Relative L2 error:\s*([0-9.eE+-]+)", out).group(1))
    red, fwd, failed = (int(x) for x in re.search(r"drop-in: (\d+) launches redirected to fa_forward_cross, (\d+) forwarded, (\d+) failed", out).groups())
    assert red == 210 and failed == 0
    assert rel <= 1e-3, out
    assert "NUM_BATCH=1024, SEQ_LEN=128, M_TOTAL=16, KDIM=16, DV=16" in out


def test_gemm_fused_program_with_its_launch_redirected(fa):
    """tensorcore_gemm_fused_epilogue.cu (M = N = K = 1024): both launches go to gemm_fused_epilogue; the check the
    program's dead `#if DO_VERIFY` block would print is made by the wrapper with the reference's cpu_gemm_fused_ref
    and the reference's kernel on the program's own buffers."""
    out = _run("harness_gemm_fused")
    g = lambda pat: float(re.search(pat + r":\s*([0-9.eE+-]+)", out).group(1))
    assert "drop-in: 2 launches redirected to gemm_fused_epilogue, 0 failed" in out
    assert g("Max abs diff vs CPU fused float ref") <= 2e-3
    assert g("Relative L2 error vs CPU fused float ref") <= 1e-3
    assert g("Max abs diff vs wmma_gemm_fused_kernel") <= 2e-3
    assert "Kernel time:" in out  # the harness timed the redirected launch
