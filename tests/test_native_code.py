"""CPU check that the shipped library really is Blackwell-native: the SASS of libfa_b200.so must contain the
tcgen05 / TMA / TMEM instructions the design claims (mnemonics per /opt/skills/guides/B200_PROFILING.md) and no
legacy tensor path (HMMA = mma.sync / wmma)."""
import os
import shutil
import subprocess

import pytest

import flashattention_kernel_project_b200 as fa


@pytest.fixture(scope="module")
def sass():
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    if not os.path.exists(fa.lib_path()):
        import __graft_entry__ as g
        g.build_cuda()
    out = subprocess.run([cuobjdump, "-sass", fa.lib_path()], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
            funcs[cur] = []
        elif cur is not None:
            funcs[cur].append(line)
    return {k: "\n".join(v) for k, v in funcs.items()}


def _one(sass, needle):
    hits = [k for k in sass if needle in k]
    assert hits, f"no kernel matching {needle}"
    return [sass[k] for k in hits]


def test_attention_kernel_uses_tcgen05_tmem_and_tma(sass):
    for body in _one(sass, "fa_fwd_kernel"):
        assert "UTCHMMA" in body          # tcgen05.mma
        assert "LDTM" in body and "STTM" in body  # tcgen05.ld / tcgen05.st (S out of TMEM, P back into TMEM)
        assert "UTMALDG" in body and "UTMASTG" in body  # TMA tile loads and the TMA store of O
        assert "SYNCS" in body            # mbarrier phase waits
        assert "USETMAXREG" in body       # setmaxnreg register redistribution between roles
        assert "MUFU.EX2" in body         # exp2-domain softmax
        assert "HMMA" not in body.replace("UTCHMMA", "")  # no mma.sync / wmma path
        assert "BRA.U.ANY" not in body    # no per-thread serialisation loops around uniform-datapath instructions


def test_gemm_kernel_uses_cta_pairs(sass):
    for body in _one(sass, "gemm_fused_kernel"):
        assert "UTCHMMA.2CTA" in body     # tcgen05.mma.cta_group::2
        assert "UTMALDG.3D.2CTA" in body  # 2-SM TMA loads (batch is the third coordinate) signalling the leader's barrier
        assert "UTMASTG.3D" in body       # the epilogue leaves through TMA stores
        assert "UTCBAR.2CTA.MULTICAST" in body  # multicast tcgen05.commit
        assert "LDTM" in body
        assert "HMMA" not in body.replace("UTCHMMA", "")


def test_softmax_kernels_are_plain_bandwidth_kernels(sass):
    for body in _one(sass, "softmax_rows_reg_kernel"):
        assert "LDG.E.128" in body or "LDG.E.EF.128" in body or ".128" in body  # 128-bit global accesses
        assert "UTCHMMA" not in body and "HMMA" not in body


def test_kernels_keep_their_state_in_registers():
    """cuobjdump -res-usage: no kernel uses local memory arrays; the GEMM, the softmax and the D = 16 kernels have no
    stack frame at all; the attention kernel keeps a few 4-byte spill slots in the correction warps' per-SEGMENT code
    (80 registers: segment bookkeeping and the stream-K merge, touched once per item, never in a loop over keys) --
    test_attention_hot_roles_do_not_spill checks that the softmax, MMA and TMA roles spill nothing."""
    import re
    import subprocess
    out = subprocess.run(["cuobjdump", "-res-usage", fa.lib_path()], capture_output=True, text=True, check=True).stdout
    entries = re.findall(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", out)
    assert len(entries) >= 20
    for name, reg, stack, shared, local in entries:
        assert int(local) == 0, name
        limit = 32 if "fa_fwd_kernel" in name else 0
        assert int(stack) <= limit, f"{name}: stack frame of {stack} bytes"


def test_attention_hot_roles_do_not_spill(sass):
    """The roles are laid out in program order softmax (after the setmaxnreg.inc), correction (first dealloc), MMA,
    prefetch, TMA.  The softmax role touches no local memory at all; nothing after the correction role ever STORES
    to local memory (so the MMA / TMA / prefetch roles spill nothing), and the only local loads found there are cold
    blocks of the correction role that ptxas placed at the end of the function (they reload slots stored by it)."""
    import re
    for body in _one(sass, "fa_fwd_kernel"):
        lines = body.splitlines()
        marks = [i for i, l in enumerate(lines) if "USETMAXREG" in l]
        assert len(marks) >= 4 and "TRY_ALLOC" in lines[marks[0]]
        is_st = lambda l: re.search(r"\bSTL(\.|\s)", l) is not None
        is_ld = lambda l: re.search(r"\bLDL(\.|\s)", l) is not None
        assert not any(is_st(l) or is_ld(l) for l in lines[marks[0]:marks[1]]), "the softmax role touches local memory"
        assert not any(is_st(l) for l in lines[marks[2]:]), "a role after the correction warps spills"
        slot = lambda l: re.search(r"\[R1(\+0x[0-9a-f]+)?\]", l).group(0)
        stored = {slot(l) for l in lines[marks[1]:marks[2]] if is_st(l)}
        late = {slot(l) for l in lines[marks[2]:] if is_ld(l)}
        assert late <= stored, (late, stored)
