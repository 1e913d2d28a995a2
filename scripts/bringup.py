"""GPU bring-up checks (developer tool, not part of the pytest suite): each case runs in its own
process so a trapped kernel cannot poison the others.  Usage: python scripts/bringup.py [case ...]"""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


def _err(a, b):
    a, b = a.double(), b.double()
    rel = ((a - b).pow(2).sum() / (b.pow(2).sum() + 1e-30)).sqrt().item()
    return rel, (a - b).abs().max().item()


def _attn_ref(q, k, v, scale):
    s = torch.matmul(q.float(), k.float().transpose(-1, -2)) * scale
    p = torch.softmax(s, dim=-1)
    return torch.matmul(p, v.float())


def _attn_case(B, H, N, D, dist="uniform", f32=False, seed=0):
    import flashattention_kernel_project_b200 as fa
    g = torch.Generator(device="cuda").manual_seed(seed)
    shape = (B, H, N, D)
    if dist == "uniform":
        mk = lambda: (torch.rand(shape, device="cuda", generator=g) * 2 - 1).half()
    else:
        mk = lambda: torch.randn(shape, device="cuda", generator=g).half()
    q, k, v = mk(), mk(), mk()
    scale = 1.0 / D ** 0.5
    out = (fa.fa_forward_f32out if f32 else fa.fa_forward)(q, k, v, scale)
    torch.cuda.synchronize()
    ref = _attn_ref(q, k, v, scale)
    rel, mx = _err(out.float(), ref)
    print(f"attn B={B} H={H} N={N} D={D} {dist} f32={f32}: rel_l2={rel:.3e} max_abs={mx:.3e}", flush=True)
    if not (rel < 1e-3 and mx < 2e-3):
        bad = (out.float() - ref).abs()
        idx = bad.flatten().argmax().item()
        print("  worst flat index", idx, "out", out.flatten()[idx].item(), "ref", ref.flatten()[idx].item())
        print("  per-row-block max err:", bad.amax(dim=-1).flatten()[:16].tolist())
        print("  out[0,0,0,:8]", out[0, 0, 0, :8].tolist(), "\n  ref[0,0,0,:8]", ref[0, 0, 0, :8].tolist())
        return False
    return True


@case
def softmax():
    import flashattention_kernel_project_b200 as fa
    ok = True
    for rows, n in [(64, 4096), (128, 512), (256, 32), (33, 1000), (7, 4097), (5, 8192), (3, 20000), (16384, 4096)]:
        x = torch.randn(rows, n, device="cuda") * 3
        p = fa.softmax_rows(x)
        torch.cuda.synchronize()
        ref = torch.softmax(x.double(), dim=-1)
        rel, mx = _err(p, ref)
        print(f"softmax {rows}x{n}: rel_l2={rel:.3e} max_abs={mx:.3e}", flush=True)
        ok &= rel < 1e-5
    return ok


@case
def gemm_small():
    return _gemm(256, 512, 256)


@case
def gemm_mid():
    return _gemm(1024, 1024, 1024) and _gemm(1000, 1096, 520) and _gemm(128, 256, 64, bias=False, res=False)


def _gemm(M, N, K, bias=True, res=True):
    import flashattention_kernel_project_b200 as fa
    g = torch.Generator(device="cuda").manual_seed(1)
    a = (torch.rand(M, K, device="cuda", generator=g) * 2 - 1).half()
    b = (torch.rand(K, N, device="cuda", generator=g) * 2 - 1).half()
    bi = (torch.rand(N, device="cuda", generator=g) * 2 - 1) if bias else None
    r = (torch.rand(M, N, device="cuda", generator=g) * 2 - 1) if res else None
    c = fa.gemm_fused_epilogue(a, b, bi, r)
    torch.cuda.synchronize()
    ref = a.double() @ b.double()
    if bias:
        ref = ref + bi.double()
    ref = torch.relu(ref)
    if res:
        ref = ref + r.double()
    rel, mx = _err(c, ref)
    print(f"gemm M={M} N={N} K={K} bias={bias} res={res}: rel_l2={rel:.3e} max_abs={mx:.3e}", flush=True)
    if not (rel < 1e-3 and mx < 2e-3):
        bad = (c.double() - ref).abs()
        print("  row-block err (128 rows):", bad.amax(dim=1).view(-1, min(128, M)).amax(dim=1)[:8].tolist())
        print("  col-block err (64 cols):", bad.amax(dim=0)[: (N // 64) * 64].view(-1, 64).amax(dim=1)[:8].tolist())
        print("  c[0,:8]", c[0, :8].tolist(), "\n  ref[0,:8]", ref[0, :8].tolist())
        return False
    return True


@case
def attn_cfg1():
    return _attn_case(1, 1, 128, 64)


@case
def attn_d64():
    return _attn_case(1, 2, 256, 64) and _attn_case(2, 3, 1024, 64) and _attn_case(1, 1, 100, 64) and \
        _attn_case(1, 2, 1000, 64, "normal")


@case
def attn_d128_small():
    return _attn_case(1, 1, 128, 128) and _attn_case(1, 1, 256, 128)


@case
def attn_d128():
    return _attn_case(2, 2, 512, 128) and _attn_case(1, 3, 1000, 128, "normal") and \
        _attn_case(1, 2, 4096, 128) and _attn_case(1, 2, 300, 128, f32=True)


@case
def attn_bench():
    import flashattention_kernel_project_b200 as fa
    ok = True
    for (B, H, N, D) in [(8, 32, 4096, 128), (4, 16, 1024, 64)]:
        q = (torch.rand(B, H, N, D, device="cuda") * 2 - 1).half()
        k = (torch.rand(B, H, N, D, device="cuda") * 2 - 1).half()
        v = (torch.rand(B, H, N, D, device="cuda") * 2 - 1).half()
        o = torch.empty_like(q)
        scale = 1.0 / D ** 0.5
        for _ in range(3):
            fa.fa_forward(q, k, v, scale, out=o)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        e0.record()
        for _ in range(iters):
            fa.fa_forward(q, k, v, scale, out=o)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        fl = 4.0 * B * H * N * N * D
        print(f"attn bench B={B} H={H} N={N} D={D}: {ms:.4f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
        ref = _attn_ref(q[:1, :2], k[:1, :2], v[:1, :2], scale)
        rel, mx = _err(o[:1, :2].float(), ref)
        print(f"   check rel_l2={rel:.3e} max_abs={mx:.3e}")
        ok &= rel < 1e-3
    return ok


@case
def gemm_bench():
    import flashattention_kernel_project_b200 as fa
    M = N = K = 8192
    a = (torch.rand(M, K, device="cuda") * 2 - 1).half()
    b = (torch.rand(K, N, device="cuda") * 2 - 1).half()
    bi = torch.rand(N, device="cuda")
    r = torch.rand(M, N, device="cuda")
    c = torch.empty(M, N, device="cuda")
    for _ in range(3):
        fa.gemm_fused_epilogue(a, b, bi, r, out=c)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fa.gemm_fused_epilogue(a, b, bi, r, out=c)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"gemm bench 8192^3: {ms:.4f} ms {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
    ref = torch.relu(a[:256].float() @ b.float() + bi) + r[:256]
    rel, mx = _err(c[:256], ref)
    print(f"   check rel_l2={rel:.3e} max_abs={mx:.3e}")
    return rel < 1e-3


@case
def softmax_bench():
    import flashattention_kernel_project_b200 as fa
    x = torch.randn(16384, 4096, device="cuda")
    p = torch.empty_like(x)
    for _ in range(3):
        fa.softmax_rows(x, out=p)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fa.softmax_rows(x, out=p)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"softmax bench 16384x4096: {ms * 1e3:.1f} us {2 * x.numel() * 4 / ms / 1e6:.1f} GB/s", flush=True)
    return True


def main():
    names = sys.argv[1:] or list(CASES)
    if len(names) == 1 and names[0] in CASES and "--child" in sys.argv:
        pass
    if "--child" in sys.argv:
        name = [n for n in sys.argv[1:] if n != "--child"][0]
        import flashattention_kernel_project_b200 as fa
        try:
            ok = CASES[name]()
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(f"EXCEPTION in {name}: {type(e).__name__}: {e}")
            try:
                print("  debug record:", [hex(w) for w in fa.last_debug()])
            except Exception:  # noqa: BLE001
                pass
            ok = False
        print(f"CASE {name}: {'PASS' if ok else 'FAIL'}", flush=True)
        sys.exit(0 if ok else 1)
    results = {}
    for name in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, name, "--child"], timeout=180)
            results[name] = "PASS" if r.returncode == 0 else f"FAIL(rc={r.returncode})"
        except subprocess.TimeoutExpired:
            results[name] = "TIMEOUT"
        print(f"== {name}: {results[name]} ({time.time() - t0:.1f}s)", flush=True)
    print("SUMMARY", results)


if __name__ == "__main__":
    main()
