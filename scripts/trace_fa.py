"""Developer tool: per-iteration role timeline of one CTA of fa_fwd_kernel (built with -DFA_TRACE
into libfa_b200_trace.so).  Usage on the GPU box: python scripts/trace_fa.py [cta] > gpurun_out/trace.txt"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libname = sys.argv[2] if len(sys.argv) > 2 else "libfa_b200_trace.so"
lib = ctypes.CDLL(os.path.join(ROOT, "flashattention_kernel_project_b200", libname))
lib.fa_forward.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 4 + [ctypes.c_float, ctypes.c_void_p]
lib.fa_trace_config.argtypes = [ctypes.c_void_p, ctypes.c_int]
B, H, N, D = 8, 32, int(os.environ.get("FA_TRACE_N", "4096")), int(os.environ.get("FA_TRACE_D", "128"))
cta = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
q, k, v = ((torch.rand((B, H, N, D), device="cuda") * 2 - 1).half() for _ in range(3))
o = torch.empty_like(q)
trace = torch.zeros((6, 64, 8), dtype=torch.int64, device="cuda")
brief = len(sys.argv) > 3
run = lambda: lib.fa_forward(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, H, N, D, D ** -0.5, None)
for _ in range(3):
    assert run() == 0
torch.cuda.synchronize()
lib.fa_trace_config(trace.data_ptr(), cta)
assert run() == 0
torch.cuda.synchronize()
t = trace.cpu()
t0 = int(t[t > 0].min())
names = ["soft0", "soft1", "corr0", "corr1", "mma0", "mma1"]
pts = {"soft": ["wait_s", "s_ready", "ld_done", "stats_done", "chunk0_pub", "p_arrived", "max_done"],
       "corr": ["stats_seen", "corr_arrived"], "mma": ["wait_p", "p_seen", "corr_seen", "pv_issued", "qk_issued"]}
for j in range(0, 0 if brief else 32):
    print(f"--- iteration {j}")
    for r, name in enumerate(names):
        labels = pts[name[:4] if name[:4] in pts else name[:3]]
        vals = [int(t[r, j, k]) - t0 if int(t[r, j, k]) > 0 else None for k in range(len(labels))]
        print(f"  {name}: " + "  ".join(f"{l}={v}" for l, v in zip(labels, vals)))
# summary: steady-state period and phase durations over iterations 8..28
import statistics as st
def col(r, k):
    return [int(t[r, j, k]) for j in range(8, 29)]
for i in (0, 1):
    s_ready, ld, sd, c0, pa, mx = (col(i, k) for k in (1, 2, 3, 4, 5, 6))
    per = [b - a for a, b in zip(s_ready, s_ready[1:])]
    d = lambda x, y: st.mean(b - a for a, b in zip(x, y))
    if os.environ.get("FA_TRACE_VERDICT"):
        vw = col(i, 7)
        print(f"soft{i}: held_compute={d(ld, mx):.0f} verdict_wait={d(mx, vw):.0f} after_verdict={d(vw, sd):.0f}")
    print(f"soft{i}: period={st.mean(per):.0f}  ld={d(s_ready, ld):.0f} max={d(ld, mx):.0f} stats={d(mx, sd):.0f} "
          f"chunk0={d(sd, c0):.0f} chunks1-3={d(c0, pa):.0f} wait_for_S={d(pa, s_ready[1:]):.0f}")
    p_seen, pv, qk = (col(4 + i, k) for k in (1, 3, 4))
    print(f"mma{i}: chunk0_pub->seen={d(c0, p_seen):.0f} last_pub->pv_all_issued={d(pa, pv):.0f} "
          f"pv_issued->qk_issued={d(pv, qk):.0f} qk_issued->s_ready(next)={d(qk, s_ready[1:]):.0f}")

top = [int(t[4, j, 5]) for j in range(8, 29)]
aft = [int(t[4, j, 6]) for j in range(8, 29)]
if all(x > 0 for x in top):
    print(f"mma warp: iteration period={st.mean(b - a for a, b in zip(top, top[1:])):.0f} kv_full waits={st.mean(b - a for a, b in zip(top, aft)):.0f}")
