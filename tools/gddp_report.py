"""Generic DDP (model (B)): solves/s and backward-pass roofline on one B200.
    python tools/gddp_report.py [--batch 4096] [--knots 100] [--precision fp32] [--model quad]"""
import argparse, sys
import numpy as np
sys.path.insert(0, ".")
from direct_b200 import gddp  # noqa: E402
from direct_b200.capi import Solver  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--knots", type=int, default=100)
ap.add_argument("--precision", default="fp32")
ap.add_argument("--model", default="quad")
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
gp = gddp.make_quad_batch(a.batch, a.knots) if a.model == "quad" else gddp.make_dint_batch(a.batch, a.knots)
if a.precision == "fp32":
    import dataclasses
    gp = dataclasses.replace(gp, tol=1e-5)
s = Solver(0, a.precision)
peak = s.fma_peak_tflops(a.precision)
best = None
for _ in range(a.reps):
    g = gddp.solve(s, gp)
    st = s.stats()
    if best is None or st.kernel_ms < best[0]:
        best = (st.kernel_ms, g, st)
ms, g, st = best
knots = int(g.stats[:, 2].sum())
tf = gddp.bwd_flops_per_knot(gp.nx, gp.nu) * knots / (ms * 1e-3) / 1e12
print(f"[{a.model} {a.precision}] B={a.batch} N={a.knots}: kernel {ms:.3f} ms = {a.batch / ms * 1e3:.0f} solves/s; grid {st.grid_blocks}x{st.block_threads}, "
      f"smem/block {st.smem_bytes_per_block}; converged {(g.rtn == 1).mean():.4f}, iters {g.iters.mean():.2f}, sweeps {g.stats[:, 0].mean():.2f}, "
      f"rollouts {g.stats[:, 1].mean():.2f}")
print(f"   backward pass: {knots} knots x {gddp.bwd_flops_per_knot(gp.nx, gp.nu):.0f} flop = {tf:.2f} TFLOP/s algorithmic = {tf / peak * 100:.1f} % of the measured "
      f"{a.precision} FMA peak ({peak:.1f} TF); h2d {st.h2d_ms:.2f} ms, d2h {st.d2h_ms:.2f} ms")
s.close()
