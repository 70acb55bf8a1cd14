"""BASELINE.json configs[4]: horizon x batch sweep (N in {50,100,200,400} x B in {1k,4k,16k,64k}), two-stage protocol,
device kernel time of the best of `reps` launches -> markdown table rows.
    python tools/sweep.py [--precision fp64] [--kind box] [--max-work 8e6]
"""
import argparse
import sys

sys.path.insert(0, ".")
from direct_b200 import make_batch  # noqa: E402
from direct_b200.capi import Solver  # noqa: E402
from bench import bwd_flops_per_knot  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="fp64")
ap.add_argument("--kind", default="box")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--max-work", type=float, default=3.0e7, help="skip points with B*N above this")
a = ap.parse_args()
s = Solver(0, a.precision)
peak = s.fma_peak_tflops(a.precision)
print(f"| N | B | kernel ms | solves/s | bwd TFLOP/s (algorithmic) | % of {a.precision} FMA peak ({peak:.1f} TF) | converged |")
print("|---|---|---|---|---|---|---|")
for N in (50, 100, 200, 400):
    for B in (1024, 4096, 16384, 65536):
        if B * N > a.max_work:
            print(f"| {N} | {B} | skipped (B*N > {a.max_work:.0e}) | | | | |")
            continue
        pb = make_batch(B, N, a.kind)
        best = None
        for _ in range(a.reps):
            _, g = s.solve_two_stage(pb, want_stage0=False)
            st = s.stats()
            if best is None or st.kernel_ms < best[0]:
                best = (st.kernel_ms, st.bwd_knots, float((g.rtn == 1).mean()))
        ms, bk, conv = best
        tf = bwd_flops_per_knot(float(pb.nplanes.mean())) * bk / (ms * 1e-3) / 1e12
        print(f"| {N} | {B} | {ms:.1f} | {B / ms * 1e3:.0f} | {tf:.2f} | {tf / peak * 100:.1f} | {conv:.3f} |", flush=True)
s.close()
