"""First GPU contact: parity vs the oracle on a few shapes + a quick throughput probe."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from direct_b200 import make_batch
from direct_b200.capi import Solver
from oracle import oracle_py as O

def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))

prec = sys.argv[1] if len(sys.argv) > 1 else "fp64"
s = Solver(0, prec)
for kind in ("box", "poly"):
    for N in (4, 30, 100):
        pb = make_batch(16, N, kind, first=100)
        a0, a1 = O.two_stage_batch(pb, nthreads=O.max_threads())
        g0, g1 = s.solve_two_stage(pb)
        print(kind, N, "rtn0 eq", (a0.rtn == g0.rtn).all(), "it0 eq", (a0.iters == g0.iters).all(),
              "rtn1 eq", (a1.rtn == g1.rtn).all(), "it1 eq", (a1.iters == g1.iters).all())
        for nm in ("cost", "poly_coeff", "bez_coeff", "poly_time", "jerk", "x_final"):
            print("   %-10s s0 %.2e  s1 %.2e" % (nm, rel(getattr(g0, nm), getattr(a0, nm)), rel(getattr(g1, nm), getattr(a1, nm))))
        print("   stats eq", (a0.stats == g0.stats).all(), (a1.stats == g1.stats).all())
for B in (4096,):
    pb = make_batch(B, 100, "box")
    for rep in range(3):
        t = time.time(); g0, g1 = s.solve_two_stage(pb, want_stage0=False); dt = time.time() - t
        st = s.stats()
        print(prec, "B", B, "wall %.3fs kernel %.2f ms -> %.0f solves/s (kernel), h2d %.2f ms d2h %.2f ms, grid %d x %d smem %d slots %d"
              % (dt, st.kernel_ms, B / st.kernel_ms * 1e3, st.h2d_ms, st.d2h_ms, st.grid_blocks, st.block_threads, st.smem_bytes_per_block, st.workspace_slots))
        print("   bwd_knots", st.bwd_knots, "fwd_knots", st.fwd_knots, "rtn hist", np.unique(g1.rtn, return_counts=True))
