"""Stress of the tail mechanisms (cooperation, speculative line search, speculative backward sweep): many small and odd batch
shapes, each solved with everything on and with DIRECT_DDP_COOP=0 (plain one-warp-per-trajectory), results compared bit for bit.
    python tools/stress_tail.py [--reps 2]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from direct_b200 import make_batch  # noqa: E402
from direct_b200.capi import Solver  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
FIELDS = ("rtn", "iters", "infeas_out", "cost", "poly_coeff", "bez_coeff", "poly_time", "x_final", "jerk")
s = Solver(0, "fp64")
bad = 0
cases = [(B, N, kind) for kind in ("box", "poly") for N in (5, 20, 100) for B in (1, 2, 3, 5, 17, 64, 149, 301, 1185)]
cases += [(7, 33, "poly40"), (130, 12, "poly40"), (2400, 40, "box")]
t0 = time.time()
for B, N, kind in cases:
    pb = make_batch(B, N, kind, first=B * 7 + N)
    os.environ["DIRECT_DDP_COOP"] = "0"
    r0, r1 = s.solve_two_stage(pb, want_stage0=True)
    os.environ.pop("DIRECT_DDP_COOP")
    used = [0, 0, 0]
    for _ in range(a.reps):
        g0, g1 = s.solve_two_stage(pb, want_stage0=True)
        st = s.stats()
        used = [used[0] + st.helper_units, used[1] + st.spec_trials, used[2] + st.spec_sweeps_used]
        for x, y in ((r0, g0), (r1, g1)):
            same = all(np.array_equal(getattr(x, f), getattr(y, f)) for f in FIELDS) and np.array_equal(x.stats[:, :4], y.stats[:, :4])
            if not same:
                bad += 1
                print("DIFFERENT", B, N, kind, flush=True)
    print(f"B {B:5d} N {N:3d} {kind:6s}: identical; helper units {used[0]}, remote trials {used[1]}, sweeps used {used[2]}; "
          f"iters max {int(g1.iters.max())}", flush=True)
print(f"{len(cases)} shapes x {a.reps} runs in {time.time() - t0:.1f} s: {'ALL IDENTICAL' if bad == 0 else str(bad) + ' DIFFERENT'}")
s.close()
sys.exit(1 if bad else 0)
