"""Small pass over every device entry point, meant to be run under compute-sanitizer (tools/gpu/sanitize.sh):
two-stage solves with cooperation and the speculative line search active (batch far below the grid), a ragged batch,
corridors with more than 32 planes, the line-initialised single stage, time allocation, Bezier sampling, model (B)."""
import sys

import numpy as np

sys.path.insert(0, ".")
from direct_b200 import make_batch  # noqa: E402
from direct_b200.capi import Solver  # noqa: E402

knots = int(sys.argv[1]) if len(sys.argv) > 1 else 40
s = Solver(0, "fp64")
pb = make_batch(48, knots, "box")
g0, g1 = s.solve_two_stage(pb, want_stage0=True)
print("box two-stage:", np.bincount(g1.rtn + 4, minlength=7).tolist(), "iters max", int(g1.iters.max()), "stats", s.stats().coop_jobs, s.stats().helper_units)
pb = make_batch(24, 20, "poly40")
g0, g1 = s.solve_two_stage(pb)
print("poly40 two-stage:", np.bincount(g1.rtn + 4, minlength=7).tolist())
pb = make_batch(24, 30, "poly")
g0, g1 = s.solve_two_stage(pb)
print("poly two-stage:", np.bincount(g1.rtn + 4, minlength=7).tolist())
s.close()
print("done")
