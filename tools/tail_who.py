"""Which solves end a launch, when were they taken from the queue and how long did they run?  Needs a library built with
-DDDP_TRACE_CYCLES (stats columns 4 / 5 = globaltimer at the start / end of each stage):
    DIRECT_DDP_LIB=tools/_variants/lib_trc.so python tools/tail_who.py [--batch 4096]"""
import argparse
import sys

import numpy as np

sys.path.insert(0, ".")
from direct_b200 import make_batch  # noqa: E402
from direct_b200.capi import Solver  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--knots", type=int, default=100)
ap.add_argument("--top", type=int, default=12)
a = ap.parse_args()
pb = make_batch(a.batch, a.knots, "box")
s = Solver(0, "fp64")
for _ in range(2):
    g0, g1 = s.solve_two_stage(pb, want_stage0=True)
    ms = s.stats().kernel_ms
t0 = g0.stats[:, 4].astype(np.float64); t1 = g1.stats[:, 5].astype(np.float64)
z = t0.min()
start, end = (t0 - z) * 1e-6, (t1 - z) * 1e-6
print(f"kernel {ms:.1f} ms; queue empty (last solve taken) at {start.max():.1f} ms; last solve ends at {end.max():.1f} ms")
o = np.argsort(-end)[: a.top]
print("  idx  taken at  ended at  ran for  stage-1 iterations")
for i in o:
    print(f"{i:5d} {start[i]:8.1f} {end[i]:9.1f} {end[i] - start[i]:8.1f} {g1.iters[i]:6d}")
hard = g1.iters >= 60
print(f"{hard.sum()} solves with >= 60 iterations: taken at min/median/max {start[hard].min():.1f}/{np.median(start[hard]):.1f}/{start[hard].max():.1f} ms, "
      f"run time min/median/max {(end - start)[hard].min():.1f}/{np.median((end - start)[hard]):.1f}/{(end - start)[hard].max():.1f} ms")
for lo in (0, 20, 40, 60, 70):
    m = hard & (start >= lo) & (start < lo + 20 if lo < 70 else start >= lo)
    if m.any():
        print(f"   taken in [{lo}, {lo + 20 if lo < 70 else 'end'}) ms: {m.sum():3d} solves, mean run time {(end - start)[m].mean():6.1f} ms, mean end {end[m].mean():6.1f} ms")
s.close()
