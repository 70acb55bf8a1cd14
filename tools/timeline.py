"""Per-iteration timeline of ONE hard trajectory inside a full batch (it is made trajectory 0 of the batch, the one the
device trace records): python tools/timeline.py [--first 547] [--batch 4096]"""
import argparse, sys
import numpy as np
sys.path.insert(0, ".")
from direct_b200 import make_batch  # noqa: E402
from direct_b200.capi import Solver  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--first", type=int, default=547)
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--knots", type=int, default=100)
ap.add_argument("--cycles", action="store_true", help="library built with -DDDP_TRACE_CYCLES: costq/logcost/err/opterr carry kilo-cycles per phase")
a = ap.parse_args()
pb = make_batch(a.batch, a.knots, "box", first=a.first)
s = Solver(0, "fp64", trace=1)
for _ in range(2):
    g0, g1 = s.solve_two_stage(pb, want_stage0=True)
    st = s.stats()
rows = s.trace()
print(f"kernel {st.kernel_ms:.1f} ms; trajectory 0: stage-1 iterations {g1.iters[0]}, total cycles {(g0.stats[0, 6] + g1.stats[0, 6]) / 1e6:.1f} M; trace rows {len(rows)}")
prev = 0
print("iter  t_ms  dt_us step failed n_bwd" + ("  | kcycles: backward (riccati)  forward (sequential part / wait)" if a.cycles else ""))
for k, r in enumerate(rows):
    extra = (f"  | {r['costq']:6.0f} ({r['err']:5.0f}) {r['logcost']:6.0f} ({r['opterr']:5.0f})"
             f"  | recursion: assembly {r['cost']:4.0f} rounds {r['mu']:4.0f} gains {r['reg']:4.0f} backup {r['stepsize']:4.0f}") if a.cycles else ""
    print(f"{k:4d} {r['t_us'] / 1e3:6.1f} {r['t_us'] - prev:6d} {r['step']:3d} {r['fp_failed']:3d} {r['n_bwd']:3d}" + extra)
    prev = r['t_us']
s.close()
