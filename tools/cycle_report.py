"""Where the solve kernel's time goes: per-trajectory SM-cycle accounting of one launch of the bench workload.
    python tools/cycle_report.py [--precision fp64] [--batch 4096] [--knots 100] [--blocks-per-sm K] [--warps-per-block W]
Prints kernel time, the busy fraction of the warp slots (sum of per-solve cycles / (slots * kernel cycles)), the
longest single solve (the tail bound) and cycles per backward knot / rollout knot."""
import argparse
import sys

import numpy as np

sys.path.insert(0, ".")
from direct_b200 import make_batch  # noqa: E402
from direct_b200.capi import Solver  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--knots", type=int, default=100)
ap.add_argument("--kind", default="box")
ap.add_argument("--precision", default="fp64")
ap.add_argument("--blocks-per-sm", type=int, default=0)
ap.add_argument("--warps-per-block", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--tag", default="")
a = ap.parse_args()
pb = make_batch(a.batch, a.knots, a.kind)
s = Solver(0, a.precision, warps_per_block=a.warps_per_block, blocks_per_sm=a.blocks_per_sm)
best = None
for _ in range(a.reps):
    g0, g1 = s.solve_two_stage(pb, want_stage0=True)
    st = s.stats()
    if best is None or st.kernel_ms < best[0]:
        best = (st.kernel_ms, g0.stats.copy(), g1.stats.copy(), st)
ms, s0, s1, st = best
cyc = s0[:, 4:7] + s1[:, 4:7]
tot = cyc[:, 2].astype(np.float64)
mhz = 1965.0
kcyc = ms * 1e-3 * mhz * 1e6
slots = st.workspace_slots
bk = (s0[:, 1] + s1[:, 1]).sum()
fk = (s0[:, 3] + s1[:, 3]).sum()
print(f"[{a.tag}] {a.precision} B={a.batch} N={a.knots} {a.kind}: kernel {ms:.1f} ms = {a.batch / ms * 1e3:.0f} solves/s; grid {st.grid_blocks}x{st.block_threads}, "
      f"slots {slots}, smem/block {st.smem_bytes_per_block}")
print(f"   cooperation: {st.coop_jobs} jobs posted to idle warps, {st.helper_units} units run by helpers; "
      f"speculative backward sweeps {st.spec_sweeps} posted, {st.spec_sweeps_used} used")
print(f"   slot busy fraction {tot.sum() / (slots * kcyc):.3f}; longest solve {tot.max() / kcyc:.3f} of the kernel; mean solve {tot.mean() / kcyc:.4f}; "
      f"p50/p90/p99/max Mcycles {np.percentile(tot, 50) / 1e6:.2f}/{np.percentile(tot, 90) / 1e6:.2f}/{np.percentile(tot, 99) / 1e6:.2f}/{tot.max() / 1e6:.2f}")
print(f"   cycles per backward knot {cyc[:, 0].sum() / bk:.0f} ({bk} knots); per rollout knot {cyc[:, 1].sum() / fk:.0f} ({fk} knots); "
      f"share bwd {cyc[:, 0].sum() / tot.sum():.3f} fwd {cyc[:, 1].sum() / tot.sum():.3f}")
s7 = s0[:, 7] + s1[:, 7]   # packed kilo-cycles: Riccati | sequential rollout (both stages summed field-wise)
ric = ((s0[:, 7] & 0xffffffff) + (s1[:, 7] & 0xffffffff)).sum() * 1024.0
seq = ((s0[:, 7] >> 32) + (s1[:, 7] >> 32)).sum() * 1024.0
print(f"   Riccati {ric / bk:.0f} cycles/backward knot, linearisation + rest {(cyc[:, 0].sum() - ric) / bk:.0f}; sequential rollout {seq / fk:.0f} cycles/rollout knot, "
      f"rows + rest {(cyc[:, 1].sum() - seq) / fk:.0f}")
s.close()
