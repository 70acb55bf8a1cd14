"""Tail of the bench batch: per-solve cycle accounting of the longest solves of one launch, and the same hard
trajectories solved alone (every warp of the grid idle except theirs).
    python tools/tail_report.py [--batch 4096] [--knots 100] [--top 12]"""
import argparse
import sys

import numpy as np

sys.path.insert(0, ".")
from direct_b200 import make_batch  # noqa: E402
from direct_b200.capi import Solver  # noqa: E402
from direct_b200.problems import ProblemBatch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--knots", type=int, default=100)
ap.add_argument("--kind", default="box")
ap.add_argument("--top", type=int, default=12)
a = ap.parse_args()
MHZ = 1965.0
pb = make_batch(a.batch, a.knots, a.kind)
s = Solver(0, "fp64")
best = None
for _ in range(3):
    g0, g1 = s.solve_two_stage(pb, want_stage0=True)
    st = s.stats()
    if best is None or st.kernel_ms < best[0]:
        best = (st.kernel_ms, g0.stats.copy(), g1.stats.copy(), g1.iters.copy(), st.coop_jobs, st.helper_units)
ms, s0, s1, it1, jobs, units = best
tot = (s0[:, 6] + s1[:, 6]).astype(np.float64)
kcyc = ms * 1e-3 * MHZ * 1e6
print(f"kernel {ms:.1f} ms ({kcyc / 1e6:.0f} Mcycles), {a.batch / ms * 1e3:.0f} solves/s; jobs {jobs}, helper units {units}")
o = np.argsort(-tot)[: a.top]
print("idx  it1 | total Mcyc (of kernel) | stage1: sweeps trials fwdknots | bwd Mcyc (riccati) | fwd Mcyc (sequential)")
for i in o:
    ric = (s1[i, 7] & 0xffffffff) * 1024 / 1e6
    seq = (s1[i, 7] >> 32) * 1024 / 1e6
    print(f"{i:5d} {it1[i]:3d} | {tot[i] / 1e6:7.1f} ({tot[i] / kcyc:.2f}) | {s1[i, 0]:4d} {s1[i, 2]:4d} {s1[i, 3]:6d} | "
          f"{s1[i, 4] / 1e6:7.1f} ({ric:6.1f}) | {s1[i, 5] / 1e6:7.1f} ({seq:6.1f})")
hard = np.sort(o)
sub = ProblemBatch(len(hard), pb.N, pb.P_max, np.ascontiguousarray(pb.planes[hard]), np.ascontiguousarray(pb.nplanes[hard]),
                   np.ascontiguousarray(pb.durations[hard]), np.ascontiguousarray(pb.seeds[hard]),
                   np.ascontiguousarray(pb.x0[hard]), np.ascontiguousarray(pb.xd[hard]), pb.max_vel, pb.max_acc)
for _ in range(2):
    h0, h1 = s.solve_two_stage(sub, want_stage0=True)
    st = s.stats()
t2 = (h0.stats[:, 6] + h1.stats[:, 6]).astype(np.float64)
print(f"the same {len(hard)} trajectories alone: kernel {st.kernel_ms:.1f} ms, grid {st.grid_blocks}x{st.block_threads}; "
      f"per-solve Mcycles min/mean/max {t2.min() / 1e6:.1f}/{t2.mean() / 1e6:.1f}/{t2.max() / 1e6:.1f}")
i = int(np.argmax(t2))
ric = (h1.stats[i, 7] & 0xffffffff) * 1024 / 1e6
seq = (h1.stats[i, 7] >> 32) * 1024 / 1e6
print(f"   longest alone: bwd {h1.stats[i, 4] / 1e6:.1f} Mcyc (riccati {ric:.1f}), fwd {h1.stats[i, 5] / 1e6:.1f} Mcyc (sequential {seq:.1f}), "
      f"sweeps {h1.stats[i, 0]}, trials {h1.stats[i, 2]}, fwd knots {h1.stats[i, 3]}")
s.close()
