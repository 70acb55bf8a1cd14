"""What would perfect hardness prediction buy?  Solves the bench batch, then the SAME batch permuted so that the solves that
took longest come first (longest-processing-time-first on the kernel's work queue), and prints both kernel times.
    python tools/lpt_probe.py [--batch 4096] [--knots 100]"""
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
a = ap.parse_args()


def permuted(pb, o):
    return ProblemBatch(len(o), pb.N, pb.P_max, np.ascontiguousarray(pb.planes[o]), np.ascontiguousarray(pb.nplanes[o]),
                        np.ascontiguousarray(pb.durations[o]), np.ascontiguousarray(pb.seeds[o]),
                        np.ascontiguousarray(pb.x0[o]), np.ascontiguousarray(pb.xd[o]), pb.max_vel, pb.max_acc)


def best_of(s, pb, n=3):
    best = None
    for _ in range(n):
        g0, g1 = s.solve_two_stage(pb, want_stage0=True)
        st = s.stats()
        if best is None or st.kernel_ms < best[0]:
            best = (st.kernel_ms, (g0.stats[:, 6] + g1.stats[:, 6]).astype(np.float64), g1.iters.copy(), g0.iters.copy())
    return best


pb = make_batch(a.batch, a.knots, "box")
s = Solver(0, "fp64")
ms, cyc, it1, it0 = best_of(s, pb)
print(f"queue in index order : kernel {ms:.1f} ms")
print("stage-1 iterations: p50/p90/p99/max", np.percentile(it1, [50, 90, 99, 100]), " stage-0:", np.percentile(it0, [50, 90, 99, 100]))
for k in (10, 15, 20, 30, 40, 60):
    print(f"   solves with more than {k} stage-1 iterations: {(it1 > k).sum()}")
o = np.argsort(-cyc, kind="stable")
ms2, cyc2, _, _ = best_of(s, permuted(pb, o))
print(f"longest first (oracle): kernel {ms2:.1f} ms; longest solve {cyc2.max() / 1e6:.1f} Mcycles")
o = np.argsort(-it1.astype(np.int64), kind="stable")
ms3, cyc3, _, _ = best_of(s, permuted(pb, o))
print(f"most iterations first : kernel {ms3:.1f} ms; longest solve {cyc3.max() / 1e6:.1f} Mcycles")
hard = np.flatnonzero(it1 > 30)
rest = np.flatnonzero(it1 <= 30)
ms4, cyc4, _, _ = best_of(s, permuted(pb, np.concatenate([hard, rest])))
print(f"the {len(hard)} solves with > 30 iterations first, rest in index order: kernel {ms4:.1f} ms; longest solve {cyc4.max() / 1e6:.1f} Mcycles")
s.close()
