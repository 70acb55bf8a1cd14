"""Latency of ONE corridor (B = 1), the way the ROS node calls the path (teach_repeat_planner.cpp:853-951: two
polyCurveGeneration calls per corridor, single thread):
  dropin     the B200 drop-in translation unit driven through the reference's ddpTrajOptimizer class API
             (oracle/_ref/libddp_dropin.so: two B = 1 host-buffer C-ABI calls, each with H2D + kernel + D2H)
  fused      direct_ddp_solve_two_stage with B = 1 (both stages in one kernel, one H2D / D2H)
  port       the C restatement (oracle/ipddp_oracle.c), one host thread
  reference  the reference's own ddp_optimizer.cpp against the eager Eigen stand-in, one host thread (lower bound of its speed)
    python tools/latency_b1.py [--knots 10 30 50 100] [--reps 7] [--kind box]
Prints a markdown table (median wall time per corridor in ms) and one JSON line."""
import argparse
import json
import statistics
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from direct_b200 import make_batch  # noqa: E402
from direct_b200.capi import Solver  # noqa: E402
from direct_b200.problems import STAGE0, STAGE1  # noqa: E402
from oracle import oracle_py as O  # noqa: E402  (checker / CPU baselines only)

ap = argparse.ArgumentParser()
ap.add_argument("--knots", type=int, nargs="+", default=[10, 30, 50, 100])
ap.add_argument("--reps", type=int, default=7)
ap.add_argument("--kind", default="box")
ap.add_argument("--first", type=int, default=7)
a = ap.parse_args()
O.build(ref=False)


def two_calls(pb, **kw):
    r0 = O.solve_batch(pb, infeas=1, zero_init=1, **STAGE0, **kw)
    dur = np.where((r0.rtn == 2)[:, None], r0.poly_time, pb.durations)
    return O.solve_batch(pb, infeas=r0.infeas_out, zero_init=0, init_bez=r0.bez_coeff, durations=dur, **STAGE1, **kw)


def med(fn, reps):
    fn()   # warm-up
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(ts)


s = Solver(0, "fp64")
rows = []
for N in a.knots:
    pb = make_batch(1, N, a.kind, first=a.first)
    row = {"knots": N}
    row["fused_ms"] = med(lambda: s.solve_two_stage(pb, want_stage0=False), a.reps)
    st = s.stats()
    row["fused_kernel_ms"] = st.kernel_ms
    if O.dropin_available():
        row["dropin_ms"] = med(lambda: two_calls(pb, use_dropin=True), a.reps)
    row["port_1thread_ms"] = med(lambda: O.two_stage_batch(pb, nthreads=1), max(3, a.reps // 2))
    if O.ref_available():
        row["reference_tu_1thread_ms"] = med(lambda: two_calls(pb, use_ref=True), 3)
    g = s.solve_two_stage(pb, want_stage0=False)[1]
    row["stage1_iters"] = int(g.iters[0]); row["rtn"] = int(g.rtn[0])
    rows.append(row)
s.close()
cols = ["knots", "dropin_ms", "fused_ms", "fused_kernel_ms", "port_1thread_ms", "reference_tu_1thread_ms", "stage1_iters", "rtn"]
print("| " + " | ".join(cols) + " |")
print("|" + "---|" * len(cols))
for r in rows:
    print("| " + " | ".join(f"{r[c]:.2f}" if isinstance(r.get(c), float) else str(r.get(c, "-")) for c in cols) + " |")
print(json.dumps({"latency_b1": rows, "kind": a.kind}))
