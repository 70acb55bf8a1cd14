"""Generate tests/golden/*.npz from the REFERENCE's own translation unit.

Run in the dev container (where /root/reference is mounted):

    make -C oracle ref && PYTHONPATH=. python tests/golden/make_golden.py

Every fixture holds the inputs of one small batch and the outputs of ddpTrajOptimizer::polyCurveGeneration
as computed by /root/reference/global_planner/src/ddp_optimizer.cpp compiled unmodified against oracle/shim
(oracle/_ref/libddp_ref.so).  The reference ships no golden vectors of its own for this path (SURVEY.md
section 4), so these files are the pin: tests compare the C oracle, the emulated kernel and the CUDA path
against them.  Stage 1 is always started from the reference's own stage-0 output.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from direct_b200.problems import STAGE0, STAGE1, make_batch  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_FIELDS = ("rtn", "infeas_out", "line_failed_out", "iters", "cost", "poly_coeff", "bez_coeff", "poly_time")

CASES = {
    # name: (B, N, kind, first, overrides for both stages)
    "box_n5": (4, 5, "box", 0, {}),
    "poly_n12": (4, 12, "poly", 40, {}),
    "box_n50_single": (1, 50, "box", 7, {}),          # BASELINE.json configs[0] shape: one trajectory, 50 knots
    "poly_n30_minvo": (3, 30, "poly", 90, {"minvo": 1}),
    "box_n8_timepower1": (3, 8, "box", 300, {"time_power": 1}),
    "box_n100": (2, 100, "box", 1000, {}),
    "poly40_n10": (3, 10, "poly40", 700, {}),         # polytopes with up to 40 planes (m_c = 6 P + 55 up to 295 rows per knot)
    "poly_n200": (2, 200, "poly", 4200, {}),           # BASELINE.json configs[3] shape: 200 knots, polyhedral corridor (P <= 14)
}


def run_case(name, B, N, kind, first, ov):
    pb = make_batch(B, N, kind, first=first)
    s0 = dict(STAGE0, **ov)
    s1 = dict(STAGE1, **ov)
    r0 = O.solve_batch(pb, nthreads=4, use_ref=True, infeas=1, zero_init=1, **s0)
    dur1 = np.where((r0.rtn == 2)[:, None], r0.poly_time, pb.durations)
    r1 = O.solve_batch(pb, nthreads=4, use_ref=True, infeas=r0.infeas_out, zero_init=0, init_bez=r0.bez_coeff,
                       durations=dur1, **s1)
    d = dict(B=B, N=N, P_max=pb.P_max, planes=pb.planes, nplanes=pb.nplanes, durations=pb.durations, seeds=pb.seeds,
             x0=pb.x0, xd=pb.xd, max_vel=pb.max_vel, max_acc=pb.max_acc, minvo=ov.get("minvo", 0),
             time_power=ov.get("time_power", 2), dur1=dur1)
    for f in OUT_FIELDS:
        d["s0_" + f] = getattr(r0, f)
        d["s1_" + f] = getattr(r1, f)
    d["s0_jerk_sum"] = r0.jerk[:, 0]
    d["s1_jerk_sum"] = r1.jerk[:, 0]
    d["s0_terminal_norm"] = r0.x_final[:, 0]
    d["s1_terminal_norm"] = r1.x_final[:, 0]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(name, "rtn0", r0.rtn, "rtn1", r1.rtn, "iters", r0.iters, r1.iters)


def line_init_case():
    """line_init_flag = true (ddp_optimizer.cpp:195-247, :255-269, :381-388): feasible IPDDP from a straight line."""
    pb = make_batch(3, 6, "box", first=500)
    kw = dict(w_snap=1.0, w_terminal=100.0, w_time=50.0, iter_max=60)
    r = O.solve_batch(pb, nthreads=2, use_ref=True, infeas=1, zero_init=0, line_init=1, **kw)
    d = dict(B=3, N=6, P_max=pb.P_max, planes=pb.planes, nplanes=pb.nplanes, durations=pb.durations, seeds=pb.seeds,
             x0=pb.x0, xd=pb.xd, max_vel=pb.max_vel, max_acc=pb.max_acc)
    for f in OUT_FIELDS:
        d["s1_" + f] = getattr(r, f)
    np.savez_compressed(os.path.join(HERE, "box_n6_lineinit.npz"), **d)
    print("line_init rtn", r.rtn, "line_failed", r.line_failed_out, "iters", r.iters)


if __name__ == "__main__":
    if not O.ref_available():
        O.build(ref=True)
    only = sys.argv[1:]   # optional: names of the cases to (re)generate; default all
    for name, (B, N, kind, first, ov) in CASES.items():
        if not only or name in only:
            run_case(name, B, N, kind, first, ov)
    if not only or "box_n6_lineinit" in only:
        line_init_case()
