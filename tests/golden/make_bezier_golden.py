"""Generate tests/golden/bezier_ref.npz from the REFERENCE's own Bernstein evaluators.

    make -C oracle bezier && PYTHONPATH=. python tests/golden/make_bezier_golden.py

Inputs: Bezier coefficients and segment times of solved trajectories (the oracle's two-stage result of a small synthetic
batch, i.e. exactly what getBezCoeff()/getPolyTime() hand to the sampling code in teach_repeat_planner.cpp:1514-1569);
outputs: position / velocity / acceleration at S = 9 parameters per segment as computed by
/root/reference/global_planner/include/global_planner/utils/bezier_base.h:77-115 compiled unmodified
(oracle/_ref/libbezier_ref.so).  The GPU kernel and the numpy restatement are compared against this file.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from direct_b200.problems import make_batch  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    O.build(ref=True)
    pb = make_batch(6, 12, "poly", first=77)
    _, r = O.two_stage_batch(pb, nthreads=4)
    S = 9
    pos, vel, acc = O.bezier_sample_ref(r.bez_coeff, r.poly_time, S)
    np.savez_compressed(os.path.join(HERE, "bezier_ref.npz"), bez_coeff=r.bez_coeff, poly_time=r.poly_time, S=S,
                        pos=pos, vel=vel, acc=acc)
    print("bezier_ref.npz:", pos.shape, float(np.abs(pos).max()), float(np.abs(vel).max()), float(np.abs(acc).max()))
