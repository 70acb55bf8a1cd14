"""Speculative line search (DIRECT_DDP_GSPEC) or speculative backward sweep (DIRECT_DDP_SPEC) on/off: same bits, different time.
    python tools/spec_check.py [--batch 4096] [--knob DIRECT_DDP_SPEC]"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, ".")
from direct_b200 import make_batch  # noqa: E402
from direct_b200.capi import Solver  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--knots", type=int, default=100)
ap.add_argument("--kind", default="box")
ap.add_argument("--knob", default="DIRECT_DDP_GSPEC")
a = ap.parse_args()
pb = make_batch(a.batch, a.knots, a.kind)
s = Solver(0, "fp64")
res = {}
for mode in ("0", "1", "0", "1"):
    os.environ[a.knob] = mode
    g0, g1 = s.solve_two_stage(pb, want_stage0=True)
    st = s.stats()
    print(f"{a.knob}={mode}: kernel {st.kernel_ms:.1f} ms = {a.batch / st.kernel_ms * 1e3:.0f} solves/s; searches posted {st.spec_searches}, "
          f"remote trials {st.spec_trials}, fwd trials {st.fwd_trials}, fwd knots {st.fwd_knots}", flush=True)
    res.setdefault(mode, []).append((g0, g1))
(a0, a1), (b0, b1) = res["0"][0], res["1"][0]
same = True
for f in ("rtn", "iters", "cost", "poly_coeff", "bez_coeff", "poly_time", "x_final", "jerk"):
    for x, y in ((a0, b0), (a1, b1)):
        if not np.array_equal(getattr(x, f), getattr(y, f)):
            same = False
            d = np.argwhere(getattr(x, f) != getattr(y, f))
            print("DIFF", f, len(d), d[:3].tolist())
for x, y in ((a0, b0), (a1, b1)):
    if not np.array_equal(x.stats[:, :4], y.stats[:, :4]):
        same = False
        print("DIFF counters", np.argwhere(x.stats[:, :4] != y.stats[:, :4])[:5].tolist())
print("bit-identical with and without speculation:", same)
s.close()
