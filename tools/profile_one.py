"""One two-stage solve of the bench workload (after warm-up launches) for ncu captures:
    ncu --set full --clock-control none --import-source on -k regex:ipddp -s 1 -c 1 -o gpurun_out/prof \
        python tools/profile_one.py [--batch 4096] [--knots 100] [--precision fp64]
"""
import argparse
import sys

sys.path.insert(0, ".")
from direct_b200 import make_batch  # noqa: E402
from direct_b200.capi import Solver  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--knots", type=int, default=100)
ap.add_argument("--kind", default="box")
ap.add_argument("--precision", default="fp64")
ap.add_argument("--launches", type=int, default=2)
ap.add_argument("--blocks-per-sm", type=int, default=0)
ap.add_argument("--first", type=int, default=0, help="index of the first trajectory in the generator (547: a hard one)")
a = ap.parse_args()
pb = make_batch(a.batch, a.knots, a.kind, first=a.first)
s = Solver(0, a.precision, blocks_per_sm=a.blocks_per_sm)
for _ in range(a.launches):
    _, g = s.solve_two_stage(pb, want_stage0=False)
    st = s.stats()
    print(f"kernel {st.kernel_ms:.2f} ms, {a.batch / st.kernel_ms * 1e3:.0f} solves/s (kernel only)")
s.close()
