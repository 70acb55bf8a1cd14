"""Kernel time of single-corridor solves (B = 1, 100 knots): two hard corridors (stage 1 runs to iter_max) and an easy one.
    [DIRECT_DDP_LIB=...] python tools/b1_kernel.py"""
import os
import sys

sys.path.insert(0, ".")
from direct_b200 import make_batch  # noqa: E402
from direct_b200.capi import Solver  # noqa: E402

s = Solver(0, "fp64")
for B, first in ((1, 547), (1, 1137), (1, 7)):
    pb = make_batch(B, 100, "box", first=first)
    best = 1e9
    for _ in range(5):
        s.solve_two_stage(pb)
        best = min(best, s.stats().kernel_ms)
    print(f"{os.environ.get('DIRECT_DDP_LIB', 'in-tree library')}: B {B} first {first}: kernel {best:.2f} ms", flush=True)
s.close()
