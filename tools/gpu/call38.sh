#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=gpurun_out/r2z_corr.log
: > $L
timeout 200 python tools/compare_builds.py tools/_variants/lib_prev.so tools/_variants/lib_corr.so >> $L 2>&1
timeout 200 python tools/compare_builds.py tools/_variants/lib_prev.so tools/_variants/lib_corr.so --batch 64 --knots 100 --kind poly >> $L 2>&1
cat > /tmp/b1.py <<'PY'
import os, sys
sys.path.insert(0, ".")
from direct_b200 import make_batch
from direct_b200.capi import Solver
s = Solver(0, "fp64")
for B, first in ((1, 547), (1, 1137), (1, 7)):
    pb = make_batch(B, 100, "box", first=first)
    best = 1e9
    for _ in range(5):
        s.solve_two_stage(pb); best = min(best, s.stats().kernel_ms)
    print(f"{os.environ.get('DIRECT_DDP_LIB','cur')}: B {B} first {first}: kernel {best:.2f} ms", flush=True)
s.close()
PY
for rep in 1 2; do
for v in prev corr; do
  DIRECT_DDP_LIB=tools/_variants/lib_$v.so timeout 100 python /tmp/b1.py >> $L 2>&1
  DIRECT_DDP_LIB=tools/_variants/lib_$v.so timeout 200 python tools/cycle_report.py --batch 4096 --tag ${v}_4096 >> $L 2>&1
done; done
for v in prev corr; do DIRECT_DDP_LIB=tools/_variants/lib_$v.so timeout 200 python tools/cycle_report.py --batch 16384 --tag ${v}_16384 >> $L 2>&1; done
DIRECT_DDP_LIB=tools/_variants/lib_corrtrc.so timeout 200 python tools/timeline.py --cycles --batch 1 --first 547 > gpurun_out/r2z_timeline_b1_corr.log 2>&1
grep -v "slot busy\|cooperation\|cycles per\|Riccati " $L; sed -n 41,44p gpurun_out/r2z_timeline_b1_corr.log
