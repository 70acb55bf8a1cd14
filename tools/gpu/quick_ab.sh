#!/bin/bash
# Fail-fast A/B of kernel variants: tools/gpu/quick_ab.sh TAG lib1 lib2 ...   ("cur" = the in-tree library)
# Every step runs under its own short timeout so that a hanging kernel costs minutes, not the whole call.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=$1; shift
LOG=gpurun_out/${TAG}_ab.log
: > $LOG
for v in "$@"; do
  if [ "$v" = cur ]; then unset DIRECT_DDP_LIB; else export DIRECT_DDP_LIB=tools/_variants/lib_$v.so; fi
  echo "== $v: smoke" >> $LOG
  timeout 150 python - >> $LOG 2>&1 <<'PY'
import sys, time
sys.path.insert(0, ".")
import numpy as np
from direct_b200 import make_batch
from direct_b200.capi import Solver
from oracle import oracle_py as O
s = Solver(0, "fp64")
pb = make_batch(32, 20, "poly", first=11)
g0, g1 = s.solve_two_stage(pb)
a0, a1 = O.two_stage_batch(pb, nthreads=8)
same = (g1.rtn == a1.rtn) & (g1.iters == a1.iters) & (np.abs(g1.cost - a1.cost) <= 1e-5 * np.abs(a1.cost))
print("small batch: identical decisions + cost within 1e-5 on", int(same.sum()), "of 32; stats equal on", int((g1.stats[:, :4] == a1.stats[:, :4]).all(1).sum()))
s.close()
PY
  rc=$?; echo "smoke rc $rc" >> $LOG
  if [ $rc -ne 0 ]; then echo "== $v: SKIPPED (smoke failed or hung)" >> $LOG; continue; fi
  timeout 200 python tools/cycle_report.py --batch 4096 --tag ${v}_4096 >> $LOG 2>&1; echo "rc $?" >> $LOG
done
unset DIRECT_DDP_LIB
cat $LOG
