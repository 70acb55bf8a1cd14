#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
LOG=gpurun_out/r2b_dbg.log
: > $LOG
for v in dbg cur; do
if [ "$v" = cur ]; then unset DIRECT_DDP_LIB; else export DIRECT_DDP_LIB=tools/_variants/lib_$v.so; fi
for cfg in "1 5 box" "4 5 box" "32 20 poly" "512 100 box"; do
  set -- $cfg
  echo "== $v: B=$1 N=$2 $3" >> $LOG
  timeout 60 python - $1 $2 $3 >> $LOG 2>&1 <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np
from direct_b200 import make_batch
from direct_b200.capi import Solver
from oracle import oracle_py as O
B, N, kind = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
s = Solver(0, "fp64")
pb = make_batch(B, N, kind, first=11)
g0, g1 = s.solve_two_stage(pb)
st = s.stats()
a0, a1 = O.two_stage_batch(pb, nthreads=16)
same = (g1.rtn == a1.rtn) & (g1.iters == a1.iters) & (np.abs(g1.cost - a1.cost) <= 1e-5 * np.abs(a1.cost))
print(f"kernel {st.kernel_ms:.2f} ms; identical decisions + cost within 1e-5 on {int(same.sum())} of {B}; stats equal on {int((g1.stats[:, :4] == a1.stats[:, :4]).all(1).sum())}")
s.close()
PY
  echo "rc $?" >> $LOG
done
done
cat $LOG
