#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
bash tools/gpu/quick_ab.sh r2e base noring cur > /dev/null 2>&1
O=gpurun_out
for v in noring cur; do
  if [ "$v" = cur ]; then unset DIRECT_DDP_LIB; else export DIRECT_DDP_LIB=tools/_variants/lib_$v.so; fi
  timeout 120 python tools/cycle_report.py --batch 16384 --tag ${v}_16384 >> $O/r2e_ab.log 2>&1
done
unset DIRECT_DDP_LIB
timeout 200 python tools/tail_report.py > $O/r2e_tail.log 2>&1
cat $O/r2e_ab.log; tail -22 $O/r2e_tail.log
