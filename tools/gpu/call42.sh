#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=gpurun_out/r2z_idle.log
: > $L
for rep in 1 2; do
for v in idle4000 idle1000 idle500; do
  DIRECT_DDP_LIB=tools/_variants/lib_$v.so timeout 100 python tools/b1_kernel.py >> $L 2>&1
  DIRECT_DDP_LIB=tools/_variants/lib_$v.so timeout 200 python tools/cycle_report.py --batch 4096 --tag ${v}_4096 >> $L 2>&1
done; done
grep -v "slot busy\|cooperation\|cycles per\|Riccati " $L
