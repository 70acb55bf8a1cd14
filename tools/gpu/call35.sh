#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=gpurun_out/r2y_hungry.log
: > $L
for rep in 1 2; do
for v in h0 h12 h12a600 h20a600; do
  DIRECT_DDP_LIB=tools/_variants/lib_$v.so timeout 200 python tools/cycle_report.py --batch 4096 --tag ${v}_4096 >> $L 2>&1
done; done
for v in h0 h12a600; do
  DIRECT_DDP_LIB=tools/_variants/lib_$v.so timeout 200 python tools/cycle_report.py --batch 1024 --tag ${v}_1024 >> $L 2>&1
  DIRECT_DDP_LIB=tools/_variants/lib_$v.so timeout 200 python tools/cycle_report.py --batch 16384 --tag ${v}_16384 >> $L 2>&1
done
grep "kernel" $L
