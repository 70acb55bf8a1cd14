#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=gpurun_out/r2y_pool.log
: > $L
for p in 0 8 16 32 64 96; do
  DIRECT_DDP_POOL=$p timeout 200 python tools/cycle_report.py --batch 4096 --tag pool${p}_4096 >> $L 2>&1
done
for p in 0 32; do
  DIRECT_DDP_POOL=$p timeout 200 python tools/cycle_report.py --batch 16384 --tag pool${p}_16384 >> $L 2>&1
done
grep "kernel\|slot busy" $L
