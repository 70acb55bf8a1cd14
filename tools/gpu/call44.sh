#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=gpurun_out/r2z_fp32.log
: > $L
timeout 200 python tools/cycle_report.py --batch 4096 --precision fp32 --tag fp32_4096 >> $L 2>&1
timeout 200 python tools/cycle_report.py --batch 4096 --precision fp64 --tag fp64_4096 >> $L 2>&1
timeout 200 python tools/cycle_report.py --batch 16384 --precision fp32 --tag fp32_16384 >> $L 2>&1
cat $L
