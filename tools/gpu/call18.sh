#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=gpurun_out/r2u_spec3.log
: > $L
timeout 150 python tools/spec_check.py --batch 256 --knots 40 --knob DIRECT_DDP_SPEC >> $L 2>&1; echo "small rc $?" >> $L
if ! grep -q "bit-identical with and without speculation: True" $L; then cat $L; exit 0; fi
timeout 300 python tools/spec_check.py --knob DIRECT_DDP_SPEC >> $L 2>&1; echo "full rc $?" >> $L
timeout 200 python tools/cycle_report.py --batch 4096 --tag spec_4096 >> $L 2>&1
timeout 200 python tools/timeline.py > gpurun_out/r2u_timeline.log 2>&1
grep -v "DIFF" $L | tail -14; sed -n 1,1p gpurun_out/r2u_timeline.log; sed -n 60,104p gpurun_out/r2u_timeline.log
