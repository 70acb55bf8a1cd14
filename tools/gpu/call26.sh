#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=gpurun_out/r2v_sweep2.log
: > $L
timeout 150 python tools/spec_check.py --batch 256 --knots 40 --knob DIRECT_DDP_SPEC >> $L 2>&1; echo "small rc $?" >> $L
if ! grep -q "bit-identical with and without speculation: True" $L; then cat $L; exit 0; fi
timeout 300 python tools/spec_check.py --knob DIRECT_DDP_SPEC >> $L 2>&1; echo "full rc $?" >> $L
timeout 200 python tools/cycle_report.py --batch 4096 --tag sweep2_4096 >> $L 2>&1
timeout 200 python tools/cycle_report.py --batch 1024 --tag sweep2_1024 >> $L 2>&1
DIRECT_DDP_LIB=tools/_variants/lib_trc.so timeout 200 python tools/timeline.py --cycles > gpurun_out/r2v_timeline_cycles2.log 2>&1
grep -v DIFF $L | tail -16; sed -n 1,2p gpurun_out/r2v_timeline_cycles2.log; sed -n 75,103p gpurun_out/r2v_timeline_cycles2.log
