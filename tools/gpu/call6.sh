#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
bash tools/gpu/quick_ab.sh r2g cur > /dev/null 2>&1
O=gpurun_out
timeout 120 python tools/cycle_report.py --batch 16384 --tag cur_16384 >> $O/r2g_ab.log 2>&1
timeout 120 python tools/cycle_report.py --batch 4096 --kind poly --tag cur_poly >> $O/r2g_ab.log 2>&1
cat $O/r2g_ab.log
