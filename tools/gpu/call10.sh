#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
bash tools/gpu/quick_ab.sh r2k cur > /dev/null 2>&1
O=gpurun_out
timeout 120 python tools/cycle_report.py --batch 16384 --tag cur_16384 >> $O/r2k_ab.log 2>&1
grep -v "cooperation\|smoke\|small batch\|^rc" $O/r2k_ab.log
