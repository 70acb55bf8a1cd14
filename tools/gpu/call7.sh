#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
bash tools/gpu/quick_ab.sh r2h noring cur d8 d2 > /dev/null 2>&1
O=gpurun_out
timeout 120 python tools/cycle_report.py --batch 16384 --tag cur_16384 >> $O/r2h_ab.log 2>&1
DIRECT_DDP_LIB=tools/_variants/lib_d8.so timeout 120 python tools/cycle_report.py --batch 16384 --tag d8_16384 >> $O/r2h_ab.log 2>&1
grep -v "cooperation\|smoke\|small batch\|^rc" $O/r2h_ab.log
