#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
bash tools/gpu/quick_ab.sh r2j cur s2c1 s1c2 s2c2 > /dev/null 2>&1
O=gpurun_out
for v in s2c1 s2c2; do DIRECT_DDP_LIB=tools/_variants/lib_$v.so timeout 120 python tools/cycle_report.py --batch 16384 --tag ${v}_16384 >> $O/r2j_ab.log 2>&1; done
grep -v "cooperation\|smoke\|small batch\|^rc" $O/r2j_ab.log
