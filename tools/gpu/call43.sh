#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=gpurun_out/r2z_rem.log
: > $L
timeout 200 python tools/compare_builds.py tools/_variants/lib_prev.so tools/_variants/lib_remq.so --batch 256 --knots 100 >> $L 2>&1
timeout 200 python tools/compare_builds.py tools/_variants/lib_prev.so tools/_variants/lib_remq.so --batch 64 --knots 200 --kind poly >> $L 2>&1
timeout 200 python tools/compare_builds.py tools/_variants/lib_prev.so tools/_variants/lib_remq.so --batch 300 --knots 37 --kind poly >> $L 2>&1
for rep in 1 2; do
for v in prev remq; do
  DIRECT_DDP_LIB=tools/_variants/lib_$v.so timeout 200 python tools/cycle_report.py --batch 4096 --tag ${v}_4096 >> $L 2>&1
done; done
for v in prev remq; do DIRECT_DDP_LIB=tools/_variants/lib_$v.so timeout 200 python tools/cycle_report.py --batch 16384 --tag ${v}_16384 >> $L 2>&1; done
for v in prev remq; do DIRECT_DDP_LIB=tools/_variants/lib_$v.so timeout 200 python tools/cycle_report.py --batch 4096 --knots 96 --tag ${v}_4096_n96 >> $L 2>&1; done
grep -v "slot busy\|cooperation\|cycles per" $L
