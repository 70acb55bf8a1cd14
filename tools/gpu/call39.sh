#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=gpurun_out/r2z_flat.log
: > $L
timeout 200 python tools/compare_builds.py tools/_variants/lib_prev.so tools/_variants/lib_flat.so >> $L 2>&1
timeout 200 python tools/compare_builds.py tools/_variants/lib_prev.so tools/_variants/lib_flat.so --batch 64 --knots 100 --kind poly >> $L 2>&1
for rep in 1 2; do
for v in pipe flat; do
  DIRECT_DDP_LIB=tools/_variants/lib_$v.so timeout 100 python tools/b1_kernel.py >> $L 2>&1
  DIRECT_DDP_LIB=tools/_variants/lib_$v.so timeout 200 python tools/cycle_report.py --batch 4096 --tag ${v}_4096 >> $L 2>&1
done; done
for v in pipe flat; do DIRECT_DDP_LIB=tools/_variants/lib_$v.so timeout 200 python tools/cycle_report.py --batch 16384 --tag ${v}_16384 >> $L 2>&1; done
DIRECT_DDP_LIB=tools/_variants/lib_flattrc.so timeout 200 python tools/timeline.py --cycles --batch 1 --first 547 > gpurun_out/r2z_timeline_b1_flat.log 2>&1
grep -v "slot busy\|cooperation\|cycles per\|Riccati " $L; sed -n 41,44p gpurun_out/r2z_timeline_b1_flat.log
