#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=gpurun_out/r2v_excl.log
: > $L
DIRECT_DDP_LIB=tools/_variants/lib_excl256.so timeout 150 python tools/spec_check.py --batch 256 --knots 40 --knob DIRECT_DDP_GSPEC >> $L 2>&1; echo "small rc $?" >> $L
if ! grep -q "bit-identical with and without speculation: True" $L; then cat $L; exit 0; fi
for rep in 1 2; do
for v in excl100000 excl256 excl0; do
  DIRECT_DDP_LIB=tools/_variants/lib_$v.so timeout 200 python tools/cycle_report.py --batch 4096 --tag ${v}_4096 >> $L 2>&1
done; done
for v in excl100000 excl256 excl0; do
  DIRECT_DDP_LIB=tools/_variants/lib_$v.so timeout 200 python tools/cycle_report.py --batch 1024 --tag ${v}_1024 >> $L 2>&1
done
DIRECT_DDP_LIB=tools/_variants/lib_excl256.so timeout 200 python tools/timeline.py > gpurun_out/r2v_timeline.log 2>&1
grep "kernel\|bit-ident" $L; sed -n 75,103p gpurun_out/r2v_timeline.log
