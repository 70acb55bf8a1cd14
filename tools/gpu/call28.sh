#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=gpurun_out/r2w_mingrid.log
: > $L
for g in 1 16 64 148; do
  echo "== DIRECT_DDP_MIN_GRID=$g (easy corridor, first=7)" >> $L
  DIRECT_DDP_MIN_GRID=$g timeout 200 python tools/latency_b1.py --knots 30 100 --reps 5 2>&1 | grep "^| [0-9]" >> $L
  echo "== DIRECT_DDP_MIN_GRID=$g (hard corridor, first=547)" >> $L
  DIRECT_DDP_MIN_GRID=$g timeout 200 python tools/latency_b1.py --knots 100 --reps 5 --first 547 2>&1 | grep "^| [0-9]" >> $L
  echo "== DIRECT_DDP_MIN_GRID=$g (hard corridor, first=1137)" >> $L
  DIRECT_DDP_MIN_GRID=$g timeout 200 python tools/latency_b1.py --knots 100 --reps 5 --first 1137 2>&1 | grep "^| [0-9]" >> $L
done
cat $L
