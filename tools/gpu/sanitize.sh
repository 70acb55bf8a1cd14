#!/bin/bash
# compute-sanitizer over a small pass of the solver: tools/gpu/sanitize.sh [tools...]   (default: memcheck racecheck synccheck initcheck)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TOOLS=${@:-memcheck racecheck synccheck initcheck}
for t in $TOOLS; do
  LOG=gpurun_out/sanitize_$t.log
  timeout 420 compute-sanitizer --tool $t --print-limit 20 python tools/sanitize_run.py 24 > $LOG 2>&1
  echo "rc $?" >> $LOG
  echo "== $t"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|rc |two-stage|done|Invalid|hazard|Uninit" $LOG | head -20
done
