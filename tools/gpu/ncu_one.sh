#!/bin/bash
# one ncu --set full capture of the solve kernel (second launch) + launch list; TAG = output prefix
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-r2d}; shift
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ipddp -s 1 -c 1 -o gpurun_out/prof_$TAG -f python tools/profile_one.py "$@" > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc $?" >> gpurun_out/ncu_$TAG.log
tail -5 gpurun_out/ncu_$TAG.log
