#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
T=r2z
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${T}_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ipddp -s 1 -c 1 -o $O/prof_$T -f python tools/profile_one.py > $O/ncu_$T.log 2>&1
tail -2 $O/ncu_$T.log; wc -l $O/${T}_launches.csv
bash tools/gpu/final1.sh r2z
