#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
bash tools/gpu/quick_ab.sh r2c base cur > /dev/null 2>&1
timeout 120 python tools/cycle_report.py --batch 16384 --tag cur_16384 >> $O/r2c_ab.log 2>&1
timeout 120 python tools/cycle_report.py --batch 4096 --kind poly --tag cur_poly >> $O/r2c_ab.log 2>&1
timeout 400 python -m pytest tests -x -q -m gpu > $O/r2c_pytest_all.log 2>&1; echo "all exit $?" >> $O/r2c_pytest_all.log
timeout 300 python bench.py > $O/r2c_bench.json 2> $O/r2c_bench.err
cat $O/r2c_ab.log; tail -5 $O/r2c_pytest_all.log; head -c 700 $O/r2c_bench.json
