#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2w_pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/r2w_pytest_gpu.log
tail -3 gpurun_out/r2w_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2w_smoke.log 2>&1; echo "exit $?" >> gpurun_out/r2w_smoke.log; tail -4 gpurun_out/r2w_smoke.log
timeout 300 python tools/stress_tail.py --reps 2 > gpurun_out/r2w_stress.log 2>&1; tail -2 gpurun_out/r2w_stress.log
timeout 400 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; head -c 300 gpurun_out/r2w_bench.json
