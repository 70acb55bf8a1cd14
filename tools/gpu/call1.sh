#!/bin/bash
# round 2, GPU call 1: parity of the TMA-ring build, base vs new cycle accounting, full gpu test suite, bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r2a_host.txt; nproc >> $O/r2a_host.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $O/r2a_pytest_parity.log 2>&1; echo "parity exit $?" >> $O/r2a_pytest_parity.log
timeout 300 python tools/spec_check.py > $O/r2a_spec.log 2>&1
for B in 4096 16384; do
  DIRECT_DDP_LIB=tools/_variants/lib_base.so timeout 300 python tools/cycle_report.py --batch $B --tag base_$B >> $O/r2a_cyc.log 2>&1
  timeout 300 python tools/cycle_report.py --batch $B --tag tma1_$B >> $O/r2a_cyc.log 2>&1
done
DIRECT_DDP_LIB=tools/_variants/lib_base.so timeout 300 python tools/cycle_report.py --batch 4096 --kind poly --tag base_poly >> $O/r2a_cyc.log 2>&1
timeout 300 python tools/cycle_report.py --batch 4096 --kind poly --tag tma1_poly >> $O/r2a_cyc.log 2>&1
timeout 300 python tools/tail_report.py > $O/r2a_tail.log 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r2a_pytest_all.log 2>&1; echo "all exit $?" >> $O/r2a_pytest_all.log
timeout 600 python bench.py > $O/r2a_bench.json 2> $O/r2a_bench.err
tail -3 $O/r2a_pytest_parity.log; cat $O/r2a_cyc.log; tail -3 $O/r2a_pytest_all.log; head -c 600 $O/r2a_bench.json
