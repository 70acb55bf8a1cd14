#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 tools/gather_bench.py > $O/r2y_gather2.log 2>&1
grep ranks $O/r2y_gather2.log || tail -5 $O/r2y_gather2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2y_bench_2gpu.json 2> $O/r2y_bench_2gpu.err
head -c 260 $O/r2y_bench_2gpu.json; echo; tail -2 $O/r2y_bench_2gpu.err
