#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
nvidia-smi -L > $O/r2x_host8.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --steps 5 --warmup 3 > $O/r2x_bench_8gpu.json 2> $O/r2x_bench_8gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 8 --batch 8192 --steps 3 --warmup 3 > $O/r2x_bench_8gpu_cfg2.json 2> $O/r2x_bench_8gpu_cfg2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29633 bench.py --gpus 4 --steps 5 --warmup 3 > $O/r2x_bench_4gpu.json 2> $O/r2x_bench_4gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29634 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2x_bench_2gpu.json 2> $O/r2x_bench_2gpu.err
timeout 300 python tools/multi_gpu_capi.py --total 65536 --gpus 8 --reps 2 > $O/r2x_capi_8.jsonl 2> $O/r2x_capi_8.err
for f in 8gpu 8gpu_cfg2 4gpu 2gpu; do head -c 260 $O/r2x_bench_$f.json; echo; done; cat $O/r2x_capi_8.jsonl | head -c 600
