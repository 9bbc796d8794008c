#!/bin/bash
# N GPUs: the driver's torchrun bench command (crowd), short
N=${1:-8}
TAG=${2:-r2n$N}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --secondary none > gpurun_out/${TAG}_bench_crowd.log 2>&1
grep -o '"n_gpus": [0-9]*\|"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/${TAG}_bench_crowd.log | head -4
tail -c 300 gpurun_out/${TAG}_bench_crowd.log
