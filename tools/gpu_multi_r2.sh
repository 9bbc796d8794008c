#!/bin/bash
# N GPUs of one box: the NCCL parity tests, then the driver's torchrun bench command at N (crowd default, age secondary off)
N=${1:-2}
TAG=${2:-r2n$N}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
timeout 1200 python -m pytest tests/test_gpu_dist.py -m gpu -q --tb=short --timeout 900 2>&1 | tail -6 > gpurun_out/${TAG}_pytest_dist.log
tail -3 gpurun_out/${TAG}_pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --secondary none > gpurun_out/${TAG}_bench_crowd.log 2>&1
tail -c 1500 gpurun_out/${TAG}_bench_crowd.log
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --secondary none 2>&1 | grep -o '"ms_per_step": [0-9.]*'
