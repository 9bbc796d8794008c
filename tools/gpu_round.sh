#!/bin/bash
# One GPU-box round: unit + parity tests, smoke, short benches.  Usage: tools/gpu_round.sh [tag]
TAG=${1:-r}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --tb=short --timeout 600 2>&1 | tail -150 > gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1
timeout 600 python bench.py --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_fp32.log 2>&1
timeout 600 python bench.py --precision bf16 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_bf16.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_smoke.log; tail -2 gpurun_out/${TAG}_bench_fp32.log; tail -2 gpurun_out/${TAG}_bench_bf16.log
