#!/bin/bash
# Round-2 GPU-box pass: full -m gpu suite, smoke, the driver's bench commands (product arm = crowd + secondary age; reference arm).
# Usage: tools/gpu_round2.sh [tag] [pytest args...]
TAG=${1:-r2}
shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt
timeout 2400 python -m pytest tests -m gpu -q --tb=short --timeout 900 -x "$@" 2>&1 | tail -150 > gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.log 2>&1
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_ref.log 2>&1
tail -8 gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_smoke.log; tail -c 3000 gpurun_out/${TAG}_bench.log; tail -c 1500 gpurun_out/${TAG}_bench_ref.log
