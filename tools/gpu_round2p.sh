#!/bin/bash
TAG=${1:-r2t}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bn_gemm.py -q -x --tb=short --timeout 600 2>&1 | tail -8 > gpurun_out/${TAG}_pytest_k.log
tail -4 gpurun_out/${TAG}_pytest_k.log
timeout 300 python tools/bn_gemm_bench.py 10 2>&1 | grep "bn_dgrad" > gpurun_out/${TAG}_bn_dgrad.txt
cat gpurun_out/${TAG}_bn_dgrad.txt
timeout 1200 python -m pytest tests -m gpu -q -x --tb=short --timeout 900 -k "crowd" 2>&1 | tail -8 > gpurun_out/${TAG}_pytest_crowd.log
tail -4 gpurun_out/${TAG}_pytest_crowd.log
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --secondary none 2>&1 | grep -o '"ms_per_step": [0-9.]*'
