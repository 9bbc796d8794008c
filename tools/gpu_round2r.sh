#!/bin/bash
# final-state check: bn kernel tests, crowd parity, smoke, bench (exact / packed scale)
TAG=${1:-r2w}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bn_gemm.py tests/test_gpu_flat3x3.py -q -x --tb=short --timeout 600 2>&1 | tail -3
timeout 1200 python -m pytest tests -m gpu -q -x --tb=short --timeout 900 -k "crowd" 2>&1 | tail -3
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log | cut -c1-300
for v in "X=0" "SRGAN_BN_PACKED_SCALE=1"; do
echo "== $v"; env $v timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --secondary none 2>&1 | grep -o '"ms_per_step": [0-9.]*'
done
