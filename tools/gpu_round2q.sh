#!/bin/bash
TAG=${1:-r2v}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bn_gemm.py -q -x --tb=short --timeout 600 2>&1 | tail -4
timeout 300 python tools/bn_gemm_bench.py 10 2>&1 | grep "bn_conv\|weight grad" > gpurun_out/${TAG}_bn_bench.txt; cat gpurun_out/${TAG}_bn_bench.txt
for v in "SRGAN_FUSE_BN=0" "SRGAN_BN_EXACT_SCALE=1" "X=0"; do
echo "== $v"; env $v timeout 600 python tools/bf16_parity_report.py 2>&1 | grep -A12 "precision mode bf16" | grep "crowd\|age" | cut -c1-330
done > gpurun_out/${TAG}_parity_variants.txt 2>&1
cat gpurun_out/${TAG}_parity_variants.txt
for v in "X=0" "SRGAN_FUSE_BN=5" "SRGAN_BN_EXACT_SCALE=1"; do
echo "== $v"; env $v timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --secondary none 2>&1 | grep -o '"ms_per_step": [0-9.]*'
done
