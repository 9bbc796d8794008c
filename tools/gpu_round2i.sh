#!/bin/bash
# fused dense-layer kernels: kernel tests, micro-bench, crowd parity tests, crowd bench with SRGAN_FUSE_BN = 1 / 2
TAG=${1:-r2k}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bn_gemm.py -q -x --tb=short --timeout 600 2>&1 | tail -15 > gpurun_out/${TAG}_pytest_bn.log
tail -5 gpurun_out/${TAG}_pytest_bn.log
timeout 600 python tools/bn_gemm_bench.py 10 > gpurun_out/${TAG}_bn_bench.txt 2>&1; cat gpurun_out/${TAG}_bn_bench.txt
timeout 1200 python -m pytest tests -m gpu -q -x --tb=short --timeout 900 -k "crowd" 2>&1 | tail -8 > gpurun_out/${TAG}_pytest_crowd.log
tail -4 gpurun_out/${TAG}_pytest_crowd.log
for f in 2 3 4; do
SRGAN_FUSE_BN=$f timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --secondary none > gpurun_out/${TAG}_bench_fuse$f.log 2>&1
done
grep -o '"ms_per_step": [0-9.]*' gpurun_out/${TAG}_bench_fuse*.log
timeout 500 ncu --clock-control none --set full --import-source on -k regex:bn_conv_down -s 2 -c 1 -f -o gpurun_out/${TAG}_bn_conv_down python tools/bn_gemm_bench.py 2 > gpurun_out/${TAG}_ncu.log 2>&1; tail -2 gpurun_out/${TAG}_ncu.log
