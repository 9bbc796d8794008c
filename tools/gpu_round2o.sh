#!/bin/bash
TAG=${1:-r2s}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bn_gemm.py tests/test_gpu_flat3x3.py -q -x --tb=short --timeout 600 2>&1 | tail -8 > gpurun_out/${TAG}_pytest_k.log
tail -4 gpurun_out/${TAG}_pytest_k.log
{ echo "== resident B"; timeout 300 python tools/bn_gemm_bench.py 10 2>&1 | grep "bn_dgrad"; echo "== streamed B"; SRGAN_NO_RESIDENT_B=1 timeout 300 python tools/bn_gemm_bench.py 10 2>&1 | grep "bn_dgrad"; } > gpurun_out/${TAG}_bn_dgrad_resb.txt
cat gpurun_out/${TAG}_bn_dgrad_resb.txt
timeout 1200 python -m pytest tests -m gpu -q -x --tb=short --timeout 900 -k "crowd or window" 2>&1 | tail -8 > gpurun_out/${TAG}_pytest_crowd.log
tail -4 gpurun_out/${TAG}_pytest_crowd.log
for v in "X=0" "SRGAN_NO_FLAT3X3=1" "SRGAN_NO_RESIDENT_B=1"; do
echo "== $v"; env $v timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --secondary none 2>&1 | grep -o '"ms_per_step": [0-9.]*'
done
