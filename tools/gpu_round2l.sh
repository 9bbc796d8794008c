#!/bin/bash
# bn kernel tests, crowd parity tests at the default level and level 5, crowd bench per level
TAG=${1:-r2p}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bn_gemm.py -q -x --tb=short --timeout 600 2>&1 | tail -15 > gpurun_out/${TAG}_pytest_bn.log
tail -5 gpurun_out/${TAG}_pytest_bn.log
SRGAN_FUSE_BN=5 timeout 1200 python -m pytest tests -m gpu -q -x --tb=short --timeout 900 -k "crowd" 2>&1 | tail -8 > gpurun_out/${TAG}_pytest_crowd5.log
tail -4 gpurun_out/${TAG}_pytest_crowd5.log
LEVELS="3 5" bash tools/gpu_round2j.sh ${TAG}
