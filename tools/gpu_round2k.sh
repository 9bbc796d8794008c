#!/bin/bash
# bn_conv_down sensitivity: stages / MT / no-transform
TAG=${1:-r2n}
run() { echo "== $*"; env "$@" timeout 300 python tools/bn_gemm_bench.py 10 2>&1 | grep "bn_conv_down  \|unfused affine\|rows=  50176.*weight\|rows=  50176.*wgrad"; }
{
run X=0
run SRGAN_BF_STAGES=4
run SRGAN_BF_STAGES=2
run SRGAN_BF_MT=2
run SRGAN_BF_MT=1
run SRGAN_BF_NOXF=1
run SRGAN_BF_NOXF=1 SRGAN_BF_MT=2
} > gpurun_out/${TAG}_fprop_sens.txt 2>&1
cat gpurun_out/${TAG}_fprop_sens.txt
