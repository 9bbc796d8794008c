#!/bin/bash
# crowd evidence for profiles/: per-op CUDA-event timing of one eager step (B=32) and one ncu --set full capture of the
# BatchNorm-affine streaming kernel and of a trunk 1x1 GEMM.  usage: tools/gpu_crowd_evidence.sh tag
TAG=${1:-c}
mkdir -p gpurun_out
SRGAN_OPTIME=1 timeout 600 python tools/crowd_bench.py 32 bf16 1 > gpurun_out/${TAG}_crowd_optime_b32.txt 2>&1
SRGAN_NO_GRAPH=1 timeout 600 ncu --clock-control none --set full --import-source on -k regex:affine2d_kernel -s 300 -c 1 -f -o gpurun_out/${TAG}_crowd_affine \
  python tools/crowd_bench.py 32 bf16 1 > gpurun_out/${TAG}_crowd_affine.log 2>&1
SRGAN_NO_GRAPH=1 timeout 600 ncu --clock-control none --set full --import-source on -k regex:affine_grad_kernel -s 100 -c 1 -f -o gpurun_out/${TAG}_crowd_affine_grad \
  python tools/crowd_bench.py 32 bf16 1 > gpurun_out/${TAG}_crowd_affine_grad.log 2>&1
tail -3 gpurun_out/${TAG}_crowd_optime_b32.txt | cut -c1-200; ls -la gpurun_out/${TAG}_*
