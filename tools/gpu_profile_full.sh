#!/bin/bash
# ncu evidence for profiles/: (1) launch list of the age bf16 step (eager launches, so every kernel is a separate row),
# (2) one `--set full` capture of each dominant kernel.  usage: tools/gpu_profile_full.sh tag
TAG=${1:-p}
mkdir -p gpurun_out
NCU="ncu --clock-control none"
SRGAN_NO_GRAPH=1 timeout 900 $NCU --metrics gpu__time_duration.sum -s 440 -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --precision bf16 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
full() {  # name kernel-regex skip cmd...
  local name=$1 k=$2 skip=$3; shift 3
  timeout 600 $NCU --set full --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/${TAG}_${name} "$@" > gpurun_out/${TAG}_${name}.log 2>&1
  tail -2 gpurun_out/${TAG}_${name}.log | cut -c1-200
}
full conv_l2_fprop 'umma_conv_persistent' 3 python tools/conv_bench.py D.l2 down --iters 2
full conv_l4_fprop 'umma_conv_persistent' 3 python tools/conv_bench.py D.l4 down --iters 2
full wgrad_l2 'umma_wgrad' 3 python tools/conv_bench.py D.l2 wgrad --iters 2
SRGAN_NO_GRAPH=1 full colsum 'colsum_kernel' 40 python bench.py --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline
SRGAN_NO_GRAPH=1 full adam 'adam_kernel' 60 python bench.py --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline
full coef_step 'coef_step_kernel' 7 python tools/coef_bench.py srgan
ls -la gpurun_out/${TAG}_*
