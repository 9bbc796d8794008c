#!/bin/bash
# N GPUs: crowd bench with / without the multi-rank DNN || GAN overlap (same seeds: the last_scalars must agree)
N=${1:-2}
TAG=${2:-r2n$N}
mkdir -p gpurun_out
for v in "X=0" "SRGAN_NO_OVERLAP_MULTI=1"; do
env $v timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --secondary none > gpurun_out/${TAG}_bench_${v%%=*}.log 2>&1
echo "== $v"; grep -o '"ms_per_step": [0-9.]*\|"last_scalars": {[^}]*}' gpurun_out/${TAG}_bench_${v%%=*}.log | head -3
done
