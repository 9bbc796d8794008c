#!/bin/bash
# crowd bench per fusion level; on failure re-run eagerly with blocking launches to name the failing kernel
TAG=${1:-r2m}
mkdir -p gpurun_out
for f in ${LEVELS:-2 3 4}; do
SRGAN_FUSE_BN=$f timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --secondary none > gpurun_out/${TAG}_bench_fuse$f.log 2>&1
if ! grep -q ms_per_step gpurun_out/${TAG}_bench_fuse$f.log; then
  echo "fuse=$f FAILED; blocking re-run"
  CUDA_LAUNCH_BLOCKING=1 SRGAN_NO_GRAPH=1 SRGAN_FUSE_BN=$f timeout 600 python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --secondary none > gpurun_out/${TAG}_bench_fuse${f}_blocking.log 2>&1
  grep -v "^frame" gpurun_out/${TAG}_bench_fuse${f}_blocking.log | grep -i "error\|failed\|timeout" | head -5
fi
done
grep -o '"ms_per_step": [0-9.]*' gpurun_out/${TAG}_bench_fuse*.log
