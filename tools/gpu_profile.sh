#!/bin/bash
# launch list (per-kernel device time) of a short bf16 bench; usage: tools/gpu_profile.sh tag
TAG=${1:-p}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 390 -c 260 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --precision bf16 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_bench.log | cut -c1-300
wc -l gpurun_out/${TAG}_launches.csv
