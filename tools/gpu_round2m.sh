#!/bin/bash
# crowd bench at the default fusion level + ncu launch list of one eager crowd step
TAG=${1:-r2q}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --secondary none > gpurun_out/${TAG}_bench.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/${TAG}_bench.log | head -1
NCU="ncu --clock-control none"
timeout 900 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${TAG}_crowd_launches.csv \
  python tools/crowd_step_profile.py 64 > gpurun_out/${TAG}_crowd_profile.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_crowd_launches.csv > gpurun_out/${TAG}_crowd_launches_summary.txt 2>&1
head -45 gpurun_out/${TAG}_crowd_launches_summary.txt
