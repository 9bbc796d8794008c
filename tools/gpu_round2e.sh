#!/bin/bash
# swapped-role conv2 wgrad: kernel tests, crowd parity, bench A/B, ncu of the dgrad GEMM after the href prefetch
TAG=${1:-r2e}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short --timeout 600 -x -k "windows or conv_down_up" 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short --timeout 600 -k "crowd" 2>&1 | tail -8 >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --secondary none > gpurun_out/${TAG}_bench.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/${TAG}_bench.log | head -1 | sed "s/^/swap /"
SRGAN_NO_WGRAD_SWAP=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --secondary none > gpurun_out/${TAG}_bench_noswap.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/${TAG}_bench_noswap.log | head -1 | sed "s/^/noswap /"
NCU="ncu --clock-control none"
timeout 600 $NCU --set full --import-source on -k regex:umma_conv_persistent -s 4 -c 1 -f -o gpurun_out/${TAG}_trunk_dgrad \
  python tools/trunk_gemm_bench.py 50176 1024 2 > gpurun_out/${TAG}_ncu_trunk.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_trunk.log | cut -c1-200
timeout 900 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${TAG}_crowd_launches.csv \
  python tools/crowd_step_profile.py 64 > gpurun_out/${TAG}_crowd_profile.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_crowd_launches.csv > gpurun_out/${TAG}_crowd_launches_summary.txt 2>&1
head -14 gpurun_out/${TAG}_crowd_launches_summary.txt
