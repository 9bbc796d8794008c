#!/bin/bash
# two-chain D step: targeted tests + crowd bench A/B
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --tb=short --timeout 900 -k "crowd or two_chain or mixin or age_full or dist" 2>&1 | tail -40 > gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --secondary none > gpurun_out/${TAG}_bench_split.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/${TAG}_bench_split.log | head -1 | sed "s/^/split /"
SRGAN_NO_SPLIT_CHAINS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --secondary none > gpurun_out/${TAG}_bench_nosplit.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/${TAG}_bench_nosplit.log | head -1 | sed "s/^/nosplit /"
SRGAN_NO_OVERLAP=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --secondary none > gpurun_out/${TAG}_bench_split_nodnn.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/${TAG}_bench_split_nodnn.log | head -1 | sed "s/^/split, dnn serial /"
tail -3 gpurun_out/${TAG}_bench_split.log | cut -c1-600
