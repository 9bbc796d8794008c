#!/bin/bash
TAG=${1:-r2r}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_flat3x3.py -q -x --tb=short --timeout 300 2>&1 | tail -15 > gpurun_out/${TAG}_pytest_flat.log
tail -8 gpurun_out/${TAG}_pytest_flat.log
{ echo "== flat"; timeout 300 python tools/flat3x3_bench.py 10; echo "== tap-per-stage"; SRGAN_NO_FLAT3X3=1 timeout 300 python tools/flat3x3_bench.py 10; } > gpurun_out/${TAG}_flat_bench.txt 2>&1
cat gpurun_out/${TAG}_flat_bench.txt
