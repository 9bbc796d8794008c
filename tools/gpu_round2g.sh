#!/bin/bash
# full -m gpu suite, smoke, bf16 parity error report, the driver's two bench commands
TAG=${1:-r2g}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --tb=short --timeout 900 2>&1 | tail -30 > gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -3 gpurun_out/${TAG}_smoke.log | cut -c1-300
timeout 900 python tools/bf16_parity_report.py > gpurun_out/${TAG}_bf16_parity_errors.txt 2>&1; cat gpurun_out/${TAG}_bf16_parity_errors.txt | cut -c1-400
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.log 2>&1
tail -c 1200 gpurun_out/${TAG}_bench.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_ref.log 2>&1
tail -c 600 gpurun_out/${TAG}_bench_ref.log
