#!/bin/bash
# href landing-zone prefetch: kernel tests + trunk GEMM A/B + crowd bench
TAG=${1:-r2d}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short --timeout 600 -x 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
for H in 1 0; do
  for shape in "50176 1024" "200704 512" "12544 1536" "802816 256"; do
    SRGAN_NO_HREF_SMEM=$H timeout 300 python tools/trunk_gemm_bench.py $shape 10 2>&1 | grep "data grad" | sed "s/^/nohrefsmem=$H /" >> gpurun_out/${TAG}_trunk_gemm.txt
  done
done
cat gpurun_out/${TAG}_trunk_gemm.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short --timeout 600 -k "crowd or age_full" 2>&1 | tail -8 >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --secondary none > gpurun_out/${TAG}_bench.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/${TAG}_bench.log | head -1
grep -o '"roofline": {[^}]*}' gpurun_out/${TAG}_bench.log | cut -c1-300
