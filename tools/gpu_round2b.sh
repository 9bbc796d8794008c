#!/bin/bash
# Round-2 experiments: full -m gpu suite; trunk GEMM A/B (L2 promotion of the activation tensor maps); crowd step A/B;
# ncu launch list of one eager crowd step; ncu --set full (+ source) of the trunk data-gradient GEMM.
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --tb=short --timeout 900 2>&1 | tail -60 > gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
for P in 128 256; do
  for shape in "50176 1024" "200704 512" "12544 1536" "802816 256"; do
    SRGAN_ACT_L2_PROMOTION=$P timeout 300 python tools/trunk_gemm_bench.py $shape 10 2>&1 | sed "s/^/promo$P /" >> gpurun_out/${TAG}_trunk_gemm.txt
  done
done
cat gpurun_out/${TAG}_trunk_gemm.txt
for P in 128 256; do
  SRGAN_ACT_L2_PROMOTION=$P timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --secondary none > gpurun_out/${TAG}_bench_promo$P.log 2>&1
  tail -c 600 gpurun_out/${TAG}_bench_promo$P.log | grep -o '"ms_per_step": [0-9.]*' | head -1 | sed "s/^/promo$P /"
done
SRGAN_NO_DIRECT_CONCAT=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --secondary none > gpurun_out/${TAG}_bench_nodirect.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/${TAG}_bench_nodirect.log | head -1 | sed "s/^/nodirect /"
NCU="ncu --clock-control none"
timeout 1200 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${TAG}_crowd_launches.csv \
  python tools/crowd_step_profile.py 64 > gpurun_out/${TAG}_crowd_profile.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_crowd_launches.csv > gpurun_out/${TAG}_crowd_launches_summary.txt 2>&1
head -30 gpurun_out/${TAG}_crowd_launches_summary.txt
timeout 600 $NCU --set full --import-source on -k regex:umma_conv_persistent -s 4 -c 1 -f -o gpurun_out/${TAG}_trunk_dgrad \
  python tools/trunk_gemm_bench.py 50176 1024 2 > gpurun_out/${TAG}_ncu_trunk.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_trunk.log | cut -c1-200
ls -la gpurun_out/${TAG}_*
