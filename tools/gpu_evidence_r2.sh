#!/bin/bash
# round-2 evidence for profiles/: (1) the driver's bench line (crowd + age secondary + baselines), (2) ncu launch list of one
# eager crowd step, (3) `--set full` in-situ captures of the fused dense-layer kernels from that step, (4) per-op timing
TAG=${1:-r2final}
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.log 2>&1
tail -c 600 gpurun_out/${TAG}_bench.log
timeout 900 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${TAG}_crowd_launches.csv \
  python tools/crowd_step_profile.py 64 > gpurun_out/${TAG}_crowd_profile.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_crowd_launches.csv > gpurun_out/${TAG}_crowd_launches_summary.txt 2>&1
head -30 gpurun_out/${TAG}_crowd_launches_summary.txt
full() {  # name kernel-regex skip
  timeout 600 $NCU --profile-from-start off --set full --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/${TAG}_$1 \
    python tools/crowd_step_profile.py 64 > gpurun_out/${TAG}_$1.log 2>&1
  tail -1 gpurun_out/${TAG}_$1.log | cut -c1-200
}
full bn_dgrad 'bn_dgrad_kernel<\(bool\)0>' 150
full bn_conv_dgrad 'bn_dgrad_kernel<\(bool\)1>' 150
full bn_conv_down 'bn_conv_down_kernel' 100
full flat3x3 'flat3x3_kernel' 100
python tools/summarize_ncu.py gpurun_out/${TAG}_bn_dgrad.ncu-rep gpurun_out/${TAG}_bn_conv_dgrad.ncu-rep gpurun_out/${TAG}_bn_conv_down.ncu-rep gpurun_out/${TAG}_flat3x3.ncu-rep > gpurun_out/${TAG}_ncu_full_summary.txt 2>&1
grep -c "==" gpurun_out/${TAG}_ncu_full_summary.txt
timeout 600 python tools/op_time.py crowd 64 > gpurun_out/${TAG}_crowd_optime_b64.txt 2>&1; head -12 gpurun_out/${TAG}_crowd_optime_b64.txt
