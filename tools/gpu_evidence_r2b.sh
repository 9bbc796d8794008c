#!/bin/bash
# in-situ `--set full` captures of two consecutive bn_dgrad_kernel launches (conv2's <1> and conv1's <0> of one dense layer)
TAG=${1:-r2final}
NCU="ncu --clock-control none"
timeout 900 $NCU --profile-from-start off --set full --import-source on -k regex:bn_dgrad_kernel -s 300 -c 2 -f -o gpurun_out/${TAG}_bn_dgrad \
    python tools/crowd_step_profile.py 64 > gpurun_out/${TAG}_bn_dgrad.log 2>&1
tail -1 gpurun_out/${TAG}_bn_dgrad.log | cut -c1-200
python tools/summarize_ncu.py gpurun_out/${TAG}_bn_dgrad.ncu-rep > gpurun_out/${TAG}_ncu_full_summary_b.txt 2>&1
grep "==\|traffic" gpurun_out/${TAG}_ncu_full_summary_b.txt
