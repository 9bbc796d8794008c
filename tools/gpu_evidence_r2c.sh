#!/bin/bash
# Commands behind the closing evidence of round 2 (profiles/r2_final_*), each run under `gpurun -- '<line>'` on one B200 unless noted.
# Outputs land in gpurun_out/ and the text / JSON summaries are copied to profiles/ by hand (see profiles/README.md).
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1                                   # r2_final_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1                      # r2_final_smoke.log
python bench.py > gpurun_out/bench.log 2>&1                                                         # r2_final_bench_crowd.json
for b in 32 128; do python bench.py --batch $b --steps 20 --warmup 5 --secondary none --no-cpu-baseline --no-gpu-baseline \
    > gpurun_out/bench_b$b.log 2>&1; done                                                           # r2_final_bench_crowd_b{32,128}.json
python tools/crowd_data_bench.py > gpurun_out/crowd_data_bench.txt 2>&1                             # r2_final_crowd_data_bench.txt
python tools/sgan_bench.py > gpurun_out/sgan_bench.txt 2>&1                                         # r2_final_sgan_bench.txt
for g in 0 1; do SRGAN_COEF_GRAPH=$g python tools/coef_bench.py > gpurun_out/coef_graph$g.txt 2>&1; done   # r2_final_coef_graph.txt
ncu --set full --clock-control none -k regex:"extract_patches_kernel|knn_maps_kernel|density_label_kernel" -c 4 \
    -o gpurun_out/ncu_data python tools/crowd_data_bench.py --iters 2                              # r2_final_ncu_data_kernels_summary.txt
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"head_wgrad_kernel|head_logits_kernel|seed_rows_multi_kernel|logits_reduce" \
    -c 12 --csv --log-file gpurun_out/sgan_launches.csv python tools/sgan_bench.py 100             # r2_final_sgan_head_launches.csv
# multi-GPU (gpurun --gpus N): weak scaling at 64 per GPU, strong scaling at global batch 64
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus N \
#       --batch {64,32} --steps 20 --warmup 5 --secondary none                                     # r2_final_bench_crowd_n{2,8}*.json
