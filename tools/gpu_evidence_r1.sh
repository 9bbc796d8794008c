#!/bin/bash
# Round-1 closing evidence on one GPU box: full -m gpu suite, smoke, benches (age fp32/bf16, crowd), ncu launch list of the
# eager age step, one `--set full` capture each of the dominant tcgen05 conv kernel and of the BN-affine streaming kernels.
# Only text summaries travel back (the .ncu-rep files stay on the box: gpurun_out/ is capped at 64 MiB).
# usage: tools/gpu_evidence_r1.sh tag
TAG=${1:-ev}
O=gpurun_out
mkdir -p $O /tmp/ncu
bash tools/gpu_round.sh $TAG
timeout 300 python bench.py --workload crowd --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_crowd.log 2>&1
tail -1 $O/${TAG}_bench_crowd.log | cut -c1-200
NCU="ncu --clock-control none"
SRGAN_NO_GRAPH=1 SRGAN_NO_OVERLAP=1 SRGAN_NO_WGRAD_STREAM=1 timeout 300 $NCU --metrics gpu__time_duration.sum -s 440 -c 300 --csv \
  --log-file $O/${TAG}_launches_age.csv python bench.py --precision bf16 --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_bench.log 2>&1
python tools/summarize_launches.py $O/${TAG}_launches_age.csv > $O/${TAG}_launches_age_summary.txt 2>&1
full() {  # name kernel-regex skip cmd...
  local name=$1 k=$2 skip=$3; shift 3
  timeout 200 $NCU --set full --import-source on -k regex:$k -s $skip -c 1 -f -o /tmp/ncu/${name} "$@" > $O/${TAG}_ncu_${name}.log 2>&1
  python tools/summarize_ncu.py /tmp/ncu/${name}.ncu-rep >> $O/${TAG}_ncu_full_summary.txt 2>&1
}
rm -f $O/${TAG}_ncu_full_summary.txt
full conv_l2_fprop 'umma_conv_persistent' 3 python tools/conv_bench.py D.l2 down --iters 2
full affine_fwd 'affine2d_kernel' 2 python tools/stream_bench.py 50176 1024 1792 2
full affine_bwd_grad 'affine_grad2d_kernel' 2 python tools/stream_bench.py 50176 1024 1792 2
rm -f $O/${TAG}_launches_age.csv.tmp
ls -la $O/${TAG}_* | head -30
