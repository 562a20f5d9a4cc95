#!/bin/bash
# ncu evidence for the round (numbers printed by a run under ncu are never bench values). One truncated bench pass per
# capture: prefill + 1 AR step (+ decode for the conv capture), CUDA graph off so that every kernel is visible.
#   1. launch lists: per-launch device time of every kernel (shares of the step) for bs=1, bs=8 and the ImageNet path;
#   2. --set full of the persistent kernel: its head instance (one DiffHead.sample, 51 evaluations) and its Qwen3 instance
#      (one AR block, 40 layers), with the SASS-level stall sampling;
#   3. --set full of the tiled GEMM (prefill), the attention kernel (prefill) and the tokenizer convolution (decode).
# gpurun brings back at most 64 MiB: the .ncu-rep files are summarised HERE (raw page -> scripts/ncu_summary.py's input,
# source page -> scripts/ncu_source_top.py) and deleted; only text comes home.
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-r02}
O=gpurun_out/${TAG}
BENCH="python bench.py --steps 1 --warmup 0 --ar-steps 1 --graph 0 --no-cpu-baseline --no-gpu-reference --no-roofline"

timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NLAUNCH:-4000} --csv \
    --log-file ${O}_launches.csv $BENCH > ${O}_ncu_list.log 2>&1
echo "launch list rc=$? lines=$(wc -l < ${O}_launches.csv)"
python scripts/launch_summary.py ${O}_launches.csv > ${O}_launch_list_summary.txt 2>&1

capture() {  # name, kernel regex, skip, count, source?, command...
  local name=$1 rx=$2 skip=$3 cnt=$4 src=$5; shift 5
  local extra=""
  [ "$src" = "1" ] && extra="--import-source on"
  timeout 1200 ncu --set full --clock-control none $extra -k regex:$rx -s $skip -c $cnt -f -o ${O}_prof_$name "$@" \
      > ${O}_ncu_$name.log 2>&1
  echo "$name capture rc=$? $(ls -la ${O}_prof_$name.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
  if [ -f ${O}_prof_$name.ncu-rep ]; then
    ncu -i ${O}_prof_$name.ncu-rep --page raw --csv > ${O}_prof_${name}_raw.csv 2>/dev/null
    if [ "$src" = "1" ]; then
      ncu -i ${O}_prof_$name.ncu-rep --page source --csv --print-source sass 2>/dev/null | \
          python scripts/ncu_source_top.py 120 > ${O}_prof_${name}_stalls.txt 2>&1
    fi
    rm -f ${O}_prof_$name.ncu-rep
  fi
  tail -3 ${O}_ncu_$name.log | cut -c1-200
}

capture stream bd_stream_kernel 0 2 1 $BENCH
capture gemm bd_gemm_kernel 4 2 0 $BENCH
capture attn bd_attn_kernel 2 1 0 $BENCH
capture conv bd_conv_kernel 20 2 0 python scripts/ae_bench.py --bs 1 --reps 1
# the CTA-pair GEMM at M = 1024 rows (bs = 8): three launches of the head's Linears inside the first evaluations
capture gemm2 bd_gemm2_kernel 60 3 0 python bench.py --bs 8 --steps 1 --warmup 0 --ar-steps 1 --graph 0 \
    --no-cpu-baseline --no-gpu-reference --no-roofline

# bs = 8 (M = 1024 rows, tensor-bound): launch list of prefill + one AR step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file ${O}_bs8_launches.csv python bench.py --bs 8 --steps 1 --warmup 0 --ar-steps 1 --graph 0 \
    --no-cpu-baseline --no-gpu-reference --no-roofline > ${O}_ncu_bs8.log 2>&1
echo "bs8 launch list rc=$?"
python scripts/launch_summary.py ${O}_bs8_launches.csv > ${O}_bs8_launch_list_summary.txt 2>&1
# ImageNet class-conditional path (B-16x, bs 64, multi-kernel head): launch list of the first AR positions
BD_IMAGENET_HEAD_ENGINES=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file ${O}_imagenet_launches.csv python scripts/imagenet_bench.py --bs 64 --reps 1 --no-warmup \
    > ${O}_ncu_imagenet.log 2>&1
echo "imagenet launch list rc=$?"
python scripts/launch_summary.py ${O}_imagenet_launches.csv > ${O}_imagenet_launch_list_summary.txt 2>&1
# keep the text, drop the bulky raw launch lists except the bs=1 one
rm -f ${O}_bs8_launches.csv ${O}_imagenet_launches.csv
du -sh gpurun_out
