#!/bin/bash
# ncu evidence for the round (numbers printed by a run under ncu are never bench values). One truncated bench pass per
# capture: prefill + 1 AR step (+ decode for the conv capture), CUDA graph off so that every kernel is visible.
#   1. launch list: per-launch device time of every kernel (shares of the step);
#   2. --set full of the persistent kernel: first launch = one DiffHead.sample (51 evaluations), second = one Qwen3 AR block;
#   3. --set full of the tiled GEMM (prefill), the attention kernel (prefill) and the tokenizer convolution (decode).
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-r02}
export BD_LLM_STREAM=${BD_LLM_STREAM:-1}
BENCH="python bench.py --steps 1 --warmup 0 --ar-steps 1 --graph 0 --no-cpu-baseline --no-gpu-reference --no-roofline"
FULL="python bench.py --steps 1 --warmup 0 --ar-steps 64 --graph 1 --no-cpu-baseline --no-gpu-reference --no-roofline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NLAUNCH:-4000} --csv \
    --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/${TAG}_launches.csv)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:bd_stream_kernel -c 3 -f \
    -o gpurun_out/${TAG}_prof_stream $BENCH > gpurun_out/${TAG}_ncu_stream.log 2>&1
echo "stream capture rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bd_gemm_kernel -s 4 -c 3 -f \
    -o gpurun_out/${TAG}_prof_gemm $BENCH > gpurun_out/${TAG}_ncu_gemm.log 2>&1
echo "gemm capture rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bd_attn_kernel -s 2 -c 2 -f \
    -o gpurun_out/${TAG}_prof_attn $BENCH > gpurun_out/${TAG}_ncu_attn.log 2>&1
echo "attn capture rc=$?"
# the conv capture needs the decode: a small script instead of a whole image (encode + decode of one 1024^2 image)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bd_conv_kernel -s 20 -c 3 -f \
    -o gpurun_out/${TAG}_prof_conv python scripts/ae_bench.py --bs 1 --reps 1 > gpurun_out/${TAG}_ncu_conv.log 2>&1
echo "conv capture rc=$?"
# bs = 8 (M = 1024 rows, tensor-bound): launch list of prefill + one AR step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/${TAG}_bs8_launches.csv python bench.py --bs 8 --steps 1 --warmup 0 --ar-steps 1 --graph 0 \
    --no-cpu-baseline --no-gpu-reference --no-roofline > gpurun_out/${TAG}_ncu_bs8.log 2>&1
echo "bs8 launch list rc=$?"
# ImageNet class-conditional path (B-16x, bs 64): launch list of the first AR positions
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv \
    --log-file gpurun_out/${TAG}_imagenet_launches.csv python scripts/imagenet_bench.py --bs 64 --reps 1 --no-warmup \
    > gpurun_out/${TAG}_ncu_imagenet.log 2>&1
echo "imagenet launch list rc=$?"
ls -la gpurun_out/${TAG}_prof_*.ncu-rep 2>/dev/null
