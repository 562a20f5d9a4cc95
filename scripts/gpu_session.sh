#!/bin/bash
# One GPU lease: tests, timelines and micro-benchmarks, each under its own timeout; logs land in gpurun_out/.
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${TAG:-r02}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt
for step in "$@"; do
  case "$step" in
    stream)   timeout 600 python -m pytest tests/test_stream_gpu.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/${TAG}_test_stream.log ;;
    head)     timeout 900 python -m pytest tests/test_head_gpu.py -m gpu -q -s 2>&1 | grep -v "^    \|^$" | tail -150 > gpurun_out/${TAG}_test_head.log ;;
    head0)    BD_HEAD_FILLERS=0 timeout 900 python -m pytest tests/test_head_gpu.py -m gpu -q -s -k "stream" 2>&1 | grep -v "^    \|^$" | tail -150 > gpurun_out/${TAG}_test_head_fill0.log ;;
    tests)    timeout 2400 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^    " > gpurun_out/${TAG}_tests.log ;;
    llmstream) BD_LLM_STREAM=1 timeout 900 python -m pytest tests/test_llm_gpu.py tests/test_pipeline_gpu.py -m gpu -q -s 2>&1 | grep -v "^    " | tail -80 > gpurun_out/${TAG}_test_llm_stream.log ;;
    benchllm) BD_LLM_STREAM=1 timeout 1500 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-gpu-reference > gpurun_out/${TAG}_bench_llmstream.json 2> gpurun_out/${TAG}_bench_llmstream.err ;;
    bench8)   timeout 1500 python bench.py --bs 8 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-reference > gpurun_out/${TAG}_bench_bs8.json 2> gpurun_out/${TAG}_bench_bs8.err ;;
    benchquick) timeout 1500 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-gpu-reference > gpurun_out/${TAG}_bench_quick.json 2> gpurun_out/${TAG}_bench_quick.err ;;
    gemmtest) timeout 900 python -m pytest tests/test_gemm_gpu.py -m gpu -q -s 2>&1 | grep -v "^    " | tail -40 > gpurun_out/${TAG}_test_gemm.log ;;
    llmtl)    timeout 900 python scripts/llm_timeline.py > gpurun_out/${TAG}_llm_timeline.txt 2>&1 ;;
    bench8g0) BD_GEMM2=0 timeout 1500 python bench.py --bs 8 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-reference > gpurun_out/${TAG}_bench_bs8_gemm1cta.json 2> gpurun_out/${TAG}_bench_bs8_gemm1cta.err ;;
    timeline) timeout 900 python scripts/head_timeline.py > gpurun_out/${TAG}_head_timeline.txt 2>&1 ;;
    ab)       timeout 900 python scripts/head_ab.py >> gpurun_out/${TAG}_head_ab.txt 2>&1 ;;
    ab_r01)   BD_LIB_PATH=$PWD/ab/libbitdance_b200_r01.so timeout 900 python scripts/head_ab.py >> gpurun_out/${TAG}_head_ab_r01.txt 2>&1 ;;
    diag)     timeout 600 python scripts/head_diag.py > gpurun_out/${TAG}_head_diag.txt 2>&1 ;;
    imagenet) timeout 1200 python scripts/imagenet_bench.py --bs 64 --modes 0 256 512 768 > gpurun_out/${TAG}_imagenet_bench.txt 2>&1
              timeout 600 python scripts/imagenet_bench.py --bs 64 256 --head-engines 0 >> gpurun_out/${TAG}_imagenet_bench.txt 2>&1 ;;
    imagenettest) timeout 900 python -m pytest tests/test_imagenet_gpu.py tests/test_head_gpu.py -m gpu -q -s 2>&1 | grep -v "^    " | tail -60 > gpurun_out/${TAG}_test_imagenet.log ;;
    ae)       timeout 900 python scripts/ae_bench.py > gpurun_out/${TAG}_ae_bench.txt 2>&1 ;;
    bench)    timeout 1500 python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err ;;
    benchref) timeout 1500 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err ;;
    profile)  TAG=${TAG} bash scripts/gpu_profile.sh > gpurun_out/${TAG}_profile_steps.log 2>&1 ;;
    bench2)   timeout 1500 python bench.py --steps 2 --warmup 3 --llm-stream 0 > gpurun_out/${TAG}_bench_chained.json 2> gpurun_out/${TAG}_bench_chained.err ;;
    smoke)    timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1 ;;
    *) echo "unknown step $step" ;;
  esac
  echo "$step rc=$?" >> gpurun_out/${TAG}_steps.log
done
tail -5 gpurun_out/${TAG}_*.log 2>/dev/null | tail -60
