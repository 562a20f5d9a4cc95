#!/bin/bash
# Final verification session of round 2 (one GPU): new interleaved tests first, the whole GPU suite as the driver runs it,
# the default bench line, the reference arm, the text-decode step. Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "=== interleaved tests"; date +%s
timeout 400 python -m pytest tests/test_interleaved_gpu.py -m gpu -q -s --no-header -p no:cacheprovider > gpurun_out/interleaved_gpu.log 2>&1
echo "rc=$?"; tail -n 25 gpurun_out/interleaved_gpu.log
echo "=== full GPU suite"; date +%s
timeout 700 python -m pytest tests -m gpu -q -s --no-header -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1
echo "rc=$?"; tail -n 12 gpurun_out/gpu_tests.log
echo "=== smoke"; date +%s
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "rc=$?"; tail -n 3 gpurun_out/smoke.log
echo "=== bench"; date +%s
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "rc=$?"; cut -c1-600 gpurun_out/bench_n1.json; tail -n 3 gpurun_out/bench_n1.err
echo "=== text decode"; date +%s
timeout 300 python scripts/text_decode_bench.py > gpurun_out/text_decode.json 2> gpurun_out/text_decode.err
echo "rc=$?"; cat gpurun_out/text_decode.json; tail -n 3 gpurun_out/text_decode.err
echo "=== reference arm"; date +%s
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "rc=$?"; cut -c1-1200 gpurun_out/bench_reference.json; tail -n 3 gpurun_out/bench_reference.err
date +%s
