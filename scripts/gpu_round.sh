#!/bin/bash
# One GPU session: parity tests, smoke, GEMM micro-benchmark, then the default bench (graph replay, eager fallback).
mkdir -p gpurun_out
TESTS="${TESTS:-$(ls tests/test_*gpu*.py | tr '\n' ' ')}" bash scripts/gpu_check.sh -s
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
if [ "${GEMM_BENCH:-1}" = "1" ]; then echo "=== gemm_bench"; timeout 900 python scripts/gemm_bench.py 2>&1 | tail -200 > gpurun_out/gemm_bench.txt; grep -c GB gpurun_out/gemm_bench.txt; fi
echo "=== bench (default)"
timeout 1500 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err
if [ $? -ne 0 ]; then tail -5 gpurun_out/bench.err; echo "=== bench (eager fallback)"; timeout 1500 python bench.py --graph 0 ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; fi
cat gpurun_out/bench.json
if [ "${EAGER_TOO:-0}" = "1" ]; then timeout 1500 python bench.py --graph 0 --no-cpu-baseline > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err; cat gpurun_out/bench_eager.json; fi
