#!/bin/bash
# One GPU session: parity tests, smoke, a truncated bench (fast signal), then the full default bench.
mkdir -p gpurun_out
TESTS="${TESTS:-tests/test_head_gpu.py tests/test_ae_gpu.py tests/test_pipeline_gpu.py}" bash scripts/gpu_check.sh -s
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "=== bench (truncated AR loop: 4 steps)"; timeout 900 python bench.py --steps 1 --warmup 1 --ar-steps 4 --no-cpu-baseline > gpurun_out/bench_trunc.json 2> gpurun_out/bench_trunc.err; tail -3 gpurun_out/bench_trunc.err; cat gpurun_out/bench_trunc.json
echo "=== bench (default)"; timeout 1500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
