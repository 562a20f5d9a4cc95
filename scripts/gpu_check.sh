#!/bin/bash
# Run each GPU test module in its own process (a CUDA fault in one does not hide the others); logs in gpurun_out/.
# usage: scripts/gpu_check.sh [pytest -k expr] ; env TESTS="tests/test_a.py tests/test_b.py" to select modules.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
TESTS=${TESTS:-$(ls tests/test_*gpu*.py)}
rc=0
for t in $TESTS; do
  name=$(basename "$t" .py)
  echo "=== $t"
  timeout ${TEST_TIMEOUT:-600} python -m pytest "$t" -m gpu -x -q --no-header -p no:cacheprovider "$@" > "gpurun_out/$name.log" 2>&1
  r=$?
  tail -n 15 "gpurun_out/$name.log"
  [ $r -ne 0 ] && rc=$r
done
exit $rc
