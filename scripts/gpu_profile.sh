#!/bin/bash
# ncu launch list (per-launch device time of every bd:: kernel over one truncated bench pass) + one --set full capture
# of the dominant GEMM. Numbers printed by a run under ncu are never bench values.
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 0 --ar-steps 1 --graph 0 --no-cpu-baseline --no-roofline"
timeout 1200 ncu --kernel-name-base mangled -k regex:_ZN2bd --metrics gpu__time_duration.sum --clock-control none \
    -c ${NLAUNCH:-9000} --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches.csv)"
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base mangled \
    -k regex:${KREGEX:-bd_gemm_kernelILi128} -s ${KSKIP:-300} -c ${KCOUNT:-12} -f -o gpurun_out/prof_gemm $BENCH > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/*.ncu-rep 2>/dev/null
