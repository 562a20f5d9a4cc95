#!/bin/bash
# ncu evidence for the round (numbers printed by a run under ncu are never bench values):
#  1. launch list: per-launch device time of every kernel of one truncated bench pass (prefill + 1 AR step);
#  2. --set full capture of ONE launch of the dominant kernel, bd_stream_kernel (= one DiffHead.sample: 51 evaluations).
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 0 --ar-steps 1 --graph 0 --no-cpu-baseline --no-roofline"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NLAUNCH:-6000} --csv \
    --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches.csv)"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:bd_stream_kernel -s ${KSKIP:-1} -c 1 -f \
    -o gpurun_out/prof_stream $BENCH > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep 2>/dev/null
