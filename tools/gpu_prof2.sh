#!/bin/bash
mkdir -p gpurun_out
C1="python tests/perf/bench_configs.py --config i16"
$C1 > gpurun_out/plain_i16.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:stream_i16 -s 0 -c 1 -f -o gpurun_out/prof_i16_k3 $C1 > gpurun_out/ncu_i16a.log 2>&1; echo "rc=$?"
$C1 > gpurun_out/plain_i16b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:stream_i16 -s 20 -c 1 -f -o gpurun_out/prof_i16_k4096 $C1 > gpurun_out/ncu_i16b.log 2>&1; echo "rc=$?"
C2="python tests/perf/bench_configs.py --config 5i --log2 30"
$C2 > gpurun_out/plain_cols.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:stream_cols -s 1 -c 1 -f -o gpurun_out/prof_cols $C2 > gpurun_out/ncu_cols.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
