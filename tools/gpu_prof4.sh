#!/bin/bash
# ncu --set full captures of the far-lag kernel (k = 60 000 and k = 300 000 on 2^27 mono float32 samples), each after a
# plain run of the same command exited 0.
mkdir -p gpurun_out
C="python tests/perf/bench_configs.py --config gen"
$C > gpurun_out/plain_far.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:stream_far -s 0 -c 1 -f -o gpurun_out/prof_far_k60000 $C > gpurun_out/ncu_far_a.log 2>&1; echo "rc=$?"
$C > gpurun_out/plain_far2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:stream_far -s 4 -c 1 -f -o gpurun_out/prof_far_k300000 $C > gpurun_out/ncu_far_b.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
