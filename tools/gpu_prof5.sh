#!/bin/bash
# ncu --set full capture of the int16 column kernel (64 channels, k = 64, 2^27 samples), after a plain run exited 0.
mkdir -p gpurun_out
C="python tests/perf/bench_configs.py --config gen"
$C > gpurun_out/plain_cols16.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:stream_cols_i16x2 -s 0 -c 1 -f -o gpurun_out/prof_cols_i16x2 $C > gpurun_out/ncu_cols16.log 2>&1; echo "rc=$?"
