#!/bin/bash
# ncu --set full captures of the reworked int16 kernel and of the few-channel kernels (each after a plain run of the
# same command exited 0).  Launch indices: bench_configs times 4 launches per k (1 warm-up + 3).
mkdir -p gpurun_out
cap() {  # name, kernel regex, skip, command...
    local name=$1 re=$2 skip=$3; shift 3
    "$@" > gpurun_out/plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$re -s $skip -c 1 -f \
        -o gpurun_out/prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1
    echo "$name rc=$?"
}
cap i16_k3_v3 stream_i16 0 python tests/perf/bench_configs.py --config i16
cap i16_k4096_v3 stream_i16 20 python tests/perf/bench_configs.py --config i16
cap fewc_f32_k64 stream_fewc_f32 8 python tests/perf/bench_configs.py --config g3
cap fewc_f32_k1024 stream_fewc_f32 20 python tests/perf/bench_configs.py --config g3
cap fewc_i16x2_k64 stream_fewc_i16x2 8 python tests/perf/bench_configs.py --config g6i
ls -la gpurun_out/*.ncu-rep
