#!/bin/bash
# round 2, call AB: ncu captures of the far-lag kernels (float32 k = 60 000, int16 stereo k = 40 000)
O=gpurun_out/r2ab; mkdir -p $O
prof() { # name regex args...
  name=$1; rx=$2; shift 2
  python tests/perf/run_shape.py "$@" > $O/$name.plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -f -o $O/prof_$name python tests/perf/run_shape.py "$@" > $O/$name.ncu.log 2>&1
  echo "$name rc=$?" >> $O/rc.log
}
prof far_f32_k60000 stream_far_f32 f32 1 134217728 60000 3
prof far_i16_c2_k40000 stream_far_i16 i16 2 67108864 40000 3
ls -la $O
