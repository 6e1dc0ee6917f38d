#!/bin/bash
# round 2, call I: ncu captures of the int16 column kernel (256 ch), the few-channel int16 pair kernel (6 ch) and the rewritten stereo kernel
O=gpurun_out/r2i; mkdir -p $O
prof() { # name regex args...
  name=$1; rx=$2; shift 2
  python tests/perf/run_shape.py "$@" > $O/$name.plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -f -o $O/prof_$name python tests/perf/run_shape.py "$@" > $O/$name.ncu.log 2>&1
  echo "$name rc=$?" >> $O/rc.log
}
prof cols_i16_c256_k64 stream_cols_i16x2 i16 256 524288 64 3
prof fewc_i16_c6_k64 stream_fewc_i16x2 i16 6 33554432 64 3
prof i16_stereo_k4096 stream_i16_kernel i16 2 134217728 4096 3
ls -la $O
