#!/bin/bash
# round 2, call R: ncu captures of the flat-stream int16 kernel at 6 and 8 channels
O=gpurun_out/r2r; mkdir -p $O
prof() { # name regex args...
  name=$1; rx=$2; shift 2
  python tests/perf/run_shape.py "$@" > $O/$name.plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -f -o $O/prof_$name python tests/perf/run_shape.py "$@" > $O/$name.ncu.log 2>&1
  echo "$name rc=$?" >> $O/rc.log
}
prof i16_c6_k64 stream_i16_kernel i16 6 22369621 64 3
prof i16_c8_k64 stream_i16_kernel i16 8 16777216 64 3
prof i16_c3_k64 stream_i16_kernel i16 3 44739242 64 3
ls -la $O
