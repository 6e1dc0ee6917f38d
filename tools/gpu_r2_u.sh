#!/bin/bash
# round 2, call U: ncu captures of the int16 column kernel with TMA stores
O=gpurun_out/r2u; mkdir -p $O
prof() { # name regex args...
  name=$1; rx=$2; shift 2
  python tests/perf/run_shape.py "$@" > $O/$name.plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -f -o $O/prof_$name python tests/perf/run_shape.py "$@" > $O/$name.ncu.log 2>&1
  echo "$name rc=$?" >> $O/rc.log
}
prof cols_i16_c256_k64 stream_cols_i16x2 i16 256 524288 64 3
prof cols_i16_c64_k64 stream_cols_i16x2 i16 64 2097152 64 3
ls -la $O
