#!/bin/bash
# round 2, call AH: stereo int16, one CTA of 512 threads per SM against two of 256
O=gpurun_out/r2ah; mkdir -p $O
for t in 0 256 0 256; do
  timeout 300 python tests/perf/bench_configs.py --config i16 --tune threads=$t >> $O/i16.log 2>> $O/i16.err
done
cut -c1-100,100-900 $O/i16.log | grep -o '"[0-9]*": {"ms": [0-9.]*' | paste - - - - - -
