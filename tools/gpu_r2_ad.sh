#!/bin/bash
# round 2, call AD: float32 prefix sum, chunk size x chunks per CTA
O=gpurun_out/r2ad; mkdir -p $O
for cfg in "32 2" "64 2" "64 1" "32 2" "64 1"; do
  set -- $cfg
  echo "chunk_kb=$1 nch=$2" >> $O/f32.log
  MAVG_SCAN_CHUNK_KB=$1 MAVG_SCAN_NCH=$2 timeout 300 python tests/perf/bench_configs.py --config scan >> $O/f32.log 2>> $O/f32.err
done
cat $O/f32.log | cut -c1-400
