#!/bin/bash
# round 2, call AD: prefix-sum chunk sizes with the vectorised kernels
O=gpurun_out/r2ad; mkdir -p $O
for kb in 16 32 64; do
  MAVG_SCAN_CHUNK_KB=$kb timeout 300 python tests/perf/bench_configs.py --config scan > $O/cfg_scan_$kb.json 2> $O/cfg_scan_$kb.err
done
cat $O/cfg_scan_*.json
