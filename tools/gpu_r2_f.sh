#!/bin/bash
# round 2, call F: prefix-sum chunk-size sweep with the conflict-free swizzle
O=gpurun_out/r2f2; mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_scan.py -m gpu -x -q ) > $O/pytest_scan.log 2>&1; echo "rc=$?" >> $O/pytest_scan.log
for kb in 32 16; do
  MAVG_SCAN_CHUNK_KB=$kb timeout 300 python tests/perf/bench_configs.py --config scan > $O/cfg_scan_$kb.json 2> $O/cfg_scan_$kb.err
  MAVG_SCAN_CHUNK_KB=$kb timeout 300 python -m pytest tests/test_gpu_scan.py -m gpu -x -q -k "ch1 or 1-" > $O/pytest_scan_$kb.log 2>&1
done
ls -la $O
