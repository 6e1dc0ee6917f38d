#!/bin/bash
# round 2, call C: rewritten prefix-sum primitive (16-byte descriptors, early aggregate, 1..8 channels) + write-bandwidth probe
O=gpurun_out/r2c; mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_scan.py -m gpu -x -q ) > $O/pytest_scan.log 2>&1; echo "rc=$?" >> $O/pytest_scan.log
timeout 300 python tests/perf/bench_configs.py --config scan > $O/cfg_scan.json 2> $O/cfg_scan.err
timeout 300 python tests/perf/write_bw_probe.py > $O/write_bw.json 2> $O/write_bw.err
ls -la $O
