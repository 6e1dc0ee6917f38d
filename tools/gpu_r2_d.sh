#!/bin/bash
# round 2, call D: moving RMS parity, prefix-sum with look-ahead, whole suite
O=gpurun_out/r2d; mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_rms.py tests/test_gpu_scan.py -m gpu -x -q ) > $O/pytest_rms_scan.log 2>&1; echo "rc=$?" >> $O/pytest_rms_scan.log
timeout 300 python tests/perf/bench_configs.py --config scan > $O/cfg_scan.json 2> $O/cfg_scan.err
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
ls -la $O
