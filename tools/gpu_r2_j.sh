#!/bin/bash
# round 2, call J: prefix-difference path for far windows
O=gpurun_out/r2j; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "prefix_difference or very_long or far_lag" ) > $O/pytest_far.log 2>&1; echo "rc=$?" >> $O/pytest_far.log
timeout 300 python tests/perf/bench_configs.py --config gen > $O/cfg_gen.json 2> $O/cfg_gen.err
ls -la $O
