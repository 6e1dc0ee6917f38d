#!/bin/bash
# round 2, call S: far-lag int16 kernel for 12 / 16 channels
O=gpurun_out/r2s; mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_random.py -m gpu -x -q -k "far or few_channel_long or random or very_long" ) > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 300 python tests/perf/bench_configs.py --config mci > $O/cfg_mci.json 2> $O/cfg_mci.err
tail -4 $O/pytest.log
