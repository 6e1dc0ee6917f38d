#!/bin/bash
# round 2, call AA: far-lag int16 kernel
O=gpurun_out/r2aa; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_random.py -m gpu -x -q -k "far or very_long or prefix_difference or random" ) > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 300 python tests/perf/bench_configs.py --config gen > $O/cfg_gen.json 2> $O/cfg_gen.err
tail -5 $O/pytest.log
