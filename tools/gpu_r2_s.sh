#!/bin/bash
# round 2, call S: two-CTA shapes for 5 / 7 channels
O=gpurun_out/r2s; mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_random.py -m gpu -x -q -k "few_channel or flat_multichannel or random" ) > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 300 python tests/perf/bench_configs.py --config mci > $O/cfg_mci.json 2> $O/cfg_mci.err
tail -4 $O/pytest.log
