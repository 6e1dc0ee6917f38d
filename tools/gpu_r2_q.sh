#!/bin/bash
# round 2, call Q: flat-stream int16 kernel generalised to 3 / 4 / 6 / 8 channels
O=gpurun_out/r2q; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py tests/test_gpu_random.py -m gpu -x -q -k "i16 or few_channel or flat_multichannel or random or multi_device" ) > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 300 python tests/perf/bench_configs.py --config mci > $O/cfg_mci.json 2> $O/cfg_mci.err
ls -la $O; tail -5 $O/pytest.log
