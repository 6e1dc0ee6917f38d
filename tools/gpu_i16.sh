#!/bin/bash
mkdir -p gpurun_out
echo "== pytest i16"; timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -q -x -k "i16 or golden or int16 or multi_device or pipeline" > gpurun_out/pytest_i16.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_i16.log
echo "== i16 bench"; timeout 600 python tools/bench_configs.py --config i16 > gpurun_out/cfg_i16_n1.json 2> gpurun_out/cfg_i16_n1.err; echo "rc=$?"; cat gpurun_out/cfg_i16_n1.json; tail -3 gpurun_out/cfg_i16_n1.err
