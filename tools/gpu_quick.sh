#!/bin/bash
mkdir -p gpurun_out
echo "== pytest subset"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "generic or interleaved or i16_bit_exact or ragged or golden" > gpurun_out/pytest_quick.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_quick.log
echo "== g3"; timeout 600 python tools/bench_configs.py --config g3 > gpurun_out/cfg_g3.json 2> gpurun_out/cfg_g3.err; echo "rc=$?"; cat gpurun_out/cfg_g3.json; tail -3 gpurun_out/cfg_g3.err
