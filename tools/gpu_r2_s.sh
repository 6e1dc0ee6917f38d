#!/bin/bash
# round 2, call S: flat-stream int16 kernel, 128-thread CTAs two per SM
O=gpurun_out/r2s; mkdir -p $O
timeout 300 python tests/perf/bench_configs.py --config mci --tune threads=128 > $O/cfg_mci_128.json 2> $O/cfg_mci_128.err
timeout 300 python tests/perf/bench_configs.py --config mci > $O/cfg_mci.json 2> $O/cfg_mci.err
ls -la $O
