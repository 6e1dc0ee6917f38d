#!/bin/bash
# round 2, call S: experiment -- the dense 72-sample shape on 4 / 8 channels
O=gpurun_out/r2s; mkdir -p $O
timeout 300 python tests/perf/bench_configs.py --config mci --tune run=72 > $O/cfg_mci_r72.json 2> $O/cfg_mci_r72.err
timeout 300 python tests/perf/bench_configs.py --config mci --tune run=72,threads=224 > $O/cfg_mci_r72_224.json 2> $O/cfg_mci_r72_224.err
ls -la $O
