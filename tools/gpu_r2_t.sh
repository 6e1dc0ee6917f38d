#!/bin/bash
# round 2, call T: int16 column kernel, 8 warps x 32 frames against 16 x 16
O=gpurun_out/r2t; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_random.py -m gpu -x -q -k "column or cols or many_channel or random" ) > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 300 python tests/perf/bench_configs.py --config ci > $O/cfg_ci.json 2> $O/cfg_ci.err
timeout 300 python tests/perf/bench_configs.py --config ci --tune direct_max_k=1 > $O/cfg_ci_direct.json 2> $O/cfg_ci_direct.err
ls -la $O; tail -5 $O/pytest.log
