#!/bin/bash
# round 2, call V: far-lag kernel with in-place results, 384 x 16 tiles, deeper rings
O=gpurun_out/r2v; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_random.py tests/test_gpu_multi.py -m gpu -x -q -k "far or very_long or random" ) > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 300 python tests/perf/bench_configs.py --config gen > $O/cfg_gen.json 2> $O/cfg_gen.err
timeout 300 python tests/perf/bench_configs.py --config gen --tune threads=512 > $O/cfg_gen_512.json 2> $O/cfg_gen_512.err
timeout 300 python tests/perf/bench_configs.py --config gen --tune prefetch=3 > $O/cfg_gen_p3.json 2> $O/cfg_gen_p3.err
ls -la $O; tail -5 $O/pytest.log
