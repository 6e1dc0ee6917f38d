#!/bin/bash
# round 2, call H: int16 column kernel on the delta-scan arithmetic -- parity, then timing
O=gpurun_out/r2h; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q -k "i16 or int16 or cols or column or random or dropin" ) > $O/pytest_i16.log 2>&1; echo "rc=$?" >> $O/pytest_i16.log
timeout 300 python tests/perf/bench_configs.py --config gen > $O/cfg_gen.json 2> $O/cfg_gen.err
ls -la $O
