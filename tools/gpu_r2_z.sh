#!/bin/bash
# round 2, call Z: tail kernel forked onto a side stream; whole GPU suite, multichannel timings
O=gpurun_out/r2z; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 300 python tests/perf/bench_configs.py --config mci > $O/cfg_mci.json 2> $O/cfg_mci.err
timeout 300 python tests/perf/bench_configs.py --config gen > $O/cfg_gen.json 2> $O/cfg_gen.err
tail -5 $O/pytest.log
