#!/bin/bash
# round 2, call K: whole GPU suite + bench + generic-shape timings after the RMS template change
O=gpurun_out/r2k2; mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 300 python tests/perf/bench_configs.py --config gen > $O/cfg_gen.json 2> $O/cfg_gen.err
timeout 300 python tests/perf/bench_configs.py --config g6i > $O/cfg_g6i.json 2> $O/cfg_g6i.err
timeout 300 python tests/perf/bench_configs.py --config g3 > $O/cfg_g3.json 2> $O/cfg_g3.err
( time timeout 600 python bench.py --steps 20 --warmup 5 ) > $O/bench_full.json 2> $O/bench_full.err
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 ) > $O/bench_ref.json 2> $O/bench_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
ls -la $O
