#!/bin/bash
# round 2, call P: whole GPU suite + bench (both arms) + smoke + scan/gen shape timings at the re-entry head
O=gpurun_out/r2p; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
( time timeout 600 python bench.py --steps 20 --warmup 5 ) > $O/bench_full.json 2> $O/bench_full.err
timeout 300 python tests/perf/bench_configs.py --config scan > $O/cfg_scan.json 2> $O/cfg_scan.err
timeout 300 python tests/perf/bench_configs.py --config gen > $O/cfg_gen.json 2> $O/cfg_gen.err
timeout 300 python tests/perf/bench_configs.py --config g6i > $O/cfg_g6i.json 2> $O/cfg_g6i.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
ls -la $O
