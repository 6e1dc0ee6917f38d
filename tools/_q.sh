python tests/perf/bench_configs.py --config g6i 2>&1 | tail -1
python tests/perf/bench_configs.py --config g6i --tune direct_max_k=16 2>&1 | tail -1
python tests/perf/bench_configs.py --config g3 2>&1 | tail -1
python tests/perf/bench_configs.py --config g3 --tune direct_max_k=16 2>&1 | tail -1
