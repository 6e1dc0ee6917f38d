set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "i16 or golden" 2>&1 | tail -8
python tests/perf/bench_configs.py --config i16 2>&1 | tail -3
python tests/perf/bench_configs.py --config g6i 2>&1 | tail -3
