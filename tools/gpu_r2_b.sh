#!/bin/bash
# round 2, call B: the rewritten int16 streaming kernel -- parity first, then timing, then the whole suite
O=gpurun_out/r2b; mkdir -p $O
( time timeout 600 python -m pytest tests -m gpu -x -q -k "i16 or int16 or dropin or random" ) > $O/pytest_i16.log 2>&1; echo "rc=$?" >> $O/pytest_i16.log
timeout 300 python tests/perf/bench_configs.py --config i16 > $O/cfg_i16.json 2> $O/cfg_i16.err
timeout 300 python tests/perf/bench_configs.py --config i16 --tune threads=256 > $O/cfg_i16_t256.json 2> $O/cfg_i16_t256.err
timeout 300 python tests/perf/bench_configs.py --config i16 --tune prefetch=1 > $O/cfg_i16_p1.json 2> $O/cfg_i16_p1.err
SK=parity,dense_k,i16,configs,e2e,steps
for t in threads=256 threads=256,run=32 chunks_per_cta=2 prefetch=1; do
  timeout 300 python bench.py --steps 20 --warmup 5 --tune $t --skip $SK --no-cpu-baseline > $O/bench_$t.json 2> $O/bench_$t.err
done
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
( time timeout 600 python bench.py --steps 20 --warmup 5 ) > $O/bench_full.json 2> $O/bench_full.err
ls -la $O
