#!/bin/bash
# round 2, call A: GPU test suite, PDL A/B on the headline, the full bench line, baselines of the kernels to be reworked
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
SK=parity,dense_k,i16,configs,e2e,steps
for ov in 3 1 2 3 1 2; do
  timeout 300 python bench.py --steps 20 --warmup 5 --overlap $ov --skip $SK --no-cpu-baseline >> $O/bench_ov$ov.json 2>> $O/bench_ov$ov.err
done
timeout 300 python bench.py --steps 20 --warmup 5 --overlap 2 --no-graph --skip $SK --no-cpu-baseline > $O/bench_ov2_nograph.json 2> $O/bench_ov2_nograph.err
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $O/bench_full.json 2> $O/bench_full.err
for c in scan gen i16 g6i; do timeout 300 python tests/perf/bench_configs.py --config $c > $O/cfg_$c.json 2> $O/cfg_$c.err; done
ls -la $O
