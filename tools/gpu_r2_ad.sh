#!/bin/bash
# round 2, call AD: int16 prefix sum, two chunks per CTA, registers capped for five resident CTAs
O=gpurun_out/r2ad; mkdir -p $O
( timeout 300 python -m pytest tests/test_gpu_scan.py tests/test_gpu_parity.py -m gpu -x -q -k "scan or prefix" ) > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
for i in 1 2 3; do timeout 300 python tests/perf/bench_configs.py --config scan >> $O/scan.log 2>> $O/scan.err; done
tail -2 $O/pytest.log; cat $O/scan.log | cut -c100-420
