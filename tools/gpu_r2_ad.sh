#!/bin/bash
# round 2, call AD: prefix sum with two chunks per CTA against one
O=gpurun_out/r2ad; mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_scan.py tests/test_gpu_parity.py -m gpu -x -q -k "scan or prefix" ) > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
for nch in 2 1 2 1; do
  MAVG_SCAN_NCH=$nch timeout 300 python tests/perf/bench_configs.py --config scan >> $O/cfg_scan_nch$nch.json 2> $O/cfg_scan_nch$nch.err
done
tail -3 $O/pytest.log; cat $O/cfg_scan_nch2.json; cat $O/cfg_scan_nch1.json
