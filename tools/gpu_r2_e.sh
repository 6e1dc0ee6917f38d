#!/bin/bash
# round 2, call E: RMS + far-lag staged context parity, ncu capture of the prefix-sum primitive
O=gpurun_out/r2e; mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_rms.py tests/test_gpu_parity.py -m gpu -x -q -k "rms or far_lag" ) > $O/pytest_rms_far.log 2>&1; echo "rc=$?" >> $O/pytest_rms_far.log
CMD="python tests/perf/scan_once.py 3"
$CMD > $O/scan_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scan_lookback -s 2 -c 1 -f -o $O/prof_scan $CMD > $O/ncu_scan.log 2>&1; echo "ncu rc=$?" >> $O/ncu_scan.log
ls -la $O
