#!/bin/bash
# round 2, call L: vectorised int16 prefix-sum kernel
O=gpurun_out/r2l2; mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_scan.py tests/test_gpu_parity.py -m gpu -x -q -k "prefix" ) > $O/pytest_scan.log 2>&1; echo "rc=$?" >> $O/pytest_scan.log
timeout 300 python tests/perf/bench_configs.py --config scan > $O/cfg_scan.json 2> $O/cfg_scan.err
MAVG_SCAN_GENERAL=1 timeout 300 python tests/perf/bench_configs.py --config scan > $O/cfg_scan_general.json 2> $O/cfg_scan_general.err
CMD4="python tests/perf/scan_once.py 3"
$CMD4 > $O/plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scan_lookback_fast -s 2 -c 1 -f -o $O/prof_scan_fast $CMD4 > $O/ncu_scan.log 2>&1; echo "scan rc=$?" >> $O/rc.log
ls -la $O
