#!/bin/bash
# round 2, call AG: ncu capture of the two-chunk int16 prefix-sum kernel
O=gpurun_out/r2ag; mkdir -p $O
CMD4="python tests/perf/scan_once.py 3"
$CMD4 > $O/plain4.log 2>&1 && ncu --set full --clock-control none -k regex:scan_lookback_fast -s 2 -c 1 -f -o $O/prof_scan $CMD4 > $O/ncu_scan.log 2>&1; echo "scan rc=$?" >> $O/rc.log
ncu -i $O/prof_scan.ncu-rep --page raw --csv > $O/prof_scan.raw.csv 2>/dev/null; rm -f $O/prof_scan.ncu-rep
ls -la $O; cat $O/rc.log
