#!/bin/bash
# round 2: the two ncu passes of B200_PROFILING.md on a short bench command (each only after the same command exited 0 without ncu).
# The reports are exported to raw CSV on the box and deleted there (gpurun brings back at most 64 MiB).
O=gpurun_out/r2prof; mkdir -p $O
SK=parity,dense_k,i16,configs,shapes,e2e,steps
CMD="python bench.py --steps 2 --warmup 3 --skip $SK --no-cpu-baseline"
$CMD > $O/plain1.json 2> $O/plain1.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu_list.log 2>&1; echo "list rc=$?" >> $O/rc.log
CMD2="python bench.py --steps 2 --warmup 3 --ks 3,64,4096 --skip $SK --no-cpu-baseline --no-graph"
$CMD2 > $O/plain2.json 2> $O/plain2.err && \
ncu --set full --clock-control none -k regex:stream_f32 -s 9 -c 3 -f -o $O/prof_stream $CMD2 > $O/ncu_full.log 2>&1; echo "full rc=$?" >> $O/rc.log
ncu -i $O/prof_stream.ncu-rep --page raw --csv > $O/prof_stream.raw.csv 2>/dev/null; rm -f $O/prof_stream.ncu-rep
# the int16 block's kernel, same way
CMD3="python tests/perf/run_shape.py i16 2 134217728 64 3"
$CMD3 > $O/plain3.log 2>&1 && ncu --set full --clock-control none -k regex:stream_i16 -s 1 -c 1 -f -o $O/prof_i16_k64 $CMD3 > $O/ncu_i16.log 2>&1; echo "i16 rc=$?" >> $O/rc.log
ncu -i $O/prof_i16_k64.ncu-rep --page raw --csv > $O/prof_i16_k64.raw.csv 2>/dev/null; rm -f $O/prof_i16_k64.ncu-rep
ls -la $O
