#!/bin/bash
# round 2, call M: L2 policies of the far-lag kernel
O=gpurun_out/r2m; mkdir -p $O
for h in 0 1 2; do
  for k in 60000 300000; do
    MAVG_FAR_HINTS=$h python tests/perf/run_shape.py f32 1 134217728 $k 5 >> $O/far_hints.log 2>&1
  done
done
MAVG_FAR_HINTS=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:stream_far -s 1 -c 1 --csv --log-file $O/far_h1_k60000.csv python tests/perf/run_shape.py f32 1 134217728 60000 3 > /dev/null 2>&1
MAVG_FAR_HINTS=2 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:stream_far -s 1 -c 1 --csv --log-file $O/far_h2_k60000.csv python tests/perf/run_shape.py f32 1 134217728 60000 3 > /dev/null 2>&1
MAVG_FAR_HINTS=0 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:stream_far -s 1 -c 1 --csv --log-file $O/far_h0_k60000.csv python tests/perf/run_shape.py f32 1 134217728 60000 3 > /dev/null 2>&1
MAVG_FAR_HINTS=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:stream_far -s 1 -c 1 --csv --log-file $O/far_h1_k300000.csv python tests/perf/run_shape.py f32 1 134217728 300000 3 > /dev/null 2>&1
MAVG_FAR_HINTS=2 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:stream_far -s 1 -c 1 --csv --log-file $O/far_h2_k300000.csv python tests/perf/run_shape.py f32 1 134217728 300000 3 > /dev/null 2>&1
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "far_lag" ) > $O/pytest_far.log 2>&1; echo "rc=$?" >> $O/pytest_far.log
ls $O
