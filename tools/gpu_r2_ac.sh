#!/bin/bash
# round 2, call AC: mavg_run_host_sweep
O=gpurun_out/r2ac; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_dropin.py -m gpu -x -q -k "sweep" ) > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -30 $O/pytest.log
