#!/bin/bash
# round 2, call X: full-size property tests (shift invariance, impulse response)
O=gpurun_out/r2x; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shift_invariance" ) > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -30 $O/pytest.log
