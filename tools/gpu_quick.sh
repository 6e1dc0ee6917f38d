#!/bin/bash
mkdir -p gpurun_out
echo "== pytest scan"; timeout 1200 python -m pytest tests/test_gpu_scan.py -m gpu -q -x > gpurun_out/pytest_scan.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_scan.log
