#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/sanitize_small.py > gpurun_out/sanitize_plain.log 2>&1 && timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -8 gpurun_out/sanitize_memcheck.log
