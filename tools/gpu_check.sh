#!/bin/bash
# First-contact GPU run: smoke, GPU parity tests, tuning sweep, bench.  Logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity.py::test_full_size_2p28_vs_oracle > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== sweep"; timeout 900 python tools/sweep.py --out gpurun_out/sweep.csv > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?"; tail -50 gpurun_out/sweep.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
