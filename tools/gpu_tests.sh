#!/bin/bash
# full GPU suite + bench (no ncu)
mkdir -p gpurun_out
echo "== pytest (all gpu tests)"; timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-700 gpurun_out/bench.json; python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print({k:v['ms'] for k,v in d['per_k'].items()}, d['roofline']['frac'], d['e2e']['value'])"; tail -3 gpurun_out/bench.err
