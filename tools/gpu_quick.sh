#!/bin/bash
mkdir -p gpurun_out
echo "== pytest subset"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "many_channel" > gpurun_out/pytest_quick.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_quick.log
echo "== cfg 5i"; timeout 600 python tools/bench_configs.py --config 5i > gpurun_out/cfg5i_n1.json 2> gpurun_out/cfg5i_n1.err; echo "rc=$?"; cat gpurun_out/cfg5i_n1.json; tail -3 gpurun_out/cfg5i_n1.err
for t in "prefetch=3" "prefetch=4"; do timeout 600 python tools/bench_configs.py --config 5i --tune "$t" 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['tune'], round(d['ms'],4), round(d['hbm_gbs_per_gpu'],1))"; done
