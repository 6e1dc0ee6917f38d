#!/bin/bash
mkdir -p gpurun_out
echo "== pytest parity"; timeout 1800 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --deselect tests/test_gpu_parity.py::test_full_size_2p28_vs_oracle > gpurun_out/pytest_quick.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_quick.log
echo "== i16 bench"; timeout 600 python tools/bench_configs.py --config i16 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k:(v['ms'],v['hbm_gbs'],v['bit_exact_head']) for k,v in d['per_k'].items()})"
echo "== bench"; timeout 900 python bench.py --e2e-steps 0 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'],{k:v['ms'] for k,v in d['per_k'].items()}, d['roofline']['frac'])"
