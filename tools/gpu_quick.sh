#!/bin/bash
mkdir -p gpurun_out
echo "== pytest i16"; timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py tests/test_gpu_scan.py -m gpu -q -x -k "i16 or golden or int16 or multi_device or prefix_difference" > gpurun_out/pytest_i16.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_i16.log
echo "== i16 bench"; timeout 600 python tools/bench_configs.py --config i16 > gpurun_out/cfg_i16_n1.json 2> gpurun_out/cfg_i16_n1.err; echo "rc=$?"; python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/cfg_i16_n1.json') if l.startswith('{')][0]
print({k:(v['ms'],v['hbm_gbs'],v['bit_exact_head']) for k,v in d['per_k'].items()})"; tail -3 gpurun_out/cfg_i16_n1.err
