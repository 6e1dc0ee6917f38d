#!/bin/bash
mkdir -p gpurun_out
for t in "" "threads=256" "threads=256,prefetch=3" "threads=256,prefetch=4"; do
echo "== i16 tune=[$t]"; timeout 600 python tools/bench_configs.py --config i16 --tune "$t" 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k:(v['ms'],v['bit_exact_head']) for k,v in d['per_k'].items()})"
done
