#!/bin/bash
mkdir -p gpurun_out
for t in "prefetch=1" "prefetch=3" "prefetch=4" "prefetch=2,chunks_per_cta=2" "prefetch=4,chunks_per_cta=32" "prefetch=3,chunks_per_cta=16"; do
echo "== cfg 5i $t"; timeout 600 python tools/bench_configs.py --config 5i --log2 30 --tune "$t" 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['tune'], round(d['ms'],4), round(d['hbm_gbs_per_gpu'],1))"
done
