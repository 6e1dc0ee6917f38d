#!/bin/bash
# round 2, call N: more tile-shape variants of the headline kernel
O=gpurun_out/r2n; mkdir -p $O
SK=parity,dense_k,i16,configs,e2e,steps
for t in "threads=256,ctas_per_sm=1,prefetch=4" "threads=256,ctas_per_sm=1,prefetch=6" "threads=256,run=32,ctas_per_sm=1,prefetch=2" "prefetch=2" "threads=256,ctas_per_sm=2,prefetch=2,chunks_per_cta=1"; do
  timeout 300 python bench.py --steps 20 --warmup 5 --tune $t --skip $SK --no-cpu-baseline > "$O/bench_$t.json" 2> "$O/bench_$t.err"
done
ls $O
