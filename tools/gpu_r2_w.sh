#!/bin/bash
# round 2, call W: bench with the shapes block; reference arm
O=gpurun_out/r2w; mkdir -p $O
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $O/bench_full.json 2> $O/bench_full.err
tail -3 $O/bench_full.err
ls -la $O
