#!/bin/bash
# round 2: bench at 4 GPUs, as the driver launches it
O=gpurun_out/r2n4; mkdir -p $O
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 20 --warmup 5 ) > $O/bench_n4.json 2> $O/bench_n4.err
tail -4 $O/bench_n4.err
