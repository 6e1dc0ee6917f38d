#!/bin/bash
# round 2: bench at N GPUs (argument), as the driver launches it
N=${1:-8}
O=gpurun_out/r2s; mkdir -p $O
nvidia-smi topo -m > $O/topo_$N.txt 2>&1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 ) > $O/bench_n$N.json 2> $O/bench_n$N.err
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > $O/pytest_multi_$N.log 2>&1; echo "rc=$?" >> $O/pytest_multi_$N.log
ls -la $O
