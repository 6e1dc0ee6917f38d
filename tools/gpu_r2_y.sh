#!/bin/bash
# round 2, call Y (2 GPUs): multi-process parity test and the bench line at N = 2 after the kernel changes of the second half
O=gpurun_out/r2y; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_dropin.py -m gpu -x -q ) > $O/pytest_multi.log 2>&1; echo "rc=$?" >> $O/pytest_multi.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 ) > $O/bench_n2.json 2> $O/bench_n2.err
tail -4 $O/pytest_multi.log; tail -3 $O/bench_n2.err
ls -la $O
