#!/bin/bash
# Multi-GPU round: torchrun bench at N ranks (IPC peer halo, then NCCL halo), and the sharded parity test.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
echo "== bench N=$N (ipc halo)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
echo "== bench N=$N (nccl halo)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 3 --halo nccl --e2e-steps 0 > gpurun_out/bench_n${N}_nccl.json 2> gpurun_out/bench_n${N}_nccl.err; echo "rc=$?"; cat gpurun_out/bench_n${N}_nccl.json; tail -5 gpurun_out/bench_n${N}_nccl.err
echo "== multi-device tests"
timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/pytest_multi_$N.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_multi_$N.log
