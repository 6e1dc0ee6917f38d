#!/bin/bash
# Scaling round on N GPUs: headline bench + configs 4, 5p, 5i + the multi-GPU parity test.
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
echo "== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; cut -c1-1500 gpurun_out/bench_n$N.json; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/bench_n$N.err | tail -5
bash tools/gpu_configs.sh $N
echo "== multi-GPU parity"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/pytest_multi_$N.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_multi_$N.log
