#!/bin/bash
# round 2, call G (2 GPUs): the new bench blocks and the multi-process parity test at N = 2; prefix-sum after the rework
O=gpurun_out/r2g; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
( time timeout 600 python -m pytest tests/test_gpu_scan.py -m gpu -x -q ) > $O/pytest_scan.log 2>&1; echo "rc=$?" >> $O/pytest_scan.log
timeout 300 python tests/perf/bench_configs.py --config scan > $O/cfg_scan.json 2> $O/cfg_scan.err
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > $O/pytest_multi.log 2>&1; echo "rc=$?" >> $O/pytest_multi.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 ) > $O/bench_n2.json 2> $O/bench_n2.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --halo nccl --skip dense_k,configs,e2e ) > $O/bench_n2_nccl.json 2> $O/bench_n2_nccl.err
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 ) > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err
ls -la $O
