#!/bin/bash
# Final 8-GPU round: headline bench, configs 4/5, multi-GPU parity, and the sweep driver over --gpus 1,2,4,8.
N=${1:-8}
mkdir -p gpurun_out
bash tools/gpu_scale.sh $N
echo "== sweep driver (drop-in binaries, single process driving 1/2/4/8 GPUs)"
rm -f gpurun_out/benchmark_data_sweep.csv
timeout 900 python -m digital_signal_processsing_b200.run_benchmarks --sizes 134217728 --dtype float32 --channels 1 \
    --grades 3,64,1024 --blocks 256 --gpus 1,2,4,$N --bins bin_vec4 --csv gpurun_out/benchmark_data_sweep.csv > gpurun_out/sweep_driver.log 2>&1
echo "rc=$?"; tail -8 gpurun_out/sweep_driver.log; cut -d, -f1,3,4,6-9,15,18,19 gpurun_out/benchmark_data_sweep.csv
