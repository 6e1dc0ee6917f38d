#!/bin/bash
# BASELINE configs 4/5 + int16 on N GPUs (N=1: plain python; N>1: torch.distributed.run)
N=${1:-1}
mkdir -p gpurun_out
run() {
  if [ "$N" = "1" ]; then timeout 900 python tests/perf/bench_configs.py "$@";
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 tests/perf/bench_configs.py "$@"; fi
}
for c in 4 5p 5i; do
  echo "== config $c N=$N"; run --config $c > gpurun_out/cfg${c}_n$N.json 2> gpurun_out/cfg${c}_n$N.err; echo "rc=$?"; cat gpurun_out/cfg${c}_n$N.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/cfg${c}_n$N.err | tail -5
done
if [ "$N" = "1" ]; then echo "== i16"; run --config i16 > gpurun_out/cfg_i16_n1.json 2> gpurun_out/cfg_i16_n1.err; echo "rc=$?"; cat gpurun_out/cfg_i16_n1.json; tail -3 gpurun_out/cfg_i16_n1.err; fi
