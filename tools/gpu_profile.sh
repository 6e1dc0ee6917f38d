#!/bin/bash
# GPU round: full parity suite, bench, then the two ncu passes of B200_PROFILING.md on a short bench command.
mkdir -p gpurun_out
echo "== pytest (all gpu tests)"; timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench --impl reference"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench_ref.json
CMD="python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline --ks 3,64,4096"
echo "== ncu launch list"
$CMD > gpurun_out/plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full"
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:stream_f32 -s 9 -c 3 -f -o gpurun_out/prof_stream $CMD > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/
