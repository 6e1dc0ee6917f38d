#!/bin/bash
# round 2, final check: whole GPU suite, kernel smoke, smoke(), bench (both arms)
O=gpurun_out/r2final; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 300 python tests/perf/smoke_all_kernels.py > $O/smoke_all.log 2>&1; echo "rc=$?" >> $O/smoke_all.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
( time timeout 600 python bench.py ) > $O/bench_default.json 2> $O/bench_default.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/bench_ref.json 2> $O/bench_ref.err
tail -3 $O/pytest.log; tail -2 $O/smoke_all.log; tail -2 $O/smoke.log; tail -3 $O/bench_default.err; cut -c1-400 $O/bench_ref.json
