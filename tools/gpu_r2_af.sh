#!/bin/bash
# round 2, call AF: kernel time against the distance between input and output buffer
O=gpurun_out/r2af; mkdir -p $O
timeout 300 python tests/perf/alias_probe.py > $O/alias.json 2> $O/alias.err
cat $O/alias.json; tail -3 $O/alias.err
