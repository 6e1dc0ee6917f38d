python tests/perf/pcie_probe.py 2>&1 | tail -3
