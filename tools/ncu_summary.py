#!/usr/bin/env python
"""Condense an ncu report (--set full) into a small CSV/markdown table for profiles/.

  python tools/ncu_summary.py gpurun_out/prof_stream.ncu-rep profiles/r01/ncu_stream_full
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    if rep.endswith(".csv"):          # already exported with `ncu -i x.ncu-rep --page raw --csv` (on the GPU box)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    cols = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    with open(out + ".csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [r[name_i].split("(")[0] for r in data])
        for m, i in cols:
            w.writerow([m, units[i]] + [r[i] for r in data])
    with open(out + ".md", "w") as f:
        f.write("| metric | unit | " + " | ".join(r[name_i].split("(")[0].replace("void ", "") for r in data) + " |\n")
        f.write("|---|---|" + "---|" * len(data) + "\n")
        for m, i in cols:
            f.write(f"| {m} | {units[i]} | " + " | ".join(r[i] for r in data) + " |\n")
    print(open(out + ".md").read())


if __name__ == "__main__":
    main()
