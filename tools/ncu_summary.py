#!/usr/bin/env python3
"""Summarise an `ncu --set full` report into the JSON kept under profiles/ (per-launch metrics of each captured
kernel).  usage: ncu_summary.py <report.ncu-rep> <out.json> "<how the report was taken>" """
import csv
import io
import json
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size", "launch__block_size",
           "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static"]


def main():
    rep, out, how = sys.argv[1:4]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    kernels = []
    for r in rows[2:]:
        k = {"kernel": r[kn].split("(")[0].replace("void ", "")}
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                k[m] = {"value": r[i], "unit": units[i]}
        kernels.append(k)
    json.dump({"source": how, "kernels": kernels}, open(out, "w"), indent=1)
    for k in kernels:
        print(k["kernel"], k.get("gpu__time_duration.sum"), k.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"))


main()
