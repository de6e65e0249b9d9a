#!/usr/bin/env python3
"""Turns an `ncu --csv --metrics ...` log of one bench step into profiles/rNN_viterbi_pipe_counts.json: the DYNAMIC
number of warp-instructions k_viterbi executed on each pipe, per decoded frame of the bench workload.  bench.py
multiplies the alu figure by the frames it decodes and divides by the kernel's CUDA-event time measured live; the peak
it divides by is measured live too (wifi_b200_alu_peak).
usage: ncu_pipe_counts.py <ncu.csv> <out.json> <frames in the profiled launch> "<how the log was taken>" """
import csv
import json
import sys

WANT = {"smsp__inst_executed.sum": "warp_inst", "smsp__inst_executed_pipe_alu.sum": "alu_warp_inst", "smsp__inst_executed_pipe_fma.sum": "fma_warp_inst",
        "smsp__inst_executed_pipe_lsu.sum": "lsu_warp_inst", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "gpu__time_duration.sum": "ns", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pct"}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "msecond": 1e6, "usecond": 1e3, "nsecond": 1.0, "second": 1e9}


def main():
    path, out, frames, how = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    kn, mn, mu, mv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    per_kernel = {}
    for r in rows[1:]:
        name = r[kn].split("(")[0].replace("void ", "")
        if r[mn] in WANT:
            v = float(r[mv].replace(",", "")) * UNIT.get(r[mu], 1.0)
            per_kernel.setdefault(name, {}).setdefault(WANT[r[mn]], []).append(v)
    k = {m: v[-1] for m, v in per_kernel["k_viterbi"].items()}        # the last captured launch
    res = {"source": how, "kernel": "k_viterbi", "psdu_len": 1528, "n_dbps": 216, "frames_in_profiled_launch": frames,
           "alu_warp_inst_per_frame": k["alu_warp_inst"] / frames, "fma_warp_inst_per_frame": k.get("fma_warp_inst", 0) / frames,
           "lsu_warp_inst_per_frame": k.get("lsu_warp_inst", 0) / frames, "warp_inst_per_frame": k["warp_inst"] / frames,
           "dram_bytes_per_frame": (k.get("dram_read", 0) + k.get("dram_write", 0)) / frames, "ncu_duration_ms": k.get("ns", 0) / 1e6,
           "ncu_alu_pipe_pct_of_peak": k.get("alu_pct"),
           "all_kernels": {n: {m: v[-1] for m, v in d.items()} for n, d in per_kernel.items()}}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps({a: res[a] for a in ("alu_warp_inst_per_frame", "warp_inst_per_frame", "dram_bytes_per_frame", "ncu_alu_pipe_pct_of_peak")}))


main()
