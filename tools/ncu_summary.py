#!/usr/bin/env python
"""Summarise an `ncu --set full` report (.ncu-rep) per kernel instance: duration, DRAM traffic,
tensor-pipe / SM / L2 utilisation, registers, achieved occupancy. Reads with `ncu -i ... --page raw --csv`."""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "dur"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "warp_inst"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    tensor_cols = [h for h in hdr if "pipe_tensor" in h and "pct" in h]
    idx = [(short, hdr.index(name)) for name, short in WANT if name in hdr]
    extra = [(h.split(".")[0][-40:], hdr.index(h)) for h in tensor_cols[:2] if h not in [w[0] for w in WANT]]
    ik, ig, ib = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size")
    print(f"# {path}")
    for r in rows[2:]:
        name = r[ik].split("(")[0][-40:]
        parts = [f"{short}={r[i]}{units[i] if short in ('dur', 'dram_rd', 'dram_wr') else ''}" for short, i in idx + extra]
        print(f"{name:40s} grid={r[ig]:>14s} block={r[ib]:>12s} " + " ".join(parts))


if __name__ == "__main__":
    main(sys.argv[1])
