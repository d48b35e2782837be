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


def gemm_traffic_json(path, out_path, min_bytes=20e6):
    """Average DRAM bytes (read + write) per launch over the encoder-sized tt::gemm_bf16_kernel launches of a
    `--set full` capture of one training step -> the `traffic` figure bench.py reports next to the roofline."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    ir, iw, idur = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, n, dur = 0.0, 0, 0.0
    for r in rows[2:]:
        if "gemm_bf16_kernel" not in r[ik]:
            continue
        b = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
        if b >= min_bytes:
            tot += b
            n += 1
            dur += float(r[idur])
    res = {"source": path.split("/")[-1], "encoder_gemm_launches": n,
           "encoder_gemm_dram_bytes_per_launch": tot / max(n, 1),
           "encoder_gemm_dram_bytes_per_step": tot,
           "note": "dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, cold L2 per replay; writes still "
                   "resident in the 126 MB L2 when a kernel ends are not counted"}
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


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
    if len(sys.argv) >= 4 and sys.argv[1] == "--gemm-traffic":
        gemm_traffic_json(sys.argv[2], sys.argv[3])
    else:
        main(sys.argv[1])
