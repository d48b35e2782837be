#!/usr/bin/env python
"""Warp-stall picture of one kernel from an .ncu-rep captured with `--set full --import-source on`:
total samples per stall reason, then the SASS instructions that collected the most samples.

    python tools/ncu_stalls.py report.ncu-rep score_topk [min_samples] [instance]

This is how the retrieval epilogue's instruction-fetch stalls (stall_no_inst) and later the MMA <-> scan
ping-pong (hot t_full AND t_empty waits) were found; see DESIGN.md section 7."""
import csv
import io
import subprocess
import sys


def main():
    rep, pattern = sys.argv[1], sys.argv[2]
    min_samples = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    instance = int(sys.argv[4]) if len(sys.argv) > 4 else -1
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pattern}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    if not starts:
        sys.exit("no kernel matched (was the report captured with --import-source on?)")
    blocks = []
    for si, s in enumerate(starts):
        e = starts[si + 1] if si + 1 < len(starts) else len(rows)
        hdr, body = rows[s + 1], rows[s + 2:e]
        ci = hdr.index("# Samples")
        blocks.append((rows[s][1], hdr, body, sum(int(r[ci]) for r in body if r[ci].isdigit())))
    # ncu lists every kernel twice (SASS / source views); pick the requested or the largest instance
    name, hdr, body, total = blocks[instance] if instance >= 0 else max(blocks, key=lambda b: b[3])
    ci, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [j for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = {}
    for r in body:
        for j in stall_cols:
            if r[j].isdigit():
                tot[hdr[j]] = tot.get(hdr[j], 0) + int(r[j])
    print(f"# {name[:90]}\n# {len(body)} SASS instructions, {total} samples")
    for k, v in sorted(tot.items(), key=lambda x: -x[1])[:10]:
        print(f"  {k:28s} {v:8d}  {100.0 * v / max(total, 1):5.1f} %")
    thr = min_samples or max(total // 200, 1)
    print(f"# instructions with >= {thr} samples: index, address, samples, executions, SASS, top-2 reasons")
    for i, r in enumerate(body):
        n = int(r[ci]) if r[ci].isdigit() else 0
        if n >= thr:
            st = sorted(((int(r[j]), hdr[j]) for j in stall_cols if r[j].isdigit()), reverse=True)[:2]
            print(f"{i:5d} {r[0][-5:]} {n:7d} {int(r[ii]):10d} {r[1][:64]:64s} {st}")


if __name__ == "__main__":
    main()
