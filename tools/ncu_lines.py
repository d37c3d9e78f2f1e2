#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares of the kernels in an .ncu-rep (read here, without a GPU).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top N]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    cur, hdr, rows = None, None, []
    for r in csv.reader(io.StringIO(out)):
        if len(r) == 2 and r[0] == "File Path":
            cur, hdr = r[1], None
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) >= 9 and r[0]:
            try:
                rows.append((int(r[7] or 0), int(r[6] or 0), cur.split("/")[-1], int(r[0]), r[1].strip()[:100], r))
            except ValueError:
                pass
    ti, ts = sum(x[0] for x in rows) or 1, sum(x[1] for x in rows) or 1
    print(f"total warp instructions {ti}, stall samples {ts}")
    print("by instructions executed:")
    for x in sorted(rows, key=lambda x: -x[0])[:top]:
        print(f"{100.0 * x[0] / ti:5.1f}% inst {100.0 * x[1] / ts:5.1f}% smp  {x[2]}:{x[3]}  {x[4]}")
    print("by stall samples:")
    for x in sorted(rows, key=lambda x: -x[1])[:top // 2]:
        print(f"{100.0 * x[0] / ti:5.1f}% inst {100.0 * x[1] / ts:5.1f}% smp  {x[2]}:{x[3]}  {x[4]}")


if __name__ == "__main__":
    main()
