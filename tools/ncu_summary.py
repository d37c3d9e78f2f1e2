#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, without a GPU) into the few numbers DESIGN.md / profiles/ quote.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--source N]   > profiles/<name>.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, zip(r, units))) for r in rows[2:]]


def main():
    rep = sys.argv[1]
    for k, d in enumerate(raw(rep)):
        name = d.get("Kernel Name", ("?", ""))[0]
        print(f"## launch {k}: `{name}`\n")
        print("| metric | value | unit |\n|---|---|---|")
        for key in KEYS:
            if key in d:
                print(f"| {key} | {d[key][0]} | {d[key][1]} |")
        print("\n| stall (warps per issue-active cycle) | value |\n|---|---|")
        for h, (v, _) in sorted(d.items()):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
                try:
                    if float(v) >= 0.05:
                        print(f"| {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} | {float(v):.2f} |")
                except ValueError:
                    pass
        print()
    if "--source" in sys.argv:
        n = int(sys.argv[sys.argv.index("--source") + 1])
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if rows:
            hdr = rows[0]
            try:
                si = hdr.index("# Samples") if "# Samples" in hdr else hdr.index("Samples")
            except ValueError:
                si = None
            if si is not None:
                body = [r for r in rows[1:] if len(r) > si and r[si].replace(".", "").isdigit()]
                body.sort(key=lambda r: -float(r[si]))
                print(f"## top {n} source/SASS lines by samples\n")
                print("| samples | line |\n|---|---|")
                for r in body[:n]:
                    print(f"| {r[si]} | `{' '.join(r[:3])[:150]}` |")


if __name__ == "__main__":
    main()
