#!/usr/bin/env python
"""Instruction and stall-sample shares of classify2_kernel by code region (line ranges located by marker comments).

    python tools/ncu_regions2.py gpurun_out/prof.ncu-rep
"""
import csv, io, subprocess, sys, os, re
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; hdr = None; rows = []
for r in csv.reader(io.StringIO(out)):
    if len(r) == 2 and r[0] == "File Path": cur = r[1]; hdr = None
    elif r and r[0] == "Line No": hdr = r
    elif hdr and len(r) >= 9 and r[0]:
        try: rows.append((cur.split('/')[-1], int(r[0]), int(r[7] or 0), int(r[6] or 0), r[1]))
        except ValueError: pass
ti = sum(x[2] for x in rows) or 1; ts = sum(x[3] for x in rows) or 1
# regions of slk_group.h from the imported source itself
marks = [("slk_g_pairrev32(uint32_t x)", "chunk scan (m-mers, window minima)"), ("struct slk_group_overflow", "overflow structs / hist"),
         ("void slk_group_classify(", "setup + lambdas"), ("auto prepare = [&]()", "prepare (next chunk loads)"),
         ("if (flush) {", "close: lookup pipeline"), ("// 2. merge:", "close: merge"),
         ("// ================= one step", "step head + run ends"), ("// ---- entries of this step", "prefix sum + emission"),
         ("// ---- advance the cursor", "advance"), ("if (l_have_cur) l_nh++;", "finish: resolve"), ("// ---- kernel body", "kernel epilogue"),
         ("slk_g_warp_alloc", "kernel epilogue")]
src = {}
for f, l, i, s, t in rows:
    src.setdefault(f, {})[l] = t
bounds = []
if "slk_group.h" in src:
    for l in sorted(src["slk_group.h"]):
        for m, nm in marks:
            if m in src["slk_group.h"][l] and not any(b[1] == nm for b in bounds): bounds.append((l, nm))
bounds.sort()
core = [(155, 180, 'compress'), (205, 222, 'hash'), (222, 234, 'next_bucket'), (244, 263, 'match'), (263, 272, 'probe_rest'), (280, 297, 'lca'), (298, 311, 'min62'), (395, 440, 'resolve_tree')]
agg = {}
for f, l, i, s, t in rows:
    name = f + ":other"
    if f == "slk_group.h":
        name = "slk_group.h:head"
        for bl, nm in bounds:
            if l >= bl: name = nm
    elif f == "slk_core.h":
        for lo, hi, nm in core:
            if lo <= l < hi: name = "core: " + nm
    a = agg.setdefault(name, [0, 0]); a[0] += i; a[1] += s
print(f"total warp instructions {ti}, stall samples {ts}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if v[0] * 1000 > ti: print(f"{100*v[0]/ti:5.1f}% inst  {100*v[1]/ts:5.1f}% samples  {k}")
