#!/usr/bin/env python
"""Attribute the executed warp instructions and stall samples of an ncu report to CUDA source
lines (no GPU needed).

    python profiles/sass_by_line.py gpurun_out/prof.ncu-rep martini_b200/libmartini_b200.so [min_pct]

ncu's CSV source page is SASS-only, so the line table comes from `nvdisasm -g` on the cubin
extracted from the library the profile was taken with; both list the kernel's instructions in
address order and are joined by position.
"""

from __future__ import annotations

import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def line_table(lib, kernel_mangled):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
        cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
    table, cur, on = [], ("?", 0), False
    for ln in dis.splitlines():
        if ln.startswith(".text."):
            on = kernel_mangled in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            table.append((int(m.group(1), 16), cur, m.group(2).strip()))
    return table


def main():
    rep, lib = sys.argv[1], sys.argv[2]
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # one section per profiled kernel, each starting with its own header row; NCU_SECTION picks one
    heads = [i for i, r in enumerate(rows) if len(r) > 3 and r[1] == "Source"]
    k = int(os.environ.get("NCU_SECTION", "0"))
    hdr = rows[heads[k]]
    end = heads[k + 1] if k + 1 < len(heads) else len(rows)
    body = [dict(zip(hdr, r)) for r in rows[heads[k] + 1:end] if len(r) == len(hdr)]
    table = line_table(lib, sys.argv[4] if len(sys.argv) > 4 else "project_kernelILb0")
    assert len(table) == len(body), (len(table), len(body))
    agg = collections.OrderedDict()
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for (off, key, sass), d in zip(table, body):
        a = agg.setdefault(key, collections.Counter())
        a["inst"] += int(d["Instructions Executed"])
        a["samples"] += int(d["# Samples"])
        for s in stall_cols:
            a[s] += int(d[s] or 0)
    ti = sum(a["inst"] for a in agg.values())
    ts = sum(a["samples"] for a in agg.values())
    print(f"total warp instructions {ti}, stall samples {ts}")
    print("file:line  inst%  samples%  top stalls")
    for key in sorted(agg, key=lambda k: (k[0] != "project.cuh", k[0], k[1])):  # (file, line) order
        a = agg[key]
        if 100 * a["inst"] / ti < min_pct and 100 * a["samples"] / ts < min_pct:
            continue
        top = sorted(((a[s], s[6:]) for s in stall_cols), reverse=True)[:3]
        tops = " ".join(f"{n}:{100 * v / max(1, a['samples']):.0f}%" for v, n in top if v)
        print(f"{key[0]}:{key[1]:<5} {100 * a['inst'] / ti:5.1f} {100 * a['samples'] / ts:6.1f}   {tops}")
    by_file = collections.Counter()
    for key, a in agg.items():
        by_file[key[0]] += a["inst"]
    print("by file:", {k: f"{100 * v / ti:.1f}%" for k, v in by_file.most_common()})


if __name__ == "__main__":
    main()
