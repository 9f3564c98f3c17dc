#!/usr/bin/env python
"""One step of the ncu launch list (`--metrics gpu__time_duration.sum`) with each kernel's share.

    python profiles/launch_shares.py gpurun_out/launches.csv [out.csv]

The last timed step of the file is reported (cube zeroing through reduce_partials_kernel).
"""

from __future__ import annotations

import csv
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith("==")]
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    launches = [(r[kn], float(r[mv].replace(",", ""))) for r in rows[1:] if len(r) > mv and r[mv]]
    unit = rows[1][hdr.index("Metric Unit")] if "Metric Unit" in hdr else "ns"
    scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
    # a step runs from the fill that zeroes the cube (just before smoothing_setup_kernel) to
    # reduce_partials_kernel; the 512 MiB L2-flush fill between steps belongs to neither
    steps = []
    for i, (k, _) in enumerate(launches):
        if "smoothing_setup_kernel" not in k:
            continue
        a = i - 1 if i > 0 and "FillFunctor<double>" in launches[i - 1][0] else i
        b = next((j for j in range(i, len(launches)) if "reduce_partials_kernel" in launches[j][0]), None)
        if b is not None and any("project_kernel<0" in k2 for k2, _ in launches[a:b + 1]):
            steps.append((a, b + 1))
    a, b = steps[-1]  # the last timed step (the diagnostic COUNT pass uses project_kernel<1, ...>)
    step = launches[a:b]
    total = sum(d for _, d in step) * scale
    out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
    out.write("kernel,duration_ns,share\n")
    for k, d in step:
        name = k.split("(")[0]
        out.write(f'"{name}",{d * scale:.0f},{100 * d * scale / total:.1f}%\n')
    out.write(f'"TOTAL ({len(step)} launches)",{total:.0f},100%\n')


if __name__ == "__main__":
    main()
