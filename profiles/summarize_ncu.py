#!/usr/bin/env python
"""Summarise an Nsight Compute report of the projection kernel into a small text file.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep [out.md]

Reads the report with `ncu -i ... --page raw --csv` and `--page source --csv` (runs here, no
GPU needed) and prints: duration, DRAM bytes, pipe utilisation, issue rate, stall mix, and
the share of executed warp instructions per kernel phase (phases are delimited by the
BAR.SYNC instructions in the SASS, in program order).
"""

from __future__ import annotations

import csv
import io
import subprocess
import sys

KEYS = (
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "sm__cycles_elapsed.max",
    "sm__cycles_active.avg",
    "sm__cycles_active.min",
    "sm__cycles_active.max",
    "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
)


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True)
    return list(csv.reader(io.StringIO(out.stdout)))


def main():
    rep = sys.argv[1]
    lines = [f"# ncu summary of `{rep.split('/')[-1]}`", ""]
    raw = ncu_csv(rep, "raw")
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        lines.append(f"## {name}")
        vals = dict(zip(hdr, row))
        un = dict(zip(hdr, units))
        for k in KEYS:
            if k in vals:
                lines.append(f"- `{k}` = {vals[k]} {un[k]}")
        lines.append("- stall mix (warps stalled per issue-active cycle):")
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                v = float(vals[h] or 0)
                if v >= 0.05:
                    lines.append(f"    - {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]}: {v:.2f}")
        lines.append("")
    src = ncu_csv(rep, "source")
    h = src[1]
    ia, isamp, isrc = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
    data = [(r[isrc].strip(), int(r[ia] or 0), int(r[isamp] or 0)) for r in src[2:] if len(r) > ia]
    tot_i = sum(d[1] for d in data) or 1
    tot_s = sum(d[2] for d in data) or 1
    lines.append("## share of executed warp instructions / stall samples between barriers (SASS order)")
    seg_i = seg_s = 0
    start = 0
    nseg = 0
    for i, d in enumerate(data):
        seg_i += d[1]
        seg_s += d[2]
        if "BAR.SYNC" in d[0] or i == len(data) - 1:
            if seg_i / tot_i > 0.002:
                lines.append(f"- SASS [{start}, {i}]: {100 * seg_i / tot_i:5.1f} % of instructions, "
                             f"{100 * seg_s / tot_s:5.1f} % of samples")
            seg_i = seg_s = 0
            start = i + 1
            nseg += 1
    ops = {}
    for d in data:
        op = d[0].split()[1] if d[0].startswith("@") and len(d[0].split()) > 1 else d[0].split()[0]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + d[1]
    lines.append("")
    lines.append("## executed warp instructions by opcode (top 14)")
    for op, n in sorted(ops.items(), key=lambda kv: -kv[1])[:14]:
        lines.append(f"- {op}: {100 * n / tot_i:.1f} %")
    proof = sorted({d[0].split()[1 if d[0].startswith('@') else 0] for d in data
                    if any(t in d[0] for t in ("UBLKCP", "SYNCS", "UTMA"))})
    lines.append("")
    lines.append("## TMA / mbarrier instructions present in the SASS: " + ", ".join(proof))
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
