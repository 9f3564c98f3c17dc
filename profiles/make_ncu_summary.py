#!/usr/bin/env python
"""Turn `ncu --set full` reports into the machine-readable entries of profiles/ncu_summaries.json.

    python profiles/make_ncu_summary.py gpurun_out/r2_col_cfg3.ncu-rep cfg3 [more.ncu-rep workload ...]

Each capture (scripts/gpu_ncu.sh) comes with a `<name>.hash` sidecar: the hash of csrc/ + the
header at capture time (bench.csrc_hash()).  bench.py quotes `roofline.traffic` from the entry
whose hash matches the sources it runs and whose kernel is the step's dominant one; no match,
no number.  Reads the report with `ncu -i ... --page raw --csv` (runs here, no GPU needed).
"""

from __future__ import annotations

import csv
import io
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "lsu_data_pipe_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "lsu_wavefronts_shared",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum": "lsu_wavefronts_global_ld",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum": "thread_dfma",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
}
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9,
              "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}


def summarise(rep, workload):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    e = {"workload": workload, "file": os.path.basename(rep), "kernel": vals[hdr.index("Kernel Name")].split("(")[0]}
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS:
            x = float(v.replace(",", "")) * UNIT_SCALE.get(u, 1.0)
            e[KEYS[h]] = x
        elif h.startswith("smsp__average_warp") and "per_issue_active" in h:
            name = h.split("issue_stalled_")[-1].split("_per_issue_active")[0] if "issue_stalled_" in h else None
            if name and float(v) >= 0.1:
                e.setdefault("stalls_per_issue", {})[name] = round(float(v), 2)
    e["duration_ms"] = e.pop("duration") * 1e3
    hash_file = os.path.splitext(rep)[0] + ".hash"
    e["csrc_hash"] = open(hash_file).read().strip() if os.path.exists(hash_file) else None
    return e


def main():
    path = os.environ.get("NCU_SUMMARY_OUT") or os.path.join(HERE, "ncu_summaries.json")  # (the GPU box writes under gpurun_out/)
    db = json.load(open(path)) if os.path.exists(path) else {"captures": []}
    args = sys.argv[1:]
    for rep, workload in zip(args[0::2], args[1::2]):
        e = summarise(rep, workload)
        db["captures"] = [c for c in db["captures"] if not (c["file"] == e["file"])] + [e]
        print(json.dumps(e))
    with open(path, "w") as f:
        json.dump(db, f, indent=1)


if __name__ == "__main__":
    main()
