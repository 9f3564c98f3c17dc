#!/usr/bin/env python
"""Static SASS statistics of one kernel of a built library (no GPU needed):
   python profiles/sass_stats.py martini_b200/libmartini_b200.so [mangled-name-substring]
Counts instructions by mnemonic family -- the evidence that the TMA bulk copies, mbarrier
transactions, warp collectives and FP64 FMAs are in the shipped binary, and how a build-time
variant changes the barrier / load mix."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else "project_kernelILb0ELi0E"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
on, counts, total = False, collections.Counter(), 0
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        on = want in m.group(1)
        continue
    if not on:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        total += 1
        counts[m.group(1).split(".")[0]] += 1
groups = {
    "FP64 (DFMA/DMUL/DADD/DSETP)": ("DFMA", "DMUL", "DADD", "DSETP"),
    "shared-memory loads / stores (LDS/STS)": ("LDS", "STS"),
    "global / table loads (LDG)": ("LDG",),
    "global stores (STG)": ("STG",),
    "block barriers (BAR)": ("BAR",),
    "TMA bulk copies (UBLKCP)": ("UBLKCP",),
    "mbarrier ops (SYNCS)": ("SYNCS",),
    "warp collectives (SHFL/VOTE/VOTEU/REDUX/MATCH)": ("SHFL", "VOTE", "VOTEU", "REDUX", "MATCH"),
    "conversions (F2F/I2F/F2I)": ("F2F", "I2F", "F2I", "I2FP", "F2FP"),
    "local memory, spills (LDL/STL)": ("LDL", "STL"),
    "atomics (ATOMG/ATOMS/RED)": ("ATOMG", "ATOMS", "RED"),
}
print(f"{lib}: kernel *{want}*: {total} SASS instructions")
for name, keys in groups.items():
    print(f"  {name:50s} {sum(counts[k] for k in keys)}")
