#!/bin/bash
# Build the experimental kernel variants as martini_b200/lib_var_<name>.so for
# scripts/try_variants.sh (which times bench.py and runs the GPU parity tests with each).
# A variant is a set of -D switches of csrc/; every one of them is logic-checked on the CPU by
# tests/test_emu_variants.py under the SIMT emulator before it is given GPU time.
#   usage: scripts/build_variants.sh [name=-DFLAG[,-DFLAG...]]...   (default: the queued set)
cd "$(dirname "$0")/.."
variants=("$@")
if [ ${#variants[@]} -eq 0 ]; then
  variants=(
    "base="
    "ws=-DMTN_FOOTREC=0"
    "footrec1=-DMTN_FOOTREC=1"
    "gauss_sep=-DMTN_GAUSS_SEP=1"
    "wtab_more0=-DMTN_WTAB_MORE=0"
  )
fi
for v in "${variants[@]}"; do
  name="${v%%=*}"
  flags="${v#*=}"
  echo "== lib_var_${name}.so  ${flags}"
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC \
    ${flags//,/ } -Xptxas -v -o "martini_b200/lib_var_${name}.so" martini_b200/csrc/api.cu 2>&1 |
    grep -A2 "project_kernelILb0ELi0" | grep -E "spill|registers"
done
