#!/bin/bash
# Build experimental kernel variants as martini_b200/lib_var_<name>.so (same C ABI; load one
# with MTN_B200_LIB=... ).  A variant is a set of -D switches of csrc/; each is logic-checked
# on the CPU by tests/test_emu_variants.py under the SIMT emulator before it gets GPU time.
#   usage: scripts/build_variants.sh name=-DFLAG[,-DFLAG...] ...
cd "$(dirname "$0")/.."
for v in "$@"; do
  name="${v%%=*}"
  flags="${v#*=}"
  echo "== lib_var_${name}.so  ${flags}"
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC \
    ${flags//,/ } -Xptxas -v -o "martini_b200/lib_var_${name}.so" martini_b200/csrc/api.cu 2>&1 |
    grep -A2 "project_kernelILb0ELi0" | grep -E "spill|registers"
done
