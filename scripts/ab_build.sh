#!/bin/bash
# Build named variants of the library for same-box A/B runs: scripts/ab_build.sh name "defs" [name "defs" ...]
set -e
mkdir -p multi-view-registration_b200/variants
while [ $# -ge 2 ]; do
  MVR_NVCC_DEFS="$2" python -c "
import sys; sys.path.insert(0, 'multi-view-registration_b200'); import build
print(build.build(force=True, out='multi-view-registration_b200/variants/lib_$1.so'))"
  shift 2
done
