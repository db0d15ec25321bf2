#!/bin/bash
# Development aid: rebuild the library with different brick-kernel shapes ON the GPU box and time one pairwise ICP.
# usage: scripts/exp_variants.sh "T CAP CTAS" ...
for v in "$@"; do
  set -- $v
  export MVR_NVCC_DEFS="-DMVR_BS_THREADS=$1 -DMVR_BS_CAP=$2 -DMVR_BS_CTAS_PER_SM=$3"
  python multi-view-registration_b200/build.py --force > /dev/null 2>&1 || { echo "build failed for $v"; continue; }
  echo "== threads $1 cap $2 ctas/sm $3"
  python scripts/gpu_quick.py 2>&1 | grep -E "ICP n|corr |reduce|table|sort|transform|NN m" | head -14
done
