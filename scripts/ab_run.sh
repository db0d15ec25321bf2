#!/bin/bash
# Same-box A/B of library variants: scripts/ab_run.sh out.log name [name ...]
OUT=$1; shift
: > $OUT
for round in 1 2; do
  for v in "$@"; do
    echo "== variant $v (round $round)" >> $OUT
    MVR_B200_LIB=$PWD/multi-view-registration_b200/variants/lib_$v.so MVR_GROUPS=${MVR_GROUPS:-8} python scripts/gpu_group.py 2>&1 | grep -v "^$" >> $OUT
    MVR_B200_LIB=$PWD/multi-view-registration_b200/variants/lib_$v.so python scripts/prof_icp.py >> $OUT 2>&1
    MVR_B200_LIB=$PWD/multi-view-registration_b200/variants/lib_$v.so python scripts/prof_icp.py 200000 30 0 >> $OUT 2>&1
  done
done
