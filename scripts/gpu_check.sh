#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, then (optionally) ncu launch list + full capture.
# usage: scripts/gpu_check.sh [tag] [ncu_kernel_regex]
set -u
TAG=${1:-r01}
KREGEX=${2:-k_icp}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_${TAG}.log
python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke_${TAG}.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
if [ "${NCU:-1}" = "1" ]; then
  SMALL="python bench.py --steps 1 --warmup 3 --no-e2e --cpu-sample-pairs 0"   # the bench workload itself, one timed step
  $SMALL > gpurun_out/plain_${TAG}.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches_${TAG}.csv $SMALL > gpurun_out/ncu_launches_${TAG}.log 2>&1
  echo "ncu launches rc=$?"
  $SMALL > gpurun_out/plain2_${TAG}.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s 200 -c 2 -f -o gpurun_out/prof_${TAG} $SMALL > gpurun_out/ncu_full_${TAG}.log 2>&1
  echo "ncu full rc=$?"
fi
