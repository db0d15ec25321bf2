#!/bin/bash
# One GPU-box visit that produces everything profiles/ is made from: scripts/gpu_final.sh TAG
#   parity tests, smoke, the bench line and the reference arm, the nn_sweep line, the ncu launch list of the bench command and
#   one ncu --set full capture each of the forward and the reverse half (each only after its command ran clean without ncu).
set -u
TAG=${1:-r02z}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_ref.json 2>> gpurun_out/${TAG}_bench.err; echo "reference arm rc=$?"
python bench.py --workload nn_sweep --steps 3 --warmup 3 > gpurun_out/${TAG}_nn_sweep.json 2>> gpurun_out/${TAG}_bench.err; echo "nn_sweep rc=$?"
python scripts/gpu_normals_time.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_normals.log
SMALL="python bench.py --steps 2 --warmup 1 --no-e2e --cpu-sample-pairs 0"
$SMALL > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${TAG}_launches.csv $SMALL > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
export MVR_ITERS=30 MVR_GROUP=24
python scripts/gpu_iter_profile.py > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_icp_forward -s 20 -c 1 -f -o gpurun_out/${TAG}_fwd python scripts/gpu_iter_profile.py > gpurun_out/${TAG}_ncu_fwd.log 2>&1
echo "ncu fwd rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_icp_reverse -s 20 -c 1 -f -o gpurun_out/${TAG}_rev python scripts/gpu_iter_profile.py > gpurun_out/${TAG}_ncu_rev.log 2>&1
echo "ncu rev rc=$?"
