#!/bin/bash
# One GPU-box session: breakdown, large workloads, ncu launch list + full capture.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh TAG'
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
( timeout 200 python tools/gpu_breakdown.py C2 0.95; timeout 200 python tools/gpu_breakdown.py C2 0.3; timeout 200 python tools/gpu_breakdown.py KITTI05 0.3 ) > $O/${TAG}_breakdown.log 2>&1
timeout 300 python bench.py --workload KITTI05 --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_kitti05.log 2>&1
timeout 300 python bench.py --workload C4 --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_c4.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pair_kernel|flow_kernel|step_kernel|prep_kernel' -c 12 -o $O/${TAG}_full_c2 -f python tools/gpu_profile.py C2 3 0.95 > $O/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pair_kernel|flow_kernel|step_kernel|prep_kernel' -c 8 -o $O/${TAG}_full_c4 -f python tools/gpu_profile.py C4 2 0.3 > $O/${TAG}_ncu_full_c4.log 2>&1
tail -3 $O/${TAG}_breakdown.log; tail -c 600 $O/${TAG}_bench_c4.log
