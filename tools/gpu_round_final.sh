#!/bin/bash
# Trimmed final round: benches, launch list, full captures of the two dominant kernels.
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 500 python bench.py > $O/${TAG}_bench.log 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.log 2>&1
timeout 300 python bench.py --workload C4 --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_c4.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-frames > $O/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'align_grid_kernel' -s 3 -c 1 -o $O/${TAG}_full_c2_persist -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-frames > $O/${TAG}_ncu_full.log 2>&1
CVO_B200_MODE=dense timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pair_kernel|flow_kernel|step_kernel|prep_kernel' -c 4 -o $O/${TAG}_full_c4_dense -f python tools/gpu_profile.py C4 1 1.5 >> $O/${TAG}_ncu_full.log 2>&1
tail -c 300 $O/${TAG}_bench.log
