#!/bin/bash
# Round-2 measurement pass on one B200: bench lines, launch list, ncu --set full of the dominant kernels.
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 900 python bench.py > $O/${TAG}_bench.log 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.log 2>&1
# launch list of the default bench command (serialised, cold caches: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-frames > $O/${TAG}_ncu_launch.log 2>&1
# full captures: the persistent kernel on C2 (one launch = one whole registration, cell queries)
# and on C4 (one launch = 50 iterations, tile cells built with a skin and reused)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'align_grid_kernel' -s 3 -c 1 -o $O/${TAG}_full_c2_persist -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-frames --no-anchor > $O/${TAG}_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'align_grid_kernel' -s 3 -c 1 -o $O/${TAG}_full_c4_persist -f python bench.py --workload C4 --steps 1 --warmup 3 --no-cpu-baseline --no-frames --no-anchor >> $O/${TAG}_ncu_full.log 2>&1
tail -c 400 $O/${TAG}_bench.log
