#!/bin/bash
TAG=${1:-g}
O=gpurun_out; mkdir -p $O
export PYTHONUNBUFFERED=1
CVO_B200_MODE=grid timeout 120 python tools/gpu_tails.py > $O/${TAG}_tails_grid.log 2>&1
CVO_B200_MODE=dense timeout 120 python tools/gpu_tails.py > $O/${TAG}_tails_dense.log 2>&1
CVO_B200_MODE=grid timeout 600 ncu --set full --clock-control none --import-source on -k regex:'flow_kernel|step_kernel' -c 6 -o $O/${TAG}_full_grid_c2 -f python tools/gpu_profile.py C2 3 0.95 > $O/${TAG}_ncu_grid.log 2>&1
cat $O/${TAG}_tails_grid.log
