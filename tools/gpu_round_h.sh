#!/bin/bash
# r01h: tests + default bench (with CPU baseline) + reference arm + launch list of the edge loop
TAG=${1:-r01h}
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; tail -3 $O/${TAG}_pytest.log
timeout 400 python bench.py > $O/${TAG}_bench.log 2>&1
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/${TAG}_launches_edges.csv python tools/gpu_edges.py 2 > $O/${TAG}_ncu_edges.log 2>&1
tail -c 1200 $O/${TAG}_bench.log
tail -c 300 $O/${TAG}_ncu_edges.log
