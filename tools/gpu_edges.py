"""The pose-graph edge loop alone (bench.py's edge_updates leg), for an ncu launch list:
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file X python tools/gpu_edges.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import unified_cvo_b200 as u  # noqa: E402

print(json.dumps(bench.edge_updates_leg(u, rounds=int(sys.argv[1]) if len(sys.argv) > 1 else 3, cpu=False)))
