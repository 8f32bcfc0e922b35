"""Where a round of the batched pose-graph edge loop spends its time (host side): cProfile over
update_edges on bench.py's ring.  python tools/gpu_edges_profile.py [rounds]"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import unified_cvo_b200 as u  # noqa: E402
from unified_cvo_b200 import synthetic  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 50
src, tgt, p, _ = bench.load_workload("KITTI05_TRACK")
g = u.CvoGPU(p)
I = np.eye(4)[:3]
G = np.asarray(synthetic.gt_transform(), np.float64)[:3]
frames = [u.CvoFrameGPU(g, c, P) for c, P in ((src, I), (tgt, G), (src, I), (tgt, G))]
states = [u.BinaryStateGPU(frames[i], frames[(i + 1) % 4], int(p.multiframe_num_neighbors), 0.25) for i in range(4)]
for _ in range(3):
    u.update_edges(states)
t0 = time.perf_counter()
for _ in range(rounds):
    u.update_edges(states)
dt = time.perf_counter() - t0
print(f"batched: {rounds * 4 / dt:.0f} edge updates/s, {1e3 * dt / rounds:.3f} ms per round of 4")
pr = cProfile.Profile()
pr.enable()
for _ in range(rounds):
    u.update_edges(states)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
g.close()
