"""Per-iteration time of every candidate generator at fixed states.
usage: gpu_modes.py CONFIG ell [ell ...]   (CONFIG in C2 | KITTI05 | C4 | C5)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, unified_cvo_b200 as u
from helpers import *
name = sys.argv[1]
P, N, M, seed, F, C = u.synthetic.CONFIGS[name]
src, tgt, _ = synthetic_pair(P, N, M, seed, F=F, C=C)
p = geometric_params() if F == 0 else u.read_params_yaml(os.path.join(DATA, "cvo_intensity_params_img_gpu0.yaml"))
for mode in (os.environ.get("MODES") or "dense,tile,grid,auto").split(","):
    os.environ["CVO_B200_MODE"] = mode
    g = u.CvoGPU(p); g.set_cloud(0, src); g.set_cloud(1, tgt)
    if os.environ.get("ROWS"):
        g.set_row_range(0, int(os.environ["ROWS"]))  # one rank's shard of a sharded job
    for ell in [float(x) for x in sys.argv[2:]]:
        iters = 20 if N > 50000 else 200
        g.time_iterations(np.eye(3), np.zeros(3), ell, 256, iters)
        ms, msk = g.time_iterations(np.eye(3), np.zeros(3), ell, 256, iters)
        tr = g.iterate(np.eye(3), np.zeros(3), ell, 256)
        print(name, mode, "ell", ell, "us/iter %.1f" % (ms / iters * 1e3), "generator kernel us %.1f" % (msk / iters * 1e3), "nnz", tr.nnz, flush=True)
    g.close()
