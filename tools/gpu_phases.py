"""Per-phase time of the persistent kernel (CVO_B200_STAMPS=1) at fixed states.
usage: gpu_phases.py CONFIG ell [ell ...]   (CONFIG in C2 | KITTI05 | C4 | C5)"""
import os, sys
os.environ["CVO_B200_STAMPS"] = "1"
os.environ.setdefault("CVO_B200_MODE", "grid")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, unified_cvo_b200 as u
from helpers import *
name = sys.argv[1]
P, N, M, seed, F, C = u.synthetic.CONFIGS[name]
src, tgt, _ = synthetic_pair(P, N, M, seed, F=F, C=C)
p = geometric_params() if F == 0 else u.read_params_yaml(os.path.join(DATA, "cvo_intensity_params_img_gpu0.yaml"))
g = u.CvoGPU(p); g.set_cloud(0, src); g.set_cloud(1, tgt)
for ell in [float(x) for x in sys.argv[2:]]:
    iters = 20 if N > 50000 else 200
    g.time_iterations(np.eye(3), np.zeros(3), ell, 256, iters, pair_kernel=False)
    ms, _ = g.time_iterations(np.eye(3), np.zeros(3), ell, 256, iters, pair_kernel=False)
    print(name, "ell", ell, "us/iter", ms / iters * 1e3, flush=True)
