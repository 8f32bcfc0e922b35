"""Workload for ncu: a few fixed-state iterations of a named config."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import unified_cvo_b200 as u
from helpers import *
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ell = float(sys.argv[3]) if len(sys.argv) > 3 else 0.95
P, N, M, seed, F, C = u.synthetic.CONFIGS[name]
src, tgt, _ = synthetic_pair(P, N, M, seed, F=F, C=C)
p = geometric_params() if F == 0 else u.read_params_yaml(os.path.join(DATA, "cvo_intensity_params_img_gpu0.yaml"))
g = u.CvoGPU(p)
g.set_cloud(0, src); g.set_cloud(1, tgt)
ms, msp = g.time_iterations(np.eye(3), np.zeros(3), ell, 256, iters, pair_kernel=False)
print(name, "iters", iters, "ms/iter", ms / iters)
