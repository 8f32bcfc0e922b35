import os, sys
os.environ["CVO_B200_DEBUG_TAILS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, unified_cvo_b200 as u
from helpers import *
src, tgt, _ = synthetic_pair(12500, 10000, 10000, 20002)
g = u.CvoGPU(geometric_params()); g.set_cloud(0, src); g.set_cloud(1, tgt)
for ell in (0.95, 0.95, 0.3, 0.1):
    ms, _ = g.time_iterations(np.eye(3), np.zeros(3), ell, 64, 300, pair_kernel=False)
    print("ell", ell, "us/iter", ms / 300 * 1e3)
