"""Tiny case for compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import unified_cvo_b200 as u
from helpers import *
src, tgt, _ = synthetic_pair(700, 500, 600, 7)
p = geometric_params()
g = u.CvoGPU(p)
g.set_cloud(0, src); g.set_cloud(1, tgt)
tr = g.iterate(np.eye(3), np.zeros(3), 0.95, 256)
print("nnz", tr.nnz, "omega", list(tr.omega), "step", tr.step)
p.MAX_ITER = 40
g2 = u.CvoGPU(p)
ret, T, info = g2.align(src, tgt)
print("align", ret, info.iterations, info.stop_reason)
