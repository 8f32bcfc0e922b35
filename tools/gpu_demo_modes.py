"""The README demo pair (BASELINE configs[0]) through align() with every candidate generator forced
and with the automatic policy: seconds per registration, iterations, share of persistent batches.
usage: gpu_demo_modes.py [modes, comma separated: auto,dense,grid,tile,brute]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import unified_cvo_b200 as u
from helpers import demo_clouds, demo_params

modes = (sys.argv[1] if len(sys.argv) > 1 else "auto,dense,grid,tile,brute").split(",")
for color in (True, False):
    src, tgt = demo_clouds(color=color)
    p = demo_params(src, tgt, color=color)
    for mode in modes:
        if mode == "auto":
            os.environ.pop("CVO_B200_MODE", None)
        else:
            os.environ["CVO_B200_MODE"] = mode
        g = u.CvoGPU(p)
        g.align_host(src, tgt)
        t0 = time.perf_counter()
        ret, T, info = g.align_host(src, tgt)
        wall = time.perf_counter() - t0
        print(f"demo colour={color} mode={mode}: {info.registration_seconds:.4f} s device, {wall:.4f} s wall, "
              f"{info.iterations} iterations ({info.registration_seconds / max(info.iterations, 1) * 1e6:.1f} us each), "
              f"stop {info.stop_reason}, persistent/cell-query batches {info.cell_query_fraction:.3f}, "
              f"final ell {info.final_ell:.4f}", flush=True)
        g.close()
