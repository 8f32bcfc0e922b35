"""Small cases for compute-sanitizer (memcheck / racecheck / synccheck), one per launch structure.

    compute-sanitizer --tool memcheck  python tools/gpu_sanitize.py
    compute-sanitizer --tool racecheck python tools/gpu_sanitize.py
    compute-sanitizer --tool synccheck python tools/gpu_sanitize.py

Every candidate generator (dense scan, cell queries, tile cells) in both launch structures (the
persistent cooperative kernel, one launch per phase) and the one-warp-per-row exact walk of the
persistent kernel ("brute"), with and without colour, then the calls that
reuse the pairwise pass (inner product, association export in both kernels) and the pose-graph
edge update (per edge and batched).  Sizes are chosen so that a sanitizer run (10-100x slower)
ends within a minute; results are only printed, parity is the tests' business.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import unified_cvo_b200 as u
from unified_cvo_b200 import multiframe
from helpers import DATA, geometric_params, synthetic_pair

ITER = int(os.environ.get("SANITIZE_ITER", "12"))


def colour_params():
    p = u.read_params_yaml(os.path.join(DATA, "cvo_outdoor_params.yaml"))
    p.is_using_geometric_type = 0
    p.ell_init = 0.9
    return p


def run_mode(mode, persist, colour):
    os.environ["CVO_B200_MODE"] = mode
    os.environ["CVO_B200_PERSIST"] = persist
    if colour:
        src, tgt, _ = synthetic_pair(900, 700, 640, 11, F=5)
        p = colour_params()
    else:
        src, tgt, _ = synthetic_pair(900, 700, 640, 7)
        p = geometric_params()
    p.MAX_ITER = ITER
    g = u.CvoGPU(p)
    g.set_cloud(0, src)
    g.set_cloud(1, tgt)
    tr = g.iterate(np.eye(3), np.zeros(3), 0.95, 64)  # rows cut at the cap: the exact redo runs
    ret, T, info = g.align(src, tgt)
    ip = g.inner_product_gpu(src, tgt, np.eye(4), 0.5)
    a = g.compute_association_gpu(src, tgt, np.eye(4), 0.5)
    k = np.diag([0.3, 0.2, 0.4]).astype(np.float32)
    b = g.compute_association_gpu(src, tgt, np.eye(4), k)
    print(f"mode={mode} persist={persist} colour={colour}: iterate nnz {tr.nnz}, align ret {ret} "
          f"iterations {info.iterations}, inner product {ip:.5f}, association nnz {len(a.vals)} / {len(b.vals)}",
          flush=True)
    g.close()


def run_edges():
    os.environ.pop("CVO_B200_MODE", None)
    os.environ["CVO_B200_PERSIST"] = "1"
    p = colour_params()
    p.multiframe_ell_init = 0.9
    p.multiframe_num_neighbors = 32
    g = u.CvoGPU(p)
    src, tgt, _ = synthetic_pair(900, 700, 640, 30, F=5)
    clouds = [src, tgt, src]
    frames = []
    for i, c in enumerate(clouds):
        pose = np.eye(4)[:3].copy()
        pose[0, 3] = 0.02 * i
        frames.append(multiframe.CvoFrameGPU(g, c, pose))
    edges = [multiframe.BinaryStateGPU(frames[i], frames[(i + 1) % 3]) for i in range(3)]
    one = [e.update_inner_product() for e in edges]
    total, two = multiframe.update_edges(edges)
    print(f"edges: per-edge nnz {one}, batched nnz {two} (total {total})", flush=True)
    for f in frames:
        f.release()
    g.close()


if __name__ == "__main__":
    only = os.environ.get("SANITIZE_MODES")
    for mode, persist in (("dense", "1"), ("grid", "1"), ("grid", "0"), ("tile", "1"), ("tile", "0"), ("brute", "1")):
        if only and mode not in only.split(","):
            continue
        for colour in (False, True):
            run_mode(mode, persist, colour)
    run_edges()
    print("sanitize cases done", flush=True)
