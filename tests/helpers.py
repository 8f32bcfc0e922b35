"""Shared helpers of the parity tests: named workloads and trace comparison."""
import os

import numpy as np

import oracle
import unified_cvo_b200 as u
from unified_cvo_b200 import synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "tests", "data")


def to_oracle_cloud(pc: u.CvoPointCloud) -> oracle.Cloud:
    return oracle.Cloud(pc.positions_, pc.features_, pc.labels_, pc.geometric_types_)


def cloud_from_dict(d) -> u.CvoPointCloud:
    return u.CvoPointCloud(d["xyz"], d["features"], d["labels"], d["geotype"])


def geometric_params(ell_init=0.95):
    """SURVEY.md §8(d) C2 parameters: cvo_outdoor_params.yaml, geometry only."""
    p = u.read_params_yaml(os.path.join(DATA, "cvo_outdoor_params.yaml"))
    p.is_using_intensity = 0
    p.is_using_geometric_type = 0
    p.ell_init = ell_init
    return p


def demo_clouds(color=True):
    src = u.CvoPointCloud.from_pcd(os.path.join(DATA, "source.pcd"), use_color=color)
    tgt = u.CvoPointCloud.from_pcd(os.path.join(DATA, "target.pcd"), use_color=color)
    return src, tgt


def demo_params(src, tgt, color=True):
    """What main_cvo_gpu_align_two_color_pcd.cpp:56-66 does to the yaml parameters."""
    p = u.read_params_yaml(os.path.join(DATA, "cvo_outdoor_params.yaml"))
    p.ell_init = float(np.linalg.norm(src.positions_.mean(0, dtype=np.float32)
                                      - tgt.positions_.mean(0, dtype=np.float32)))
    p.ell_decay_rate = p.ell_decay_rate_first_frame
    p.ell_decay_start = p.ell_decay_start_first_frame
    if not color:
        p.is_using_intensity = 0  # main_cvo_gpu_align_two_pcd.cpp:66
    return p


def synthetic_pair(P, N, M, seed, F=0, C=0, geotype=False):
    d = synthetic.make_pair(P, N, M, seed, F=F, C=C, with_geotype=geotype)
    return cloud_from_dict(d["source"]), cloud_from_dict(d["target"]), d["T_gt"]


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def compare_traces(got, ref, twist_tol=1e-4, scalar_tol=1e-6):
    """Per-iteration parity of two IterTrace records.  Returns a list of mismatch strings."""
    bad = []
    if got.nnz != ref.nnz:
        bad.append(f"nnz {got.nnz} != {ref.nnz}")
    if got.max_row_nnz != ref.max_row_nnz:
        bad.append(f"max_row_nnz {got.max_row_nnz} != {ref.max_row_nnz}")
    twist_g = list(got.omega) + list(got.v)
    twist_r = list(ref.omega) + list(ref.v)
    e = rel_err(twist_g, twist_r)
    if not e <= twist_tol:
        bad.append(f"twist rel err {e:.3e}")
    for name in ("omega_sum", "v_sum"):
        e = rel_err(list(getattr(got, name)), list(getattr(ref, name)))
        scale = np.linalg.norm(list(ref.omega_sum) + list(ref.v_sum))
        ea = np.linalg.norm(np.array(list(getattr(got, name))) - np.array(list(getattr(ref, name))))
        # float row sums in a different order (8 lanes per row, Morton-ordered candidates)
        if not ea <= 3e-5 * max(scale, 1e-30):
            bad.append(f"{name} abs err {ea:.3e} (scale {scale:.3e})")
    bcde_g = np.array([got.B, got.C, got.D, got.E])
    bcde_r = np.array([ref.B, ref.C, ref.D, ref.E])
    for n_, g_, r_ in zip("BCDE", bcde_g, bcde_r):
        if not abs(g_ - r_) <= 1e-4 * max(abs(r_), 1e-12) + 1e-9 * np.abs(bcde_r).max():
            bad.append(f"{n_} {g_:.9e} != {r_:.9e}")
    if not abs(got.step - ref.step) <= 1e-4 * abs(ref.step) + 1e-12:
        bad.append(f"step {got.step:.9e} != {ref.step:.9e}")
    if not abs(got.a_sum - ref.a_sum) <= scalar_tol * max(abs(ref.a_sum), 1e-30):
        bad.append(f"a_sum {got.a_sum:.9e} != {ref.a_sum:.9e}")
    return bad
