"""CvoParams defaults and the YAML reader, restated for the ORACLE side (pure Python).

TEST INFRASTRUCTURE.  Follows cvo::CvoParams::CvoParams() (include/UnifiedCvo/cvo/CvoParams.hpp:75-126)
and read_CvoParams_yaml (:193-303): every key optional, only the keys the reference looks up are
read, `max_step` / `step` have no default (0 here).  Duplicate keys: the FIRST occurrence wins
(yaml-cpp's map lookup returns the first equal key - a belief about yaml-cpp 0.7, un-vendored and
absent here; DESIGN.md lists it as unpinned).  tests/test_abi.py holds this reader to the
product's C reader on every shipped yaml.
"""
from __future__ import annotations

from .abi_types import Params

# CvoParams.hpp:75-126
DEFAULTS = dict(
    ell_init_first_frame=0.5, ell_init=0.5, ell_min=0.05, min_ell_iter_limit=1, ell_max=1.2, dl=0.0, dl_step=0.3,
    sigma=0.1, sp_thres=0.0006, c=7.0, d=7.0, c_ell=0.15, c_sigma=0.6, s_ell=0.1, s_sigma=0.8, MAX_ITER=10000,
    min_step=2e-5, eps=0.00005, eps_2=0.000012, max_step=0.0, step=0.0, ell_decay_rate=0.9,
    ell_decay_rate_first_frame=0.99, ell_decay_start=30, ell_decay_start_first_frame=300, indicator_window_size=15,
    indicator_stable_threshold=0.2, is_pcl_visualization_on=0, is_using_least_square=0, is_ell_adaptive=0,
    is_full_ip_matrix=0, is_using_geometry=1, is_using_intensity=0, is_using_semantics=0, is_using_range_ell=0,
    is_using_kdtree=0, is_using_geometric_type=0, is_exporting_association=0, multiframe_using_cpu=1,
    multiframe_max_iters=200, nearest_neighbors_max=512, multiframe_ell_init=0.15, multiframe_ell_min=0.05,
    multiframe_iter_per_ell=10, multiframe_ell_decay_rate=0.7, multiframe_iterations_per_ell=50,
    multiframe_iterations_per_solve=8, multiframe_downsample_voxel_size=0.5, multiframe_expected_points=1000,
    multiframe_num_neighbors=128, multiframe_min_nonzeros=300, multiframe_least_squares_num_threads=24)

# the keys read_CvoParams_yaml looks up (CvoParams.hpp:197-296)
YAML_KEYS = (
    "ell_init_first_frame ell_init ell_min min_ell_iter_limit ell_max dl dl_step sigma sp_thres c d c_ell c_sigma "
    "s_ell s_sigma MAX_ITER eps eps_2 min_step max_step ell_decay_rate ell_decay_rate_first_frame ell_decay_start "
    "ell_decay_start_first_frame indicator_window_size indicator_stable_threshold is_pcl_visualization_on "
    "is_using_least_square is_full_ip_matrix is_using_geometry is_using_intensity is_using_semantics "
    "is_using_range_ell is_using_kdtree is_using_geometric_type is_exporting_association nearest_neighbors_max "
    "multiframe_using_cpu multiframe_ell_init multiframe_max_iters multiframe_ell_min multiframe_ell_decay_rate "
    "multiframe_iterations_per_ell multiframe_iterations_per_solve multiframe_downsample_voxel_size "
    "multiframe_expected_points multiframe_num_neighbors multiframe_min_nonzeros "
    "multiframe_least_squares_num_threads").split()


def default_params() -> Params:
    p = Params()
    for k, v in DEFAULTS.items():
        setattr(p, k, v)
    return p


def read_params_yaml(path: str) -> Params:
    p = default_params()
    types = dict(Params._fields_)
    seen = set()
    with open(path) as fh:
        for line in fh:
            line = line.split("#", 1)[0].strip()
            if not line or line.startswith("%") or line.startswith("---") or ":" not in line:
                continue
            key, val = (x.strip() for x in line.split(":", 1))
            if not val or key not in YAML_KEYS or key in seen:
                continue
            seen.add(key)
            v = float(val)
            import ctypes as C
            setattr(p, key, int(v) if types[key] is C.c_int else v)
    return p
