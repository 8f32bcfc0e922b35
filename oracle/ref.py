"""ctypes face of oracle/_ref — the REFERENCE'S OWN kernels compiled by oracle/make_ref.py.

TEST INFRASTRUCTURE ONLY (same rule as the rest of oracle/): tests and bench.py's CPU legs may
use it, the product never does.  Three builds of the same reference text:
  "host"        g++ -ffp-contract=off                (runs anywhere)
  "cuda"        nvcc, the reference's own flags       (needs a GPU)
  "cuda_nofma"  nvcc --fmad=false                     (needs a GPU)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_HERE, "_ref")
_NAMES = {"host": "libcvo_ref_host.so", "cuda": "libcvo_ref_cuda.so",
          "cuda_nofma": "libcvo_ref_cuda_nofma.so"}
_libs: dict = {}

NUM_CLASSES = 19          # CMakeLists.txt:498 of the reference (compile definitions)
FEATURE_DIMENSIONS = 5


def available(kind: str = "host") -> bool:
    return os.path.exists(os.path.join(_DIR, _NAMES[kind]))


def lib(kind: str = "host") -> C.CDLL:
    if kind not in _libs:
        path = os.path.join(_DIR, _NAMES[kind])
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: run `python oracle/make_ref.py` where /root/reference exists")
        L = C.CDLL(path)
        f32p, i32p, u32p, f64p = (C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_uint),
                                  C.POINTER(C.c_double))
        L.cvo_ref_build_kind.restype = C.c_int
        L.cvo_ref_num_classes.restype = C.c_int
        L.cvo_ref_feature_dimensions.restype = C.c_int
        cloud = [C.c_int, f32p, f32p, f32p, f32p]
        L.cvo_ref_fill_A.restype = C.c_int
        L.cvo_ref_fill_A.argtypes = [C.c_void_p, *cloud, *cloud, C.c_int, C.c_int, C.c_int, C.c_float,
                                     f32p, i32p, u32p]
        L.cvo_ref_fill_A_dense.restype = C.c_int
        L.cvo_ref_fill_A_dense.argtypes = [C.c_void_p, *cloud, *cloud, C.c_int, C.c_int, C.c_int, f32p,
                                           f32p, i32p, u32p]
        L.cvo_ref_flow_rows.restype = C.c_int
        L.cvo_ref_flow_rows.argtypes = [C.c_void_p, C.c_int, f32p, C.c_int, f32p, C.c_int, f32p, i32p,
                                        f64p, f64p]
        L.cvo_ref_step_rows.restype = C.c_int
        L.cvo_ref_step_rows.argtypes = [f32p, f32p, C.c_float, C.c_float, C.c_int, C.c_int, f32p,
                                        C.c_int, f32p, C.c_int, f32p, i32p, f64p, f64p, f64p, f64p]
        if kind == "host":  # host code of the reference (LieGroup.cpp): host build only
            L.cvo_ref_indicator_sequence.restype = C.c_int
            L.cvo_ref_indicator_sequence.argtypes = [C.c_void_p, C.c_int, f32p, i32p, f32p, f32p]
            L.cvo_ref_update_tf_and_transform.restype = C.c_int
            L.cvo_ref_update_tf_and_transform.argtypes = [f32p, f32p, f32p, f32p, f32p, C.c_int, f32p, f32p]
            L.cvo_ref_transform_pose_vec.restype = C.c_int
            L.cvo_ref_transform_pose_vec.argtypes = [f32p, C.c_int, f32p, f32p]
            L.cvo_ref_exp_sek3.restype = C.c_int
            L.cvo_ref_exp_sek3.argtypes = [f32p, C.c_float, f32p]
        assert L.cvo_ref_num_classes() == NUM_CLASSES and L.cvo_ref_feature_dimensions() == FEATURE_DIMENSIONS
        _libs[kind] = L
    return _libs[kind]


def _f(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def _cloud_args(xyz, feat, lab, geo, F, Cc):
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    n = len(xyz)
    keep = [xyz]
    f = l = g = None
    if F:
        f = np.ascontiguousarray(feat, np.float32).reshape(n, F)
        keep.append(f)
    if Cc:
        l = np.ascontiguousarray(lab, np.float32).reshape(n, Cc)
        keep.append(l)
    if geo is not None:
        g = np.ascontiguousarray(geo, np.float32).reshape(n, 2)
        keep.append(g)
    return [n, _f(xyz), _f(f), _f(l), _f(g)], keep


def _dims(src, tgt):
    F = max(src.F, tgt.F)
    Cc = max(src.C, tgt.C)
    if (src.F and tgt.F and src.F != tgt.F) or (src.C and tgt.C and src.C != tgt.C):
        raise ValueError("reference kernels need equal feature / class widths on both clouds")
    if F > FEATURE_DIMENSIONS or Cc > NUM_CLASSES:
        raise ValueError("the reference is compiled for FEATURE_DIMENSIONS=5, NUM_CLASSES=19")
    return F, Cc


def _widen(a, n, w):
    """a cloud without a channel = zeros (what the reference's CvoPoint ctor leaves)."""
    return np.zeros((n, w), np.float32) if a is None and w else a


def fill_A(params, src, tgt, y_moved, num_neighbors: int, ell: float, kind: str = "host",
           kernel_inv=None) -> dict:
    """The reference's fill_in_A_mat_gpu (CvoGPU.cu:477-593; kernel_inv: the dense-kernel variant
    :217-327) on oracle.Cloud objects; y_moved replaces the target's xyz.  Returns the same dict
    as oracle.fill_A: nonzeros[N], mat[N, k], ind[N, k] (-1 = unused)."""
    L = lib(kind)
    F, Cc = _dims(src, tgt)
    a_args, k1 = _cloud_args(src.xyz, _widen(src.features, src.n, F), _widen(src.labels, src.n, Cc),
                             src.geotype, F, Cc)
    b_args, k2 = _cloud_args(y_moved, _widen(tgt.features, tgt.n, F), _widen(tgt.labels, tgt.n, Cc),
                             tgt.geotype, F, Cc)
    k = int(num_neighbors)
    n = src.n
    mat = np.zeros((n, max(k, 1)), np.float32)
    ind = np.full((n, max(k, 1)), -1, np.int32)
    nz = np.zeros(n, np.uint32)
    outs = [_f(mat), ind.ctypes.data_as(C.POINTER(C.c_int)), nz.ctypes.data_as(C.POINTER(C.c_uint))]
    if kernel_inv is None:
        rc = L.cvo_ref_fill_A(C.addressof(params), *a_args, *b_args, F, Cc, k, C.c_float(ell), *outs)
    else:
        K = np.ascontiguousarray(np.asarray(kernel_inv, np.float32).T).reshape(9)  # column-major
        rc = L.cvo_ref_fill_A_dense(C.addressof(params), *a_args, *b_args, F, Cc, k, _f(K), *outs)
    if rc != 0:
        raise RuntimeError(f"cvo_ref_fill_A failed: {rc}")
    del k1, k2
    return {"nonzeros": nz, "mat": mat[:, :k], "ind": ind[:, :k], "stride": k,
            "nonzero_sum": int(nz.astype(np.int64).sum())}


def flow_rows(params, src_xyz, y_moved, sparse: dict, kind: str = "host"):
    """compute_flow_gpu_no_eigen (CvoGPU.cu:729-790): per-row (omega_i / c, v_i / d) as doubles."""
    L = lib(kind)
    x = np.ascontiguousarray(src_xyz, np.float32).reshape(-1, 3)
    y = np.ascontiguousarray(y_moved, np.float32).reshape(-1, 3)
    mat = np.ascontiguousarray(sparse["mat"], np.float32)
    ind = np.ascontiguousarray(sparse["ind"], np.int32)
    k = mat.shape[1]
    om = np.zeros((len(x), 3), np.float64)
    vv = np.zeros((len(x), 3), np.float64)
    d = C.POINTER(C.c_double)
    rc = L.cvo_ref_flow_rows(C.addressof(params), len(x), _f(x), len(y), _f(y), k, _f(mat),
                             ind.ctypes.data_as(C.POINTER(C.c_int)), om.ctypes.data_as(d),
                             vv.ctypes.data_as(d))
    if rc != 0:
        raise RuntimeError(f"cvo_ref_flow_rows failed: {rc}")
    return om, vv


def step_rows(omega, v, ell: float, ell_init: float, is_using_range_ell: int, src_xyz, y_moved,
              sparse: dict, kind: str = "host"):
    """compute_step_size_xi + compute_step_size_poly_coeff (CvoGPU.cu:953-1082): per-row
    (B_i, C_i, D_i, E_i) as an [N, 4] double array."""
    L = lib(kind)
    x = np.ascontiguousarray(src_xyz, np.float32).reshape(-1, 3)
    y = np.ascontiguousarray(y_moved, np.float32).reshape(-1, 3)
    mat = np.ascontiguousarray(sparse["mat"], np.float32)
    ind = np.ascontiguousarray(sparse["ind"], np.int32)
    k = mat.shape[1]
    o = np.ascontiguousarray(omega, np.float32).reshape(3)
    vv = np.ascontiguousarray(v, np.float32).reshape(3)
    outs = [np.zeros(len(x), np.float64) for _ in range(4)]
    d = C.POINTER(C.c_double)
    rc = L.cvo_ref_step_rows(_f(o), _f(vv), C.c_float(ell), C.c_float(ell_init), int(is_using_range_ell),
                             len(x), _f(x), len(y), _f(y), k, _f(mat),
                             ind.ctypes.data_as(C.POINTER(C.c_int)), *[a.ctypes.data_as(d) for a in outs])
    if rc != 0:
        raise RuntimeError(f"cvo_ref_step_rows failed: {rc}")
    return np.stack(outs, axis=1)


def exp_sek3(xi, dt: float):
    """The reference's Exp_SEK3(v, dt) (LieGroup.cpp:245-274, tier 2: over oracle/ref_mini_eigen.h),
    3x4 [R | t]."""
    x = np.ascontiguousarray(xi, np.float32).reshape(6)
    out = np.zeros(12, np.float32)
    rc = lib("host").cvo_ref_exp_sek3(_f(x), C.c_float(dt), _f(out))
    assert rc == 0
    return out.reshape(4, 3).T.copy()


def indicator_sequence(params, indicators):
    """The reference's A_sparsity_indicator_ell_update (CvoGPU.cu:1167-1285, tier 1: std::queue
    only) fed with a sequence of indicators, queues empty at entry: (decisions, start sums, end sums)."""
    x = np.ascontiguousarray(indicators, np.float32)
    n = int(x.size)
    dec = np.zeros(n, np.int32)
    s0, s1 = np.zeros(n, np.float32), np.zeros(n, np.float32)
    rc = lib("host").cvo_ref_indicator_sequence(C.byref(params), n, _f(x), dec.ctypes.data_as(C.POINTER(C.c_int)),
                                                _f(s0), _f(s1))
    assert rc == 0
    return dec, s0, s1


def update_tf_and_transform(R, T, xyz):
    """The reference's update_tf (CvoGPU.cu:94-112) and, with the inverse pose it uploads, its point
    transform transform_point_R_T (CvoGPU_impl.cu:31-82) on xyz (tier 2: over the mini-Eigen).
    R 3x3, T 3; returns (Rinv 3x3, Tinv 3, transform 4x4, moved n x 3)."""
    r = np.ascontiguousarray(np.asarray(R, np.float32).T).reshape(9)  # column-major
    t = np.ascontiguousarray(T, np.float32).reshape(3)
    y = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    rinv, tinv, tf = np.zeros(9, np.float32), np.zeros(3, np.float32), np.zeros(16, np.float32)
    out = np.zeros_like(y)
    rc = lib("host").cvo_ref_update_tf_and_transform(_f(r), _f(t), _f(rinv), _f(tinv), _f(tf), int(y.shape[0]), _f(y), _f(out))
    assert rc == 0
    return rinv.reshape(3, 3).T.copy(), tinv, tf.reshape(4, 4).T.copy(), out


def transform_pose_vec(pose12, xyz):
    """The reference's transform_point_pose_vec (CvoGPU_impl.cu:84-150, tier 2: over the mini-Eigen):
    x' = P [x 1]^T with P the frame's row-major 3x4 pose."""
    P = np.ascontiguousarray(pose12, np.float32).reshape(12)
    x = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    out = np.zeros_like(x)
    rc = lib("host").cvo_ref_transform_pose_vec(_f(P), int(x.shape[0]), _f(x), _f(out))
    assert rc == 0
    return out
