"""Python face of the CPU oracle (oracle/cvo_oracle.c).

TEST INFRASTRUCTURE ONLY (what is pinned against the reference and what is not: oracle/cvo_oracle.h).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (unified_cvo_b200) never does - and this package never imports
the product: its ABI structs (abi_types.py), parameter defaults and YAML reader (params.py) are
its own, so the reference arm of bench.py runs without libcvo_b200.so in the process.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .abi_types import AlignInfo, IterTrace, Params  # the oracle's own mirrors of the ABI structs
from .params import default_params, read_params_yaml  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcvo_oracle.so")
_BASE_SO = os.path.join(_HERE, "_build", "libcvo_cpu_baseline.so")


def build(force: bool = False) -> None:
    """gcc-compile the C restatement (no reference sources involved)."""
    if force:
        subprocess.run(["make", "-C", _HERE, "clean"], check=True, capture_output=True)
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)


class _Cloud(C.Structure):
    _fields_ = [
        ("n", C.c_int),
        ("F", C.c_int),
        ("C", C.c_int),
        ("xyz", C.POINTER(C.c_float)),
        ("feat", C.POINTER(C.c_float)),
        ("labels", C.POINTER(C.c_float)),
        ("geotype", C.POINTER(C.c_float)),
    ]


class _Sparse(C.Structure):
    _fields_ = [
        ("rows", C.c_int),
        ("stride", C.c_int),
        ("mat", C.POINTER(C.c_float)),
        ("ind", C.POINTER(C.c_int)),
        ("nonzeros", C.POINTER(C.c_uint)),
        ("capacity", C.c_int),
        ("nonzero_sum", C.c_ulonglong),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        f32p = C.POINTER(C.c_float)
        L.oracle_sparse_new.restype = C.POINTER(_Sparse)
        L.oracle_sparse_new.argtypes = [C.c_int, C.c_int]
        L.oracle_sparse_free.argtypes = [C.POINTER(_Sparse)]
        L.oracle_align.restype = C.c_int
        L.oracle_align.argtypes = [C.c_void_p, C.POINTER(_Cloud), C.POINTER(_Cloud), f32p,
                                   f32p, C.POINTER(AlignInfo), C.POINTER(IterTrace), C.c_int]
        L.oracle_iterate.restype = None
        L.oracle_iterate.argtypes = [C.c_void_p, C.POINTER(_Cloud), C.POINTER(_Cloud), f32p,
                                     f32p, C.c_float, C.c_int, C.POINTER(IterTrace),
                                     C.POINTER(_Sparse)]
        L.oracle_inner_product.restype = C.c_float
        L.oracle_inner_product.argtypes = [C.c_void_p, C.POINTER(_Cloud), C.POINTER(_Cloud),
                                           f32p, C.c_float, f32p, C.POINTER(_Sparse)]
        L.oracle_function_angle.restype = C.c_float
        L.oracle_function_angle.argtypes = [C.c_void_p, C.POINTER(_Cloud),
                                            C.POINTER(_Cloud), f32p, C.c_float, C.c_int]
        L.oracle_cubic_roots.restype = C.c_int
        L.oracle_cubic_roots.argtypes = [C.POINTER(C.c_double)] * 3
        L.oracle_exp_sek3.restype = None
        L.oracle_exp_sek3.argtypes = [f32p, C.c_float, f32p]
        L.oracle_indicator_sequence.restype = None
        L.oracle_indicator_sequence.argtypes = [C.c_void_p, C.c_int, f32p, C.POINTER(C.c_int), f32p, f32p]
        L.oracle_se3_log_norm.restype = C.c_double
        L.oracle_se3_log_norm.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.oracle_update_tf.restype = None
        L.oracle_update_tf.argtypes = [f32p] * 5
        L.oracle_transform.restype = None
        L.oracle_transform.argtypes = [f32p, f32p, f32p, C.c_int, f32p]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_num_threads.restype = None
        L.oracle_set_num_threads.argtypes = [C.c_int]
        L.oracle_transform_pose_vec.restype = None
        L.oracle_transform_pose_vec.argtypes = [f32p, f32p, C.c_int, f32p]
        L.oracle_edge_update.restype = C.c_ulonglong
        L.oracle_edge_update.argtypes = [C.c_void_p, C.POINTER(_Cloud), f32p,
                                         C.POINTER(_Cloud), f32p, C.c_float, C.c_int,
                                         C.POINTER(_Sparse)]
        f64p = C.POINTER(C.c_double)
        L.oracle_flow_rows.restype = None
        L.oracle_flow_rows.argtypes = [C.c_void_p, C.POINTER(_Cloud), f32p, C.POINTER(_Sparse),
                                       f64p, f64p]
        L.oracle_step_rows.restype = None
        L.oracle_step_rows.argtypes = [C.c_void_p, C.POINTER(_Cloud), f32p, C.c_int,
                                       C.POINTER(_Sparse), f32p, f32p, C.c_float, f64p]
        L.oracle_set_device_arith.restype = None
        L.oracle_set_device_arith.argtypes = [C.c_int]
        L.oracle_device_arith.restype = C.c_int
        L.oracle_fill_A.restype = None
        L.oracle_fill_A.argtypes = [C.c_void_p, C.POINTER(_Cloud), C.POINTER(_Cloud), f32p,
                                    C.c_int, C.c_float, C.POINTER(_Sparse)]
        L.oracle_fill_A_dense_kernel.restype = None
        L.oracle_fill_A_dense_kernel.argtypes = [C.c_void_p, C.POINTER(_Cloud),
                                                 C.POINTER(_Cloud), f32p, C.c_int, f32p,
                                                 C.POINTER(_Sparse)]
        _lib = L
    return _lib


def _fp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


class Cloud:
    """Host cloud in the C-ABI's layout (row-major features / labels)."""

    def __init__(self, xyz, features=None, labels=None, geotype=None):
        self.xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        n = self.xyz.shape[0]
        self.features = (None if features is None or np.size(features) == 0
                         else np.ascontiguousarray(features, dtype=np.float32).reshape(n, -1))
        self.labels = (None if labels is None or np.size(labels) == 0
                       else np.ascontiguousarray(labels, dtype=np.float32).reshape(n, -1))
        self.geotype = (None if geotype is None
                        else np.ascontiguousarray(geotype, dtype=np.float32).reshape(n, 2))

    @property
    def n(self):
        return self.xyz.shape[0]

    @property
    def F(self):
        return 0 if self.features is None else self.features.shape[1]

    @property
    def C(self):
        return 0 if self.labels is None else self.labels.shape[1]

    def c_struct(self) -> _Cloud:
        return _Cloud(self.n, self.F, self.C, _fp(self.xyz), _fp(self.features),
                      _fp(self.labels), _fp(self.geotype))


def _sparse_to_numpy(sp) -> dict:
    s = sp.contents
    rows, stride = s.rows, s.stride
    nz = np.ctypeslib.as_array(s.nonzeros, shape=(max(rows, 1),))[:rows].copy()
    if rows * stride > 0:
        mat = np.ctypeslib.as_array(s.mat, shape=(rows * stride,)).reshape(rows, stride).copy()
        ind = np.ctypeslib.as_array(s.ind, shape=(rows * stride,)).reshape(rows, stride).copy()
    else:
        mat = np.zeros((rows, 0), np.float32)
        ind = np.zeros((rows, 0), np.int32)
    return {"nonzeros": nz, "mat": mat, "ind": ind, "stride": stride,
            "nonzero_sum": int(s.nonzero_sum)}


def sparse_to_csr(sp: dict):
    """(row_ptr, cols, vals) in the reference's row order."""
    nz = sp["nonzeros"].astype(np.int64)
    row_ptr = np.zeros(len(nz) + 1, np.int64)
    np.cumsum(nz, out=row_ptr[1:])
    cols = np.concatenate([sp["ind"][i, : nz[i]] for i in range(len(nz))] or [np.zeros(0, np.int32)])
    vals = np.concatenate([sp["mat"][i, : nz[i]] for i in range(len(nz))] or [np.zeros(0, np.float32)])
    return row_ptr, cols.astype(np.int32), vals.astype(np.float32)


def iterate(params: Params, src: Cloud, tgt: Cloud, R, T, ell: float, num_neighbors: int,
            want_matrix: bool = False):
    L = lib()
    R = np.ascontiguousarray(R, np.float32).reshape(9)
    T = np.ascontiguousarray(T, np.float32).reshape(3)
    tr = IterTrace()
    cs, ct = src.c_struct(), tgt.c_struct()
    sp = L.oracle_sparse_new(src.n, max(int(num_neighbors), 1)) if want_matrix else None
    L.oracle_iterate(C.byref(params), C.byref(cs), C.byref(ct), _fp(R), _fp(T), C.c_float(ell),
                     int(num_neighbors), C.byref(tr), sp)
    if want_matrix:
        out = _sparse_to_numpy(sp)
        L.oracle_sparse_free(sp)
        return tr, out
    return tr


def align(params: Params, src: Cloud, tgt: Cloud, T_init=None, trace_cap: int = 0):
    L = lib()
    Ti = np.eye(4, dtype=np.float32) if T_init is None else np.asarray(T_init, np.float32)
    Ti = np.ascontiguousarray(Ti.T).reshape(16)  # column-major
    To = np.zeros(16, np.float32)
    info = AlignInfo()
    trace = (IterTrace * trace_cap)() if trace_cap > 0 else None
    cs, ct = src.c_struct(), tgt.c_struct()
    ret = L.oracle_align(C.byref(params), C.byref(cs), C.byref(ct), _fp(Ti), _fp(To),
                         C.byref(info), trace, trace_cap)
    executed = info.iterations + (0 if info.stop_reason == 8 else 1)  # 8 = STOP_MAX_ITER
    n_rec = min(trace_cap, executed) if trace_cap else 0
    return ret, To.reshape(4, 4).T.copy(), info, ([trace[i] for i in range(n_rec)] if trace else [])


def inner_product(params: Params, src: Cloud, tgt: Cloud, T, ell: float, kernel3x3=None,
                  want_matrix: bool = False):
    L = lib()
    T16 = np.ascontiguousarray(np.asarray(T, np.float32).T).reshape(16)
    K = None if kernel3x3 is None else np.ascontiguousarray(np.asarray(kernel3x3, np.float32).T).reshape(9)
    cs, ct = src.c_struct(), tgt.c_struct()
    sp = L.oracle_sparse_new(src.n, max(int(params.nearest_neighbors_max), 1)) if want_matrix else None
    val = L.oracle_inner_product(C.byref(params), C.byref(cs), C.byref(ct), _fp(T16),
                                 C.c_float(ell), _fp(K), sp)
    if want_matrix:
        out = _sparse_to_numpy(sp)
        L.oracle_sparse_free(sp)
        return float(val), out
    return float(val)


def transform_pose_vec(pose12, xyz):
    """x' = P [x 1]^T, P row-major 3x4 (CvoGPU_impl.cu:84-150)."""
    P = np.ascontiguousarray(pose12, np.float32).reshape(12)
    x = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    out = np.empty_like(x)
    lib().oracle_transform_pose_vec(_fp(P), _fp(x), len(x), _fp(out))
    return out


def edge_update(params: Params, f1: Cloud, pose1, f2: Cloud, pose2, ell: float,
                num_neighbors: int):
    """One edge update of the multi-frame IRLS (IRLS_State_GPU.cu:43-79 on frames moved by
    their own poses).  Returns (nonzero_sum, sparse dict)."""
    L = lib()
    P1 = np.ascontiguousarray(pose1, np.float32).reshape(12)
    P2 = np.ascontiguousarray(pose2, np.float32).reshape(12)
    c1, c2 = f1.c_struct(), f2.c_struct()
    sp = L.oracle_sparse_new(f1.n, max(int(num_neighbors), 1))
    total = L.oracle_edge_update(C.byref(params), C.byref(c1), _fp(P1), C.byref(c2), _fp(P2),
                                 C.c_float(ell), int(num_neighbors), sp)
    out = _sparse_to_numpy(sp)
    L.oracle_sparse_free(sp)
    return int(total), out


def function_angle(params: Params, src: Cloud, tgt: Cloud, T, ell: float, is_approximate=True):
    L = lib()
    T16 = np.ascontiguousarray(np.asarray(T, np.float32).T).reshape(16)
    cs, ct = src.c_struct(), tgt.c_struct()
    return float(L.oracle_function_angle(C.byref(params), C.byref(cs), C.byref(ct), _fp(T16),
                                         C.c_float(ell), int(bool(is_approximate))))


def cubic_roots(coef):
    L = lib()
    c = (C.c_double * 4)(*[float(x) for x in coef])
    re = (C.c_double * 3)()
    im = (C.c_double * 3)()
    rc = L.oracle_cubic_roots(c, re, im)
    return rc, np.array(re[:]) + 1j * np.array(im[:])


def exp_sek3(xi, dt):
    L = lib()
    x = np.ascontiguousarray(xi, np.float32).reshape(6)
    out = np.zeros(12, np.float32)
    L.oracle_exp_sek3(_fp(x), C.c_float(dt), _fp(out))
    return out.reshape(4, 3).T.copy()  # 3x4


def indicator_sequence(params: Params, indicators):
    """A_sparsity_indicator_ell_update (CvoGPU.cu:1167-1285) over a sequence, queues empty at entry:
    (decisions, start sums, end sums) after every call."""
    L = lib()
    x = np.ascontiguousarray(indicators, np.float32)
    n = int(x.size)
    dec = np.zeros(n, np.int32)
    s0, s1 = np.zeros(n, np.float32), np.zeros(n, np.float32)
    L.oracle_indicator_sequence(C.byref(params), n, _fp(x), dec.ctypes.data_as(C.POINTER(C.c_int)), _fp(s0), _fp(s1))
    return dec, s0, s1


def se3_log_norm(dR, dT):
    L = lib()
    r = np.ascontiguousarray(np.asarray(dR, np.float64).T).reshape(9)
    t = np.ascontiguousarray(dT, np.float64).reshape(3)
    return float(L.oracle_se3_log_norm(r.ctypes.data_as(C.POINTER(C.c_double)),
                                       t.ctypes.data_as(C.POINTER(C.c_double))))


def set_device_arith(on: bool) -> None:
    """K1's float sums with (default) / without the FMA contractions of the reference's GPU build
    (see cvo_oracle.c above mul_add())."""
    lib().oracle_set_device_arith(1 if on else 0)


def device_arith() -> bool:
    return bool(lib().oracle_device_arith())


def fill_A(params: Params, src: Cloud, tgt: Cloud, y_moved, num_neighbors: int, ell: float,
           kernel_inv=None) -> dict:
    """fill_in_A_mat_gpu (kernel_inv None) / fill_in_A_mat_gpu_dense_mat_kernel on an already
    moved target; kernel_inv is a 3x3 (row-major numpy) inverse kernel."""
    L = lib()
    y = np.ascontiguousarray(y_moved, np.float32).reshape(-1, 3)
    cs, ct = src.c_struct(), tgt.c_struct()
    sp = L.oracle_sparse_new(src.n, max(int(num_neighbors), 1))
    if kernel_inv is None:
        L.oracle_fill_A(C.byref(params), C.byref(cs), C.byref(ct), _fp(y), int(num_neighbors),
                        C.c_float(ell), sp)
    else:
        K = np.ascontiguousarray(np.asarray(kernel_inv, np.float32).T).reshape(9)
        L.oracle_fill_A_dense_kernel(C.byref(params), C.byref(cs), C.byref(ct), _fp(y),
                                     int(num_neighbors), _fp(K), sp)
    out = _sparse_to_numpy(sp)
    L.oracle_sparse_free(sp)
    return out


def transform(R, T, y):
    """update_tf + transform_pointcloud_thrust (CvoGPU.cu:94-112, CvoGPU_impl.cu:31-82): the
    target moved by the INVERSE of the pose (R, T); R is a 3x3 numpy matrix (row-major)."""
    L = lib()
    Rc = np.ascontiguousarray(np.asarray(R, np.float32).T).reshape(9)  # column-major
    Tc = np.ascontiguousarray(T, np.float32).reshape(3)
    Rinv = np.zeros(9, np.float32)
    Tinv = np.zeros(3, np.float32)
    tf = np.zeros(16, np.float32)
    L.oracle_update_tf(_fp(Rc), _fp(Tc), _fp(Rinv), _fp(Tinv), _fp(tf))
    yy = np.ascontiguousarray(y, np.float32).reshape(-1, 3)
    out = np.empty_like(yy)
    L.oracle_transform(_fp(Rinv), _fp(Tinv), _fp(yy), len(yy), _fp(out))
    return out


def _sparse_from_dict(L, sp: dict):
    """An oracle_sparse holding the given matrix (stride = its column count)."""
    rows, k = sp["mat"].shape
    h = L.oracle_sparse_new(rows, max(k, 1))
    s = h.contents
    s.stride = k
    if rows * k:
        np.ctypeslib.as_array(s.mat, shape=(rows * k,))[:] = np.ascontiguousarray(sp["mat"], np.float32).reshape(-1)
        np.ctypeslib.as_array(s.ind, shape=(rows * k,))[:] = np.ascontiguousarray(sp["ind"], np.int32).reshape(-1)
    np.ctypeslib.as_array(s.nonzeros, shape=(max(rows, 1),))[:rows] = sp["nonzeros"]
    return h


def flow_rows(params: Params, src: Cloud, y_moved, sparse: dict):
    """Per-row (omega_i / c, v_i / d) of compute_flow_gpu_no_eigen, [N, 3] doubles each."""
    L = lib()
    y = np.ascontiguousarray(y_moved, np.float32).reshape(-1, 3)
    h = _sparse_from_dict(L, sparse)
    om = np.zeros((src.n, 3), np.float64)
    vv = np.zeros((src.n, 3), np.float64)
    d = C.POINTER(C.c_double)
    cs = src.c_struct()
    L.oracle_flow_rows(C.byref(params), C.byref(cs), _fp(y), h, om.ctypes.data_as(d), vv.ctypes.data_as(d))
    L.oracle_sparse_free(h)
    return om, vv


def step_rows(params: Params, src: Cloud, y_moved, sparse: dict, omega, v, ell: float):
    """Per-row (B_i, C_i, D_i, E_i) of compute_step_size_xi + _poly_coeff, [N, 4] doubles."""
    L = lib()
    y = np.ascontiguousarray(y_moved, np.float32).reshape(-1, 3)
    h = _sparse_from_dict(L, sparse)
    out = np.zeros((src.n, 4), np.float64)
    o = np.ascontiguousarray(omega, np.float32).reshape(3)
    vv = np.ascontiguousarray(v, np.float32).reshape(3)
    cs = src.c_struct()
    L.oracle_step_rows(C.byref(params), C.byref(cs), _fp(y), len(y), h, _fp(o), _fp(vv), C.c_float(ell),
                       out.ctypes.data_as(C.POINTER(C.c_double)))
    L.oracle_sparse_free(h)
    return out


class _BaselineInfo(C.Structure):
    _fields_ = [("ret", C.c_int), ("iterations", C.c_int), ("executed", C.c_int), ("threads", C.c_int),
                ("final_ell", C.c_float), ("nnz_last", C.c_longlong), ("pairs", C.c_ulonglong),
                ("seconds", C.c_double), ("t_transform", C.c_double), ("t_kdtree_build", C.c_double),
                ("t_se_kernel", C.c_double), ("t_flow", C.c_double), ("t_step", C.c_double),
                ("t_rest", C.c_double)]


_base_lib = None


def _baseline():
    global _base_lib
    if _base_lib is None:
        if not os.path.exists(_BASE_SO):
            build()
        L = C.CDLL(_BASE_SO)
        L.cpu_baseline_align.restype = C.c_int
        L.cpu_baseline_align.argtypes = [C.c_void_p, C.POINTER(_Cloud), C.POINTER(_Cloud),
                                         C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float),
                                         C.POINTER(_BaselineInfo)]
        L.oracle_set_num_threads.argtypes = [C.c_int]
        L.cpu_baseline_radius_counts.restype = None
        L.cpu_baseline_radius_counts.argtypes = [C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float), C.c_int,
                                                 C.c_float, C.POINTER(C.c_int)]
        _base_lib = L
    return _base_lib


def cpu_baseline_radius_counts(pts, queries, r2: float):
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
    q = np.ascontiguousarray(queries, np.float32).reshape(-1, 3)
    out = np.zeros(len(q), np.int32)
    _baseline().cpu_baseline_radius_counts(_fp(pts), len(pts), _fp(q), len(q), C.c_float(r2),
                                           out.ctypes.data_as(C.POINTER(C.c_int)))
    return out


def cpu_baseline_align(params: Params, src: Cloud, tgt: Cloud, T_init=None, use_semantics: bool = False,
                       threads: int | None = None):
    """The restated reference CPU path cvo::cvo::align (oracle/cvo_cpu_baseline.c, Cvo.cpp:885-1089):
    kd-tree rebuilt per iteration, no row cap, no normalisation.  Returns (ret, T 4x4, info dict)."""
    _base_lib = _baseline()
    if threads:
        _base_lib.oracle_set_num_threads(int(threads))
    Ti = None
    if T_init is not None:
        Ti = np.ascontiguousarray(np.asarray(T_init, np.float32).T).reshape(16)
    To = np.zeros(16, np.float32)
    info = _BaselineInfo()
    cs, ct = src.c_struct(), tgt.c_struct()
    ret = _base_lib.cpu_baseline_align(C.byref(params), C.byref(cs), C.byref(ct), _fp(Ti),
                                       int(bool(use_semantics)), _fp(To), C.byref(info))
    d = {k: getattr(info, k) for k, _ in _BaselineInfo._fields_}
    return int(ret), To.reshape(4, 4).T.copy(), d


def set_accel(on: bool) -> None:
    """Accelerated candidate enumeration on (default) / off (the literal dense loop)."""
    lib().oracle_set_accel(1 if on else 0)


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def use_all_host_threads() -> int:
    """OpenMP threads = the cores this process may run on, whatever OMP_NUM_THREADS says
    (torchrun sets it to 1 for its workers).  Returns the count."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().oracle_set_num_threads(int(n))
    return num_threads()
