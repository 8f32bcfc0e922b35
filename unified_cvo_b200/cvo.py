"""Host-side mirror of the reference's registration API over the C-ABI.

Same names, argument meaning and error behaviour as
  cvo::CvoGPU          include/UnifiedCvo/cvo/CvoGPU.hpp:33-232
  cvo::CvoPointCloud   include/UnifiedCvo/utils/CvoPointCloud.hpp:65-173 (in-memory subset)
  cvo::CvoParams       include/UnifiedCvo/cvo/CvoParams.hpp:12-128
  cvo::Association     include/UnifiedCvo/cvo/Association.hpp:7-11
so the parity tests read like the reference's drivers
(src/experiments/main_cvo_gpu_align_two_color_pcd.cpp:36-105).

Everything numeric happens inside libcvo_b200.so; this file only marshals numpy arrays.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _abi
from ._abi import AlignInfo, IterTrace, Params

CvoParams = Params


class CvoError(RuntimeError):
    pass


def default_params() -> Params:
    p = Params()
    _abi.load_library().cvo_b200_params_default(C.byref(p))
    return p


def read_params_yaml(path: str) -> Params:
    """read_CvoParams_yaml (CvoParams.hpp:193-303): defaults, then the keys present."""
    lib = _abi.load_library()
    p = Params()
    lib.cvo_b200_params_default(C.byref(p))
    rc = lib.cvo_b200_params_read_yaml(str(path).encode(), C.byref(p))
    if rc != _abi.OK:
        raise CvoError(f"cannot read parameter file {path!r} (code {rc})")
    return p


def _f32(a, shape=None):
    if a is None:
        return None
    out = np.ascontiguousarray(a, dtype=np.float32)
    return out if shape is None else out.reshape(shape)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def _colmajor16(T) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(T, dtype=np.float32).reshape(4, 4).T).reshape(16)


class CvoPointCloud:
    """positions (N,3), features (N,F), labels (N,C), geometric_types (N,2) — all float32."""

    def __init__(self, positions=None, features=None, labels=None, geometric_types=None):
        self.positions_ = (np.zeros((0, 3), np.float32) if positions is None
                           else _f32(positions).reshape(-1, 3))
        n = self.positions_.shape[0]
        self.features_ = None if features is None or np.size(features) == 0 else _f32(features).reshape(n, -1)
        self.labels_ = None if labels is None or np.size(labels) == 0 else _f32(labels).reshape(n, -1)
        self.geometric_types_ = None if geometric_types is None else _f32(geometric_types).reshape(n, 2)

    # --- accessors (CvoPointCloud.hpp:125-146)
    def num_points(self) -> int:
        return int(self.positions_.shape[0])

    def num_classes(self) -> int:
        return 0 if self.labels_ is None else int(self.labels_.shape[1])

    def feature_dimensions(self) -> int:
        return 0 if self.features_ is None else int(self.features_.shape[1])

    def positions(self):
        return self.positions_

    def features(self):
        return self.features_

    def labels(self):
        return self.labels_

    def geometric_types(self):
        return self.geometric_types_

    # --- constructors the demo drivers use
    @classmethod
    def from_xyzrgb(cls, xyz, rgb_uint8) -> "CvoPointCloud":
        """CvoPointCloud(pcl::PointCloud<PointXYZRGB>) (CvoPointCloud.cpp:570-594):
        features = (r,g,b)/255, 0, 0; geometric type (0,1)."""
        xyz = _f32(xyz).reshape(-1, 3)
        n = xyz.shape[0]
        f = np.zeros((n, 5), np.float32)
        rgb = np.asarray(rgb_uint8).reshape(n, 3).astype(np.int32).astype(np.float32)
        f[:, :3] = (rgb.astype(np.float64) / 255.0).astype(np.float32)
        g = np.zeros((n, 2), np.float32)
        g[:, 1] = 1.0
        return cls(xyz, f, None, g)

    @classmethod
    def from_xyz(cls, xyz) -> "CvoPointCloud":
        """CvoPointCloud(pcl::PointCloud<PointXYZ>) (CvoPointCloud.cpp:634-652): no features,
        geometric type (1,0)."""
        xyz = _f32(xyz).reshape(-1, 3)
        g = np.zeros((xyz.shape[0], 2), np.float32)
        g[:, 0] = 1.0
        return cls(xyz, None, None, g)

    @classmethod
    def from_pcd(cls, path: str, use_color: bool = True) -> "CvoPointCloud":
        """ASCII PCD with FIELDS x y z [rgb] (the demo_data flavour)."""
        fields, n, rows = None, None, []
        with open(path, "r") as fh:
            data = False
            for line in fh:
                if data:
                    if line.strip():
                        rows.append(line.split())
                    continue
                tok = line.split()
                if not tok or tok[0].startswith("#"):
                    continue
                if tok[0] == "FIELDS":
                    fields = tok[1:]
                elif tok[0] == "POINTS":
                    n = int(tok[1])
                elif tok[0] == "DATA":
                    if tok[1] != "ascii":
                        raise CvoError("only ASCII PCD files are supported")
                    data = True
        if fields is None or n is None or len(rows) != n:
            raise CvoError(f"malformed PCD file {path!r}")
        ix, iy, iz = fields.index("x"), fields.index("y"), fields.index("z")
        xyz = np.array([[float(r[ix]), float(r[iy]), float(r[iz])] for r in rows], np.float32)
        if use_color and "rgb" in fields:
            ir = fields.index("rgb")
            packed = np.array([int(float(r[ir])) if "." in r[ir] or "e" in r[ir] else int(r[ir])
                               for r in rows], np.uint64).astype(np.uint32)
            rgb = np.stack([(packed >> 16) & 255, (packed >> 8) & 255, packed & 255], axis=1)
            return cls.from_xyzrgb(xyz, rgb.astype(np.uint8))
        return cls.from_xyz(xyz)

    # --- CvoPointCloud::transform (CvoPointCloud.cpp:1366-1381) and operator+
    @staticmethod
    def transform(pose, inp: "CvoPointCloud") -> "CvoPointCloud":
        T = np.asarray(pose, np.float32).reshape(4, 4)
        p = inp.positions_ @ T[:3, :3].T + T[:3, 3]
        return CvoPointCloud(p.astype(np.float32), inp.features_, inp.labels_, inp.geometric_types_)

    def __add__(self, other: "CvoPointCloud") -> "CvoPointCloud":
        def cat(a, b):
            return None if a is None or b is None else np.concatenate([a, b], axis=0)
        return CvoPointCloud(np.concatenate([self.positions_, other.positions_], axis=0),
                             cat(self.features_, other.features_), cat(self.labels_, other.labels_),
                             cat(self.geometric_types_, other.geometric_types_))


@dataclass
class Association:
    """cvo::Association: inlier index lists + the sparse N x M weight matrix (CSR).  The index
    lists are views of the CSR (rows that hold an entry; every entry's column, in row order -
    gpu_association_to_cpu, CvoGPU_impl.cu:366-427) and are built when first read."""
    row_ptr: Optional[np.ndarray] = None
    cols: Optional[np.ndarray] = None
    vals: Optional[np.ndarray] = None
    shape: tuple = (0, 0)

    @property
    def source_inliers(self) -> list:
        return [] if self.row_ptr is None else np.nonzero(np.diff(self.row_ptr))[0].tolist()

    @property
    def target_inliers(self) -> list:
        return [] if self.cols is None else self.cols.tolist()

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.vals, self.cols, self.row_ptr), shape=self.shape)


class CvoGPU:
    """cvo::CvoGPU.  One instance = one device handle (one CUDA stream, one set of buffers)."""

    def __init__(self, param_file_or_params, device: int = 0):
        self._lib = _abi.load_library()
        if isinstance(param_file_or_params, Params):
            self.params = param_file_or_params.copy()
        else:
            self.params = read_params_yaml(param_file_or_params)
        h = C.c_void_p()
        rc = self._lib.cvo_b200_create(C.byref(self.params), int(device), C.byref(h))
        if rc != _abi.OK:
            raise CvoError(f"cvo_b200_create failed ({rc}): "
                           f"{self._lib.cvo_b200_global_error().decode(errors='replace')}")
        self._h = h
        self._src_id = self._tgt_id = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cvo_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- parameters (CvoGPU.hpp:52-54, CvoGPU.cu:73-77)
    def get_params(self) -> Params:
        return self.params

    def write_params(self, p: Optional[Params] = None):
        if p is not None and p is not self.params:
            self.params = p.copy()
        self._check(self._lib.cvo_b200_write_params(self._h, C.byref(self.params)))

    # --- helpers
    def _check(self, rc: int):
        if rc != _abi.OK:
            msg = self._lib.cvo_b200_last_error(self._h).decode(errors="replace")
            raise CvoError(f"cvo_b200 error {rc}: {msg}")

    def set_cloud(self, which: int, pc: CvoPointCloud):
        F, Cn = pc.feature_dimensions(), pc.num_classes()
        self._keep = getattr(self, "_keep", {})
        self._keep[which] = pc  # keep the arrays alive during the call
        self._check(self._lib.cvo_b200_set_cloud(
            self._h, which, pc.num_points(), _ptr(pc.positions_), F, _ptr(pc.features_), Cn,
            _ptr(pc.labels_), _ptr(pc.geometric_types_)))

    def set_row_range(self, begin: int, end: int):
        self._check(self._lib.cvo_b200_set_row_range(self._h, int(begin), int(end)))

    # --- the hot path
    def iterate(self, R, T, ell: float, num_neighbors: int) -> IterTrace:
        """One CVO iteration at an explicit state on the clouds already set."""
        Rc = np.ascontiguousarray(np.asarray(R, np.float32).reshape(3, 3).T).reshape(9)
        Tc = _f32(T).reshape(3)
        tr = IterTrace()
        self._check(self._lib.cvo_b200_iterate(self._h, _ptr(Rc), _ptr(Tc), C.c_float(ell),
                                               int(num_neighbors), C.byref(tr)))
        return tr

    def align(self, source: CvoPointCloud, target: CvoPointCloud, T_target_to_source=None,
              association: Optional[Association] = None, trace_cap: int = 0, resident: bool = False):
        """CvoGPU::align (CvoGPU.cu:1605-1632).  Returns (ret, transform 4x4, info[, trace]).

        resident=True skips the upload and registers the clouds already on the device."""
        Ti = np.eye(4, dtype=np.float32) if T_target_to_source is None else T_target_to_source
        Ti = _colmajor16(Ti)
        To = np.zeros(16, np.float32)
        To[[0, 5, 10, 15]] = 1.0
        info = AlignInfo()
        if source.num_points() == 0 or target.num_points() == 0:
            return (0, To.reshape(4, 4).T.copy(), info) + (([],) if trace_cap else ())
        self.write_params()
        if not resident:
            self.set_cloud(0, source)
            self.set_cloud(1, target)
        trace = (IterTrace * trace_cap)() if trace_cap > 0 else None
        self._check(self._lib.cvo_b200_align(self._h, _ptr(Ti), _ptr(To), C.byref(info), trace, trace_cap))
        Tm = To.reshape(4, 4).T.copy()
        if association is not None and self.params.is_exporting_association:
            # align_impl exports the LAST iteration's matrix (CvoGPU.cu:1552-1556)
            self._fill_association(association, source.num_points(), target.num_points(),
                                   lambda *a: self._lib.cvo_b200_align_association(self._h, *a))
        if trace_cap:
            executed = info.iterations + (0 if info.stop_reason == _abi.STOP_MAX_ITER else 1)
            n = min(trace_cap, executed)
            return info.ret, Tm, info, [trace[i] for i in range(n)]
        return info.ret, Tm, info

    def align_host(self, source: CvoPointCloud, target: CvoPointCloud, T_target_to_source=None):
        """The single C-ABI call a CvoGPU::align(const CvoPointCloud&, ...) maps to: host
        buffers in, pose out, uploads inside."""
        Ti = np.eye(4, dtype=np.float32) if T_target_to_source is None else T_target_to_source
        Ti = _colmajor16(Ti)
        To = np.zeros(16, np.float32)
        To[[0, 5, 10, 15]] = 1.0
        info = AlignInfo()
        if source.num_points() == 0 or target.num_points() == 0:  # CvoGPU.cu:1614-1617
            return 0, To.reshape(4, 4).T.copy(), info
        Fs, Ft = source.feature_dimensions(), target.feature_dimensions()
        Cs, Ct = source.num_classes(), target.num_classes()
        if (Fs and Ft and Fs != Ft) or (Cs and Ct and Cs != Ct):
            # one stride is passed for both clouds: unequal widths would be read out of bounds
            raise CvoError("source and target feature / class dimensions differ")
        F, Cn = max(Fs, Ft), max(Cs, Ct)
        self.write_params()
        self._check(self._lib.cvo_b200_align_host(
            self._h, source.num_points(), _ptr(source.positions_), F, _ptr(source.features_), Cn,
            _ptr(source.labels_), _ptr(source.geometric_types_), target.num_points(),
            _ptr(target.positions_), _ptr(target.features_), _ptr(target.labels_),
            _ptr(target.geometric_types_), _ptr(Ti), _ptr(To), C.byref(info)))
        return info.ret, To.reshape(4, 4).T.copy(), info

    def inner_product_gpu(self, source, target, T_target_to_source, ell: float) -> float:
        self.write_params()
        self.set_cloud(0, source)
        self.set_cloud(1, target)
        out = C.c_float(0.0)
        self._check(self._lib.cvo_b200_inner_product(self._h, _ptr(_colmajor16(T_target_to_source)),
                                                     C.c_float(ell), C.byref(out)))
        return float(out.value)

    def function_angle(self, source, target, T_target_to_source, ell: float,
                       is_approximate: bool = True, is_gpu: bool = True) -> float:
        """CvoGPU::function_angle (CvoGPU.cu:1814-1846)."""
        if not is_gpu:
            raise CvoError("the CPU variant (inner_product_cpu) is outside the hot path")
        if source.num_points() == 0 or target.num_points() == 0:
            return 0.0
        self.write_params()
        self.set_cloud(0, source)
        self.set_cloud(1, target)
        out = C.c_float(0.0)
        self._check(self._lib.cvo_b200_function_angle(
            self._h, _ptr(_colmajor16(T_target_to_source)), C.c_float(ell),
            int(bool(is_approximate)), C.byref(out)))
        return float(out.value)

    def compute_association_gpu(self, source, target, T_target_to_source, ell_or_kernel) -> Association:
        """CvoGPU::compute_association_gpu: float -> isotropic kernel with that length-scale
        (CvoGPU.cu:1876-1911); 3x3 matrix -> Mahalanobis kernel (:1975-1995)."""
        assoc = Association(shape=(source.num_points(), target.num_points()))
        if source.num_points() == 0 or target.num_points() == 0:
            return assoc
        self.write_params()
        self.set_cloud(0, source)
        self.set_cloud(1, target)
        T16 = _colmajor16(T_target_to_source)
        if np.ndim(ell_or_kernel) == 0:
            ell, K = float(ell_or_kernel), None
        else:
            ell, K = 0.0, np.ascontiguousarray(np.asarray(ell_or_kernel, np.float32).reshape(3, 3).T).reshape(9)
        self._fill_association(assoc, source.num_points(), target.num_points(),
                               lambda *a: self._lib.cvo_b200_association(
                                   self._h, _ptr(T16), C.c_float(ell), _ptr(K), *a))
        return assoc

    def _fill_association(self, assoc: Association, n: int, m: int, call):
        """Two-call CSR protocol of cvo_b200_association / cvo_b200_align_association, filled
        like gpu_association_to_cpu (CvoGPU_impl.cu:366-427)."""
        nnz = C.c_int64(0)
        row_ptr = np.zeros(n + 1, np.int32)
        i32p = C.POINTER(C.c_int32)
        self._check(call(C.byref(nnz), row_ptr.ctypes.data_as(i32p), None, None))
        cols = np.zeros(max(nnz.value, 1), np.int32)
        vals = np.zeros(max(nnz.value, 1), np.float32)
        if nnz.value > 0:
            self._check(call(C.byref(nnz), row_ptr.ctypes.data_as(i32p),
                             cols.ctypes.data_as(i32p), _ptr(vals)))
        assoc.shape = (n, m)
        assoc.row_ptr = row_ptr.astype(np.int64)
        assoc.cols = cols[: nnz.value]
        assoc.vals = vals[: nnz.value]
        return assoc

    # --- measurement helpers (bench.py)
    def time_iterations(self, R, T, ell: float, num_neighbors: int, iters: int, pair_kernel: bool = True):
        Rc = np.ascontiguousarray(np.asarray(R, np.float32).reshape(3, 3).T).reshape(9)
        Tc = _f32(T).reshape(3)
        ms_total, ms_pair = C.c_float(0), C.c_float(0)
        self._check(self._lib.cvo_b200_time_iterations(
            self._h, _ptr(Rc), _ptr(Tc), C.c_float(ell), int(num_neighbors), int(iters),
            C.byref(ms_total), C.byref(ms_pair) if pair_kernel else None))
        return float(ms_total.value), float(ms_pair.value)

    def launch_count(self) -> int:
        return int(self._lib.cvo_b200_launch_count(self._h))

    def last_candidate_builds(self) -> int:
        """Iterations of the last align() that built candidate cells (persistent tile mode)."""
        return int(self._lib.cvo_b200_last_candidate_builds(self._h))

    def stream(self) -> int:
        return int(self._lib.cvo_b200_stream(self._h) or 0)

    def fma_peak(self, kind: int = 1, iters: int = 4096) -> float:
        out = C.c_double(0.0)
        self._check(self._lib.cvo_b200_fma_peak(self._h, int(kind), int(iters), C.byref(out)))
        return float(out.value)

    # --- multi-GPU plumbing (one process per GPU; ids travel over torch.distributed)
    @staticmethod
    def comm_unique_id() -> bytes:
        lib = _abi.load_library()
        buf = (C.c_char * 128)()
        rc = lib.cvo_b200_comm_unique_id(buf)
        if rc != _abi.OK:
            raise CvoError(f"ncclGetUniqueId failed: {lib.cvo_b200_global_error().decode()}")
        return bytes(buf.raw)

    def comm_init(self, rank: int, world: int, unique_id: bytes):
        buf = (C.c_char * 128)(*unique_id[:128])
        self._check(self._lib.cvo_b200_comm_init(self._h, int(rank), int(world), buf))

    def comm_mailbox_handle(self) -> bytes:
        """CUDA IPC handle (64 bytes) of this rank's exchange mailbox (fused multi-GPU path)."""
        buf = (C.c_char * 64)()
        self._check(self._lib.cvo_b200_comm_mailbox_handle(self._h, buf))
        return bytes(buf.raw)

    def comm_open_peers(self, handles):
        """handles: the 64-byte mailbox handles of all ranks, in rank order."""
        self._check(self._lib.cvo_b200_comm_open_peers(self._h, b"".join(bytes(x[:64]) for x in handles)))

    def comm_shard_inner_products(self, on: bool = True):
        """inner_product_gpu / function_angle become collective calls of the job: every rank scans
        its shard of the source rows, one all-gather of the ranks' sums."""
        self._check(self._lib.cvo_b200_comm_shard_inner_products(self._h, int(bool(on))))

    def comm_destroy(self):
        self._check(self._lib.cvo_b200_comm_destroy(self._h))
