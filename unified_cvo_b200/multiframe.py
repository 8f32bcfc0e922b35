"""Host mirror of the reference's pose-graph edge update.

`CvoFrameGPU` = cvo::CvoFrameGPU (include/UnifiedCvo/cvo/CvoFrameGPU.hpp, CvoFrameGPU.cu:7-62):
a point cloud kept on the device plus the frame's pose, `pose_vec`, a row-major 3x4 [R t] in
double that the outer solver (Ceres in the reference, out of scope here) mutates in place.
`BinaryStateGPU` = cvo::BinaryStateGPU (IRLS_State_GPU.hpp, IRLS_State_GPU.cu:16-79,
IRLS_State_GPU.cpp:54-57): one edge of the graph; `update_inner_product()` refills the capped
kernel matrix between the two moved frames and returns its number of non-zeros.
`update_edges` is the edge loop of one outer iteration of CvoBatchIRLS::solve
(IRLS.cpp:104-125) without the Ceres residuals.

Everything here only marshals numpy arrays into include/cvo_b200.h; the arithmetic is in
libcvo_b200.so (cvo_b200_frame_set / cvo_b200_edge_update).
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional

import numpy as np

from .cvo import Association, CvoGPU, CvoPointCloud, _ptr


class CvoFrameGPU:
    """cvo::CvoFrameGPU(pts, poses[12]): uploads the cloud once; the pose stays on the host."""

    def __init__(self, gpu: CvoGPU, points: CvoPointCloud, pose_vec=None):
        self.gpu = gpu
        self.points = points
        self.set_pose_vec(pose_vec)
        self.frame_id = getattr(gpu, "_next_frame_id", 0)  # ids are per handle
        gpu._next_frame_id = self.frame_id + 1
        F, Cn = points.feature_dimensions(), points.num_classes()
        gpu._check(gpu._lib.cvo_b200_frame_set(
            gpu._h, self.frame_id, points.num_points(), _ptr(points.positions_), F,
            _ptr(points.features_), Cn, _ptr(points.labels_), _ptr(points.geometric_types_)))

    def set_pose_vec(self, pose_vec=None):
        """The optimiser writes the frame's pose between outer iterations (CvoFrame::pose_vec)."""
        p = np.eye(4)[:3] if pose_vec is None else np.asarray(pose_vec, np.float64)
        if p.size == 16:
            p = p.reshape(4, 4)[:3]
        self.pose_vec = np.ascontiguousarray(p.reshape(12), np.float64)  # row-major 3x4 [R t]

    def pose_float(self) -> np.ndarray:
        """CvoFrameGPU.cu:47-53: the double pose narrowed to float for the device."""
        return np.ascontiguousarray(self.pose_vec, np.float64).astype(np.float32)

    def release(self):
        if self.gpu._h:
            self.gpu._check(self.gpu._lib.cvo_b200_frame_clear(self.gpu._h, self.frame_id))


class BinaryStateGPU:
    """cvo::BinaryStateGPU(pc1, pc2, params_cpu, params_gpu, num_neighbor, init_ell)."""

    def __init__(self, frame1: CvoFrameGPU, frame2: CvoFrameGPU, num_neighbor: Optional[int] = None,
                 init_ell: Optional[float] = None):
        assert frame1.gpu is frame2.gpu, "both frames of an edge live on one handle"
        self.gpu = frame1.gpu
        self.frame1, self.frame2 = frame1, frame2
        p = self.gpu.params
        # CvoGPU.cu:1663-1666: params.multiframe_num_neighbors, params.multiframe_ell_init
        self.init_num_neighbors_ = int(p.multiframe_num_neighbors if num_neighbor is None else num_neighbor)
        self.num_neighbors_ = self.init_num_neighbors_
        self.ell_ = float(p.multiframe_ell_init if init_ell is None else init_ell)
        self.iter_ = 0
        self.last_max_row_nnz = 0
        self.A_result_cpu_ = Association(shape=(frame1.points.num_points(), frame2.points.num_points()))
        self.A_result_cpu_.row_ptr = np.zeros(frame1.points.num_points() + 1, np.int64)
        self.A_result_cpu_.cols = np.zeros(0, np.int32)
        self.A_result_cpu_.vals = np.zeros(0, np.float32)

    def update_inner_product(self) -> int:
        """IRLS_State_GPU.cu:43-79.  Returns nonzero_sum; the matrix is in A_result_cpu_."""
        # :45-47 cap of this update from the LAST matrix' fullest row
        if self.last_max_row_nnz > 0:
            self.num_neighbors_ = min(self.init_num_neighbors_, int(self.last_max_row_nnz * 1.1))
        g = self.gpu
        n, m = self.frame1.points.num_points(), self.frame2.points.num_points()
        p1, p2 = self.frame1.pose_float(), self.frame2.pose_float()
        mx = C.c_int32(0)
        call = lambda *a: g._lib.cvo_b200_edge_update(  # noqa: E731
            g._h, self.frame1.frame_id, _ptr(p1), self.frame2.frame_id, _ptr(p2),
            C.c_float(self.ell_), int(self.num_neighbors_), a[0], C.byref(mx), *a[1:])
        g.write_params()
        g._fill_association(self.A_result_cpu_, n, m, call)
        self.last_max_row_nnz = int(mx.value)
        self.iter_ += 1
        return int(len(self.A_result_cpu_.vals))

    def update_ell(self):
        """IRLS_State_GPU.cpp:54-57."""
        p = self.gpu.params
        if self.ell_ > p.multiframe_ell_min:
            self.ell_ = self.ell_ * p.multiframe_ell_decay_rate


def update_edges(states: Iterable[BinaryStateGPU], batched: bool = True):
    """The edge loop of one outer iteration (IRLS.cpp:111-121): every edge refills its matrix
    from the frames' CURRENT poses; returns (total_nonzeros, per-edge nonzeros).

    batched (default): ONE cvo_b200_edge_update_batch call for all edges of a handle - the kernels
    of every edge are enqueued back to back and the host waits once; same matrices as the per-edge
    calls.  Falls back to the per-edge loop for edges the batch cannot serve."""
    states = list(states)
    if not batched or not states or any(s.gpu is not states[0].gpu for s in states):
        per_edge = [s.update_inner_product() for s in states]
        return int(sum(per_edge)), per_edge
    from ._abi import Edge
    g = states[0].gpu
    g.write_params()
    n = len(states)
    edges = (Edge * n)()
    for e, st in zip(edges, states):
        if st.last_max_row_nnz > 0:  # IRLS_State_GPU.cu:45-47
            st.num_neighbors_ = min(st.init_num_neighbors_, int(st.last_max_row_nnz * 1.1))
        e.frame1, e.frame2 = st.frame1.frame_id, st.frame2.frame_id
        e.pose1[:] = st.frame1.pose_float().tolist()
        e.pose2[:] = st.frame2.pose_float().tolist()
        e.ell, e.num_neighbors = st.ell_, int(st.num_neighbors_)
    rows = [st.frame1.points.num_points() for st in states]
    nnz = (C.c_int64 * n)()
    mx = (C.c_int32 * n)()
    row_ptr = np.zeros(sum(rows) + n, np.int32)
    i32p = C.POINTER(C.c_int32)
    rc = g._lib.cvo_b200_edge_update_batch(g._h, n, edges, nnz, mx, row_ptr.ctypes.data_as(i32p), None, None)
    if rc == -5:  # CVO_B200_ERR_STATE: an edge outside the cell-query regimes
        per_edge = [s.update_inner_product() for s in states]
        return int(sum(per_edge)), per_edge
    g._check(rc)
    total = int(sum(nnz))
    cols = np.zeros(max(total, 1), np.int32)
    vals = np.zeros(max(total, 1), np.float32)
    if total:
        g._check(g._lib.cvo_b200_edge_update_batch(g._h, n, edges, nnz, mx, row_ptr.ctypes.data_as(i32p),
                                                   cols.ctypes.data_as(i32p), _ptr(vals)))
    per_edge, ro, eo = [], 0, 0
    for k, st in enumerate(states):
        a = st.A_result_cpu_
        a.shape = (rows[k], st.frame2.points.num_points())
        a.row_ptr = row_ptr[ro:ro + rows[k] + 1].astype(np.int64)
        a.cols = cols[eo:eo + int(nnz[k])]  # views of the round's buffers (fresh arrays every call)
        a.vals = vals[eo:eo + int(nnz[k])]
        st.last_max_row_nnz = int(mx[k])
        st.iter_ += 1
        per_edge.append(int(nnz[k]))
        ro += rows[k] + 1
        eo += int(nnz[k])
    return total, per_edge


def update_edges_sharded(states, rank: int, world: int, dist_module=None, gather_to: Optional[int] = 0):
    """Edge-parallel edge loop over the GPUs of a job: the edges of a pose graph
    are independent, so rank r refills edges r, r + world, ... on ITS GPU with one batch call and
    there is no data-path collective.  The matrices feed a host solver (Ceres in the reference), so
    with gather_to = k the other ranks' CSR matrices travel to rank k over the host control plane
    (dist_module.gather_object) and are stored into its states; gather_to = None keeps them where
    they were computed.  Every rank must hold the frames of its own edges (states of edges it does
    not own are never touched on the device).  Returns (nonzeros of this rank's edges, their indices)."""
    from .dist import shard_edges
    states = list(states)
    mine = shard_edges(len(states), world, rank)
    total_local, _ = update_edges([states[i] for i in mine])
    if dist_module is None or world == 1 or gather_to is None:
        return total_local, mine
    # one compact buffer per rank: per edge (index, rows, nnz, fullest row), then the per-row counts
    # as uint16 (a row holds at most nearest_neighbors_max entries), the columns and the values
    head, parts = [len(mine)], []
    for i in mine:
        a = states[i].A_result_cpu_
        head += [i, len(a.row_ptr) - 1, len(a.vals), states[i].last_max_row_nnz]
        parts += [np.diff(a.row_ptr).astype(np.uint16).tobytes(), np.ascontiguousarray(a.cols, np.int32).tobytes(),
                  np.ascontiguousarray(a.vals, np.float32).tobytes()]
    payload = np.asarray(head, np.int64).tobytes() + b"".join(parts)
    gathered = [None] * world if rank == gather_to else None
    dist_module.gather_object(payload, gathered, dst=gather_to)
    if rank == gather_to:
        for r, buf in enumerate(gathered):
            if r == rank:
                continue
            k = int(np.frombuffer(buf, np.int64, 1)[0])
            meta = np.frombuffer(buf, np.int64, 4 * k, 8).reshape(k, 4)
            off = 8 * (1 + 4 * k)
            for i, rows, nnz, mx in meta.tolist():
                cnt = np.frombuffer(buf, np.uint16, rows, off)
                off += 2 * rows
                a = states[i].A_result_cpu_
                a.row_ptr = np.concatenate([[0], np.cumsum(cnt, dtype=np.int64)])
                a.cols = np.frombuffer(buf, np.int32, nnz, off)
                off += 4 * nnz
                a.vals = np.frombuffer(buf, np.float32, nnz, off)
                off += 4 * nnz
                a.shape = (rows, states[i].frame2.points.num_points())
                states[i].last_max_row_nnz = int(mx)
    return total_local, mine
