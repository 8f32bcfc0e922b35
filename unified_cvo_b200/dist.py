"""Host-side plumbing of the multi-GPU path: one process per GPU, source rows sharded,
target replicated, NCCL unique id exchanged over whatever control plane the host has
(torch.distributed here).  The data path itself (two tiny all-gathers per iteration) lives
in libcvo_b200.so."""
from __future__ import annotations

ROW_TILE = 64  # kTileRows of csrc/cvo_device.cuh


def shard_rows(n_rows: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block partition of the source rows: [begin, end) of `rank`.  Shards start on
    a multiple of 64 rows (the source tile of the pair kernel) so that every rank can use the
    per-tile bounding spheres of the Morton-ordered cloud."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    per = (n_rows + world - 1) // world
    per = (per + ROW_TILE - 1) // ROW_TILE * ROW_TILE
    begin = min(n_rows, rank * per)
    return begin, min(n_rows, begin + per)


def shard_edges(n_edges: int, world: int, rank: int) -> list[int]:
    """Edge-parallel decomposition of the pose-graph edge loop (IRLS.cpp:111-121): the edges are
    independent (each fills its own matrix from the two frames' poses), so rank r takes edges
    r, r + world, ... and uploads only the frames those edges touch.  No data-path collective:
    the per-edge matrices go to the host solver; only the per-edge non-zero counts are gathered
    (host control plane) for the loop's `total_nonzeros` test."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_edges, world))


def attach(gpu, n_rows: int, rank: int, world: int, dist_module=None, fused: bool = True) -> tuple[int, int]:
    """Create the NCCL communicator of `gpu` (a CvoGPU) and set its row shard.

    dist_module: an initialised torch.distributed (any backend) used only to broadcast the
    128-byte NCCL unique id from rank 0."""
    from .cvo import CvoGPU

    if world > 1:
        uid = [CvoGPU.comm_unique_id() if rank == 0 else None]
        dist_module.broadcast_object_list(uid, src=0)
        gpu.comm_init(rank, world, uid[0])
        if fused:  # NVLink mailboxes: the persistent kernel exchanges its records itself
            handles = [None] * world
            dist_module.all_gather_object(handles, gpu.comm_mailbox_handle())
            gpu.comm_open_peers(handles)
    b, e = shard_rows(n_rows, world, rank)
    gpu.set_row_range(b, e)
    return b, e


def gather_association(assoc, rank: int, world: int, dist_module, dst: int = 0):
    """The kernel matrix of a SHARDED align on one rank ("association rows gathered").
    After a sharded align with is_exporting_association every rank's Association holds the rows of
    its own source shard (all other rows empty, same shape); the rows are disjoint, so the whole
    matrix is their row-wise union.  The parts travel over the host control plane
    (dist_module.gather_object); returns the merged Association on `dst`, None elsewhere."""
    import numpy as np
    from .cvo import Association

    cnt = np.diff(np.asarray(assoc.row_ptr, np.int64))
    rows = np.nonzero(cnt)[0]
    part = (assoc.shape, rows.astype(np.int32), cnt[rows].astype(np.int32),
            np.ascontiguousarray(assoc.cols, np.int32), np.ascontiguousarray(assoc.vals, np.float32))
    parts = [None] * world if rank == dst else None
    dist_module.gather_object(part, parts, dst=dst)
    if rank != dst:
        return None
    shape = parts[0][0]
    total = np.zeros(shape[0], np.int64)
    for sh, r, c, _, _ in parts:
        if tuple(sh) != tuple(shape):
            raise ValueError("ranks hold associations of different shapes")
        if np.any(total[r] != 0):
            raise ValueError("two ranks hold entries of the same source row")
        total[r] = c
    row_ptr = np.concatenate([[0], np.cumsum(total)])
    cols = np.zeros(int(row_ptr[-1]), np.int32)
    vals = np.zeros(int(row_ptr[-1]), np.float32)
    for _, r, c, pc, pv in parts:
        src_off = np.concatenate([[0], np.cumsum(c, dtype=np.int64)])
        # entry k of part-row q goes to row_ptr[r[q]] + k
        dest = np.repeat(row_ptr[r] - src_off[:-1], c) + np.arange(int(src_off[-1]))
        cols[dest] = pc
        vals[dest] = pv
    out = Association(shape=tuple(shape))
    out.row_ptr, out.cols, out.vals = row_ptr, cols, vals
    return out
