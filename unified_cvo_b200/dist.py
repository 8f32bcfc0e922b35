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
