"""CPU (gloo, world_size 2) test of the multi-GPU decomposition's host logic: source rows are
sharded, every accumulated quantity is a sum (or max) over rows, ranks exchange their local
totals with one all-gather and reduce them in rank order -> identical on every rank and equal to
the unsharded result.  The per-shard numbers come from the oracle here; on the
GPU box tools/mgpu_check.py runs the same comparison through libcvo_b200 + NCCL."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from helpers import ROOT
from unified_cvo_b200.dist import shard_edges, shard_rows


def test_shard_rows_partition_is_exact():
    for n in (0, 1, 7, 64, 65, 10_000, 200_001):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_rows(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(0 <= b <= e <= n for b, e in spans)
    with pytest.raises(ValueError):
        shard_rows(10, 2, 2)


def test_shard_edges_partition_is_exact():
    for n in (0, 1, 5, 16):
        for world in (1, 2, 3, 8):
            parts = [shard_edges(n, world, r) for r in range(world)]
            assert sorted(e for part in parts for e in part) == list(range(n))
            assert max(len(x) for x in parts) - min(len(x) for x in parts) <= 1
    with pytest.raises(ValueError):
        shard_edges(4, 2, -1)


WORKER = textwrap.dedent("""
    import os, sys, pickle
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    import numpy as np
    import torch.distributed as dist
    import oracle
    import unified_cvo_b200 as u
    from unified_cvo_b200.dist import shard_rows
    from helpers import geometric_params, synthetic_pair, to_oracle_cloud
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    src, tgt, _ = synthetic_pair(900, 601, 700, 31)
    p = geometric_params()
    R, T, ell, cap = np.eye(3, dtype=np.float32).reshape(9), np.zeros(3, np.float32), 1.4, 6
    b, e = shard_rows(src.num_points(), world, rank)
    shard = u.CvoPointCloud(src.positions_[b:e])
    tr = oracle.iterate(p, to_oracle_cloud(shard), to_oracle_cloud(tgt), R, T, ell, cap)
    local = np.array(list(tr.omega_sum) + list(tr.v_sum) + [tr.a_sum, float(tr.nnz), float(tr.max_row_nnz)])
    gathered = [None] * world
    dist.all_gather_object(gathered, local)          # the 72-byte record of the GPU path
    tot = np.zeros(9)
    for g in gathered:                               # rank order: identical on every rank
        tot[:8] += g[:8]; tot[8] = max(tot[8], g[8])
    full = oracle.iterate(p, to_oracle_cloud(src), to_oracle_cloud(tgt), R, T, ell, cap)
    ref = np.array(list(full.omega_sum) + list(full.v_sum) + [full.a_sum, float(full.nnz), float(full.max_row_nnz)])
    np.testing.assert_allclose(tot[:7], ref[:7], rtol=1e-12, atol=1e-15)
    assert tot[7] == ref[7] and tot[8] == ref[8]
    sums = [None] * world
    dist.all_gather_object(sums, tot.tobytes())
    assert all(s == sums[0] for s in sums)           # bit-identical replicated control
    # edge-parallel pose-graph loop: every rank updates its own edges, the non-zero counts are
    # gathered in edge order and equal the single-process loop's
    from unified_cvo_b200.dist import shard_edges
    I = np.eye(4, dtype=np.float32)[:3].reshape(12)
    P = I.copy(); P[[3, 7, 11]] = [0.02, 0.0, 0.1]
    frames = [(src, I), (tgt, P), (src, P), (tgt, I)]
    edges = [(0, 1), (1, 2), (2, 3), (3, 0), (0, 2)]
    def edge_nnz(k):
        (a, pa), (b2, pb) = frames[edges[k][0]], frames[edges[k][1]]
        return oracle.edge_update(p, to_oracle_cloud(a), pa, to_oracle_cloud(b2), pb, 0.9, 12)[0]
    mine = {{k: edge_nnz(k) for k in shard_edges(len(edges), world, rank)}}
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    merged = {{k: v for part in parts for k, v in part.items()}}
    assert sorted(merged) == list(range(len(edges)))
    assert [merged[k] for k in range(len(edges))] == [edge_nnz(k) for k in range(len(edges))]
    assert sum(merged.values()) > 100
    # the NCCL unique-id plumbing: rank 0's 128 bytes reach every rank unchanged
    uid = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    assert uid[0] == bytes(range(128))
    dist.barrier()
    os.write(1, ("rank %d ok" % rank + os.linesep).encode())   # one write: the ranks share a pipe
""")


def test_two_rank_sharded_reduction_equals_unsharded(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    import socket
    with socket.socket() as sk:  # a free port: a fixed one can linger in TIME_WAIT between runs
        sk.bind(("127.0.0.1", 0))
        port = str(sk.getsockname()[1])
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=port, OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", port, str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout


def test_gather_association_merges_the_row_shards_of_two_ranks(tmp_path):
    """dist.gather_association on 2 gloo ranks: each rank holds the rows of its shard of one CSR
    matrix (the layout cvo_b200_align_association returns after a sharded align); rank 0 gets the
    whole matrix back, entry for entry."""
    worker = tmp_path / "worker.py"
    worker.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import numpy as np
        import torch.distributed as dist
        import unified_cvo_b200 as u
        from unified_cvo_b200.dist import gather_association
        rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        dist.init_process_group("gloo", rank=rank, world_size=world)
        rng = np.random.default_rng(5)
        n, m = 300, 90
        cnt = rng.integers(0, 7, n)
        cnt[rng.random(n) < 0.3] = 0
        row_ptr = np.concatenate([[0], np.cumsum(cnt)])
        cols = np.concatenate([np.sort(rng.choice(m, c, replace=False)) for c in cnt]).astype(np.int32)
        vals = rng.random(int(row_ptr[-1])).astype(np.float32)
        owner = rng.integers(0, world, n)   # Morton-contiguous shards are scattered in caller order
        mine = np.repeat(owner == rank, cnt)
        part = u.Association(shape=(n, m))
        part.row_ptr = np.concatenate([[0], np.cumsum(np.where(owner == rank, cnt, 0))])
        part.cols, part.vals = cols[mine], vals[mine]
        whole = gather_association(part, rank, world, dist)
        if rank == 0:
            assert np.array_equal(whole.row_ptr, row_ptr) and np.array_equal(whole.cols, cols)
            assert np.array_equal(whole.vals, vals) and whole.shape == (n, m)
            print("GATHER_ASSOC_OK", len(vals))
        else:
            assert whole is None
        dist.barrier()
    """))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29679", str(worker)],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-3000:])
    assert "GATHER_ASSOC_OK" in out.stdout
