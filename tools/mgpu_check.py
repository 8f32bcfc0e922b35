"""Multi-GPU check (torchrun, one rank per GPU): sharded align == single-GPU align."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch, torch.distributed as dist
import unified_cvo_b200 as u
from helpers import *

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("gloo", rank=rank, world_size=world)
torch.cuda.set_device(lr)
src, tgt, Tgt = synthetic_pair(12500, 10000, 10000, 20002)
p = geometric_params(); p.MAX_ITER = 60
single = u.CvoGPU(p, device=lr)
r0, T0, i0, tr0 = single.align(src, tgt, None, trace_cap=60)
g = u.CvoGPU(p, device=lr)
uid = [u.CvoGPU.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
g.comm_init(rank, world, uid[0])
from unified_cvo_b200.dist import shard_rows
g.set_row_range(*shard_rows(10000, world, rank))
r1, T1, i1, tr1 = g.align(src, tgt, None, trace_cap=60)   # NCCL all-gathers, one launch per phase
handles = [None] * world
dist.all_gather_object(handles, g.comm_mailbox_handle())
g.comm_open_peers(handles)
r2, T2, i2, tr2 = g.align(src, tgt, None, trace_cap=60)   # fused: persistent kernel + NVLink mailboxes
l0 = g.launch_count()
r2, T2, i2, tr2 = g.align(src, tgt, None, trace_cap=60)
print(f"rank {rank}: launches of the fused align: {g.launch_count() - l0}")
if os.environ.get("CVO_B200_STAMPS"):
    g.time_iterations(np.eye(3), np.zeros(3), 0.3, 64, 200, pair_kernel=False)
okf = all(not compare_traces(tr2[k], tr0[k]) for k in range(8))
assert okf, "fused path: first 8 iterations differ from the single-GPU run"
assert np.array_equal(T2, T1), "fused and NCCL paths end on different poses"
print(f"rank {rank}: fused x{world} iters {i2.iterations} t={i2.registration_seconds*1e3:.2f} ms | pose diff vs NCCL path {np.abs(T2-T1).max():.2e} | first-8 parity {'OK' if okf else 'FAIL'}")
ok = True
for k in range(8):
    bad = compare_traces(tr1[k], tr0[k])
    if bad: ok = False; print(rank, "iter", k, bad)
print(f"rank {rank}: single iters {i0.iterations} t={i0.registration_seconds*1e3:.2f} ms | sharded x{world} iters {i1.iterations} t={i1.registration_seconds*1e3:.2f} ms | pose diff {np.abs(T1-T0).max():.2e} | first-8 parity {'OK' if ok else 'FAIL'}")
# scalar inner products: sharded (one all-gather of the ranks' sums) == every rank on its own
Tpose = np.linalg.inv(T1).astype(np.float32)
fa_own = g.function_angle(src, tgt, Tpose, 0.5)
ip_own = g.inner_product_gpu(src, tgt, Tpose, 0.5)
g.comm_shard_inner_products(True)
fa_sh = g.function_angle(src, tgt, Tpose, 0.5)
ip_sh = g.inner_product_gpu(src, tgt, Tpose, 0.5)
fa_sh_exact = g.function_angle(src, tgt, Tpose, 0.5, is_approximate=False)
g.comm_shard_inner_products(False)
fa_own_exact = g.function_angle(src, tgt, Tpose, 0.5, is_approximate=False)
print(f"rank {rank}: function_angle own {fa_own:.7f} sharded {fa_sh:.7f} | inner product own {ip_own:.4f} sharded {ip_sh:.4f} | exact own {fa_own_exact:.7f} sharded {fa_sh_exact:.7f}")
assert abs(ip_sh - ip_own) <= 2e-6 * abs(ip_own) and abs(fa_sh - fa_own) <= 2e-6 * abs(fa_own), "sharded inner product differs"
assert abs(fa_sh_exact - fa_own_exact) <= 4e-6 * abs(fa_own_exact)
vals = [None] * world
dist.all_gather_object(vals, (fa_sh, ip_sh))
assert all(v == vals[0] for v in vals), "ranks disagree on the sharded inner product"
# the kernel matrix of the LAST iteration: every rank exports the rows of its shard, rank 0 merges
# them (dist.gather_association) - equal to the single-GPU export entry for entry
from unified_cvo_b200.dist import gather_association
q = p.copy(); q.is_exporting_association = 1; q.MAX_ITER = 12
a_single, a_part = u.Association(), u.Association()
single.write_params(q); g.write_params(q)
single.align(src, tgt, None, association=a_single)
g.align(src, tgt, None, association=a_part)
whole = gather_association(a_part, rank, world, dist)
if rank == 0:
    assert len(a_single.vals) > 1000 and 0 < len(a_part.vals) < len(a_single.vals)
    assert np.array_equal(whole.row_ptr, a_single.row_ptr) and np.array_equal(whole.cols, a_single.cols)
    assert np.array_equal(whole.vals.view(np.uint32), a_single.vals.view(np.uint32))
    print(f"sharded association: {len(a_part.vals)} of {len(a_single.vals)} entries on rank 0, merged == single-GPU export")
single.write_params(p); g.write_params(p)
poses = [None] * world
dist.all_gather_object(poses, T1.tobytes())
assert ok, "NCCL path: first 8 iterations differ from the single-GPU run"
assert all(b == poses[0] for b in poses), "ranks ended on different poses"
if rank == 0:
    print("all ranks bit-identical pose:", all(b == poses[0] for b in poses))
dist.barrier()
g.close(); single.close()
if rank == 0:
    print("MGPU_CHECK_OK")
