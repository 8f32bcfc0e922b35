"""GPU parity of the controller's schedule and of the branches round 1
left untested.

a12 = A_sparsity_indicator_ell_update (CvoGPU.cu:1167-1285: the two sliding windows with the
double-counted sample, clear-on-decay), the `k > ell_decay_start` gate and the decay itself
(:1486-1513), the neighbour-cap update (:1518-1529), and what align reports at the end.  The
device controller runs inside cvo_b200_align only, so these tests run the FREE loop on both
sides with schedules that fire within the first iterations (small windows, ell_decay_start 0)
and compare the records iteration by iteration while the trajectories still agree — the loop is
chaotic (tests/test_oracle.py), so a long agreeing prefix is required, not assumed: every test
asserts how long it was and that the events it is about happened inside it.
"""
import os

import numpy as np
import pytest

import oracle
import unified_cvo_b200 as u
from unified_cvo_b200 import _abi
from helpers import (DATA, compare_traces, demo_clouds, demo_params, geometric_params, synthetic_pair,
                     to_oracle_cloud)

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["dense", "grid", "grid-launches", "tile", "tile-launches", "brute"])
def candidate_mode(request, monkeypatch):
    """dense N x M scan / cell queries in the persistent kernel / cell queries as one launch per
    phase / tile cells in the persistent kernel (built with a skin and REUSED while the pose has
    drifted less than it) / tile cells as one launch per phase / every row walked exactly by one
    warp inside the persistent kernel (read by cvo_b200_create): the controller exists in each
    launch structure."""
    monkeypatch.setenv("CVO_B200_MODE", request.param.split("-")[0])
    monkeypatch.setenv("CVO_B200_PERSIST", "0" if request.param.endswith("-launches") else "1")
    return request.param


def agreeing_prefix(tr_g, tr_o):
    """Number of leading iterations whose records agree: integers and the schedule exactly, the
    twist / step polynomial within north_star's 1e-4."""
    n = 0
    for a, b in zip(tr_g, tr_o):
        if compare_traces(a, b, twist_tol=1e-4):
            break
        if (a.iter, a.num_neighbors, a.flags, a.num_neighbors_next) != (b.iter, b.num_neighbors, b.flags,
                                                                       b.num_neighbors_next):
            break
        if a.ell != b.ell or a.ell_next != b.ell_next:
            break
        n += 1
    return n


SCHEDULES = [
    # window, stable threshold, decay start, ell_init: what the case exercises
    (2, 0.2, 0, 2.5),      # decays every third iteration from k = 2 on, cap follows 1.2 * max row
    (3, 0.2, 0, 0.95),     # longer windows, the double-counted sample that fills the start queue
    (3, 0.0005, 0, 1.5),   # band so narrow that the windows mostly SLIDE (no decay)
    (2, 0.2, 7, 1.5),      # requests before k > ell_decay_start clear the queues but do not decay
]


@pytest.mark.parametrize("window,thr,start,ell0", SCHEDULES)
def test_schedule_queues_decay_and_cap_follow_the_oracle(candidate_mode, window, thr, start, ell0):
    src, tgt, _ = synthetic_pair(2500, 2000, 2000, 20002)
    p = geometric_params(ell0)
    p.ell_decay_start = start
    p.indicator_window_size = window
    p.indicator_stable_threshold = thr
    p.ell_decay_rate = 0.9
    p.MAX_ITER = 16
    cs, ct = to_oracle_cloud(src), to_oracle_cloud(tgt)
    ret_o, T_o, info_o, tr_o = oracle.align(p, cs, ct, None, trace_cap=16)
    g = u.CvoGPU(p, device=0)
    ret_g, T_g, info_g, tr_g = g.align(src, tgt, None, trace_cap=16)
    g.close()
    assert len(tr_o) == len(tr_g) == 16
    n = agreeing_prefix(tr_g, tr_o)
    decays = [t.iter for t in tr_o[:n] if t.flags & _abi.ELL_DECAYED]
    caps = {t.num_neighbors for t in tr_o[:n]}
    print(f"{candidate_mode} window {window} thr {thr} start {start}: agreeing prefix {n}/16, decays at {decays}, caps {sorted(caps)}")
    assert n >= 12, f"only {n} iterations agree"
    if thr >= 0.1:
        assert len(decays) >= 2 and all(k > start for k in decays)
        # the decay is what the record says: ell_next = max(ell * rate, ell_min)
        for t in tr_g[:n]:
            if t.flags & _abi.ELL_DECAYED:
                assert t.ell_next == np.float32(max(np.float32(t.ell) * np.float32(p.ell_decay_rate),
                                                    np.float32(p.ell_min)))
            else:
                assert t.ell_next == t.ell
    else:
        assert len(decays) <= 1
    assert len(caps) >= 3  # the cap moved: 256 -> 1.2 * max row count -> ...
    for t in tr_g[:n]:
        assert t.num_neighbors_next == min(int(p.nearest_neighbors_max), int(t.max_row_nnz * 1.2))
    if n == 16:
        # both loops ran into MAX_ITER with identical schedules: what align reports at the end
        assert ret_g == ret_o == 0
        assert info_g.iterations == info_o.iterations == 16
        assert info_g.stop_reason == info_o.stop_reason == _abi.STOP_MAX_ITER
        assert info_g.final_ell == info_o.final_ell
        assert info_g.final_num_neighbors == info_o.final_num_neighbors
        assert np.abs(T_g - T_o).max() <= 1e-5


def test_schedule_with_colour_on_the_demo_pair(candidate_mode):
    """the README demo with its first-frame schedule made to fire early: colour kernel, geometric
    types, 523 x 1080 points, ell_init = 5.76 (rows cut at the cap of 256)."""
    src, tgt = demo_clouds(True)
    p = demo_params(src, tgt, True)
    p.ell_decay_start = 0
    p.indicator_window_size = 2
    p.indicator_stable_threshold = 0.3
    p.MAX_ITER = 14
    ret_o, T_o, info_o, tr_o = oracle.align(p, to_oracle_cloud(src), to_oracle_cloud(tgt), None, trace_cap=14)
    g = u.CvoGPU(p, device=0)
    ret_g, T_g, info_g, tr_g = g.align(src, tgt, None, trace_cap=14)
    g.close()
    n = agreeing_prefix(tr_g, tr_o)
    decays = [t.iter for t in tr_o[:n] if t.flags & _abi.ELL_DECAYED]
    print(f"demo: agreeing prefix {n}/14, decays at {decays}")
    assert n >= 12 and len(decays) >= 2
    if n == 14:
        assert info_g.final_ell == info_o.final_ell and info_g.final_num_neighbors == info_o.final_num_neighbors


def test_range_ell_branch_of_the_step_polynomial(candidate_mode):
    """is_using_range_ell = 1: compute_step_size_poly_coeff scales ell per row
    (CvoGPU.cu:1036-1037) — teacher-forced iterations against the oracle, whose rows are pinned
    against the reference's own K4 text (tests/test_ref_pin.py)."""
    src, tgt, _ = synthetic_pair(2500, 2000, 2000, 20002)
    p = geometric_params()
    p.is_using_range_ell = 1
    cs, ct = to_oracle_cloud(src), to_oracle_cloud(tgt)
    _, _, _, tr = oracle.align(p, cs, ct, None, trace_cap=41)
    q = p.copy()
    q.is_using_range_ell = 0
    g = u.CvoGPU(p, device=0)
    g.set_cloud(0, src)
    g.set_cloud(1, tgt)
    differs = 0
    for k in (0, 1, 5, 20, 40):
        if k == 0:
            R9, T3 = np.eye(3, dtype=np.float32).reshape(9), np.zeros(3, np.float32)
        else:
            R9, T3 = np.array(list(tr[k - 1].R), np.float32), np.array(list(tr[k - 1].T), np.float32)
        ell, cap = float(tr[k].ell), int(tr[k].num_neighbors)
        want = oracle.iterate(p, cs, ct, R9, T3, ell, cap)
        got = g.iterate(R9.reshape(3, 3).T, T3, ell, cap)
        assert not compare_traces(got, want, twist_tol=1e-4), k
        off = oracle.iterate(q, cs, ct, R9, T3, ell, cap)
        differs += int(abs(off.B - want.B) > 1e-3 * abs(want.B))
    g.close()
    assert differs >= 3  # the branch changes B..E by far more than the tolerance: it was really taken


def test_c4_full_size_iteration_at_the_benchmarked_ell():
    """One whole iteration (flow, step polynomial, pose update) of C4 = 200 000 x 200 000 with
    colour at ell = 1.5, identity pose — the regime SCALE measures — against the oracle.  (The
    kernel matrix of the same state is held to the reference's own kernel bit for bit in
    tests/test_ref_pin_gpu.py.)"""
    src, tgt, _ = synthetic_pair(250000, 200000, 200000, 20004, F=5)
    p = u.read_params_yaml(os.path.join(DATA, "cvo_intensity_params_img_gpu0.yaml"))
    cs, ct = to_oracle_cloud(src), to_oracle_cloud(tgt)
    R9, T3 = np.eye(3, dtype=np.float32).reshape(9), np.zeros(3, np.float32)
    want = oracle.iterate(p, cs, ct, R9, T3, 1.5, int(p.nearest_neighbors_max))
    g = u.CvoGPU(p, device=0)
    g.set_cloud(0, src)
    g.set_cloud(1, tgt)
    got = g.iterate(np.eye(3), T3, 1.5, int(p.nearest_neighbors_max))
    g.close()
    assert not compare_traces(got, want, twist_tol=1e-4)
    assert got.nnz == want.nnz > 10000
    assert np.abs(np.array(list(got.R)) - np.array(list(want.R))).max() <= 1e-6
    assert np.abs(np.array(list(got.T)) - np.array(list(want.T))).max() <= 1e-5


@pytest.mark.parametrize("colour", [False, True])
def test_reused_candidate_cells_lose_no_pair(monkeypatch, colour):
    """Persistent tile mode keeps a row's candidate cells over several iterations (built with a skin,
    verlet_decide in cvo_kernels.cu).  A lost pair would show as a different nnz: the free-running
    loop must agree with the oracle record by record (nnz exactly) over a prefix that contains
    reused iterations AND rebuilds, and end where the rebuild-every-iteration run ends."""
    monkeypatch.setenv("CVO_B200_MODE", "tile")
    monkeypatch.setenv("CVO_B200_PERSIST", "1")
    if colour:
        src, tgt, _ = synthetic_pair(5000, 4000, 4000, 20005, F=5)
        p = u.read_params_yaml(os.path.join(DATA, "cvo_intensity_params_img_gpu0.yaml"))
        p.ell_init, p.c_ell = 1.5, 1.0  # random colours: loosen the colour kernel so that rows fill
    else:
        src, tgt, _ = synthetic_pair(2500, 2000, 2000, 20002)
        p = geometric_params(0.95)
    p.MAX_ITER = 48
    cs, ct = to_oracle_cloud(src), to_oracle_cloud(tgt)
    _, T_o, _, tr_o = oracle.align(p, cs, ct, None, trace_cap=48)
    out = {}
    for kappa in ("0.1", "0.02", "0"):
        monkeypatch.setenv("CVO_B200_VERLET", kappa)
        g = u.CvoGPU(p, device=0)
        _, T_g, info, tr_g = g.align(src, tgt, None, trace_cap=48)
        out[kappa] = (T_g, len(tr_g), g.last_candidate_builds(), agreeing_prefix(tr_g, tr_o), [t.nnz for t in tr_g])
        g.close()
    for kappa, (T_g, n_it, builds, prefix, nnz) in out.items():
        print(f"colour {colour} kappa {kappa}: {n_it} iterations, {builds} builds, prefix {prefix}, nnz[0] {nnz[0]}")
    base = out["0"][3]  # how long the rebuild-every-iteration run follows the oracle (the loop is chaotic)
    assert base >= 12
    for kappa, (T_g, n_it, builds, prefix, nnz) in out.items():
        assert prefix >= base - 2, (kappa, prefix, base)
        assert nnz[:prefix] == [t.nnz for t in tr_o[:prefix]] and nnz[0] > 1000
    n_it = out["0"][1]
    assert out["0"][2] == n_it                      # no skin: every iteration builds
    assert 1 <= out["0.1"][2] <= (2 * n_it) // 3    # cells reused ...
    assert out["0.1"][2] <= out["0.02"][2] <= n_it  # ... and a thinner skin is used up sooner
    assert out["0.02"][2] >= 2                      # rebuilds happened inside the run
    # reused iterations lie INSIDE the compared prefix: the same run stopped there built fewer times
    # than it iterated
    monkeypatch.setenv("CVO_B200_VERLET", "0.1")
    p.MAX_ITER = out["0.1"][3]
    g = u.CvoGPU(p, device=0)
    g.align(src, tgt, None)
    builds_in_prefix = g.last_candidate_builds()
    g.close()
    print(f"builds inside the agreeing prefix of {p.MAX_ITER}: {builds_in_prefix}")
    assert 1 <= builds_in_prefix < p.MAX_ITER


def test_stop_test_shortcut_never_changes_the_loop(candidate_mode):
    """The controller skips the SE3 log of the pose increment when step * |twist| is far from eps_2
    (controller_step); with a trace attached every iteration takes the exact path.  Both loops must
    run the same iterations, stop for the same reason and end on the same bits — on a run that ends
    through the eps_2 test and on one that ends at MAX_ITER."""
    src, tgt, _ = synthetic_pair(2500, 2000, 2000, 20002)
    for max_iter, eps_2 in ((3000, 1.2e-5), (3000, 2e-4), (40, 1.2e-5)):
        p = geometric_params(0.95)
        p.MAX_ITER, p.eps_2 = max_iter, eps_2
        g = u.CvoGPU(p, device=0)
        ret_a, T_a, info_a, tr = g.align(src, tgt, None, trace_cap=max_iter)
        ret_b, T_b, info_b = g.align(src, tgt, None)
        g.close()
        print(f"{candidate_mode} MAX_ITER {max_iter} eps_2 {eps_2}: {info_a.iterations} iterations, stop {info_a.stop_reason}, "
              f"last dist {tr[-1].dist:.3e}")
        assert (ret_a, info_a.iterations, info_a.stop_reason) == (ret_b, info_b.iterations, info_b.stop_reason)
        assert info_a.final_ell == info_b.final_ell and info_a.final_num_neighbors == info_b.final_num_neighbors
        assert np.array_equal(np.asarray(T_a, np.float32).view(np.uint32), np.asarray(T_b, np.float32).view(np.uint32))
        # the recorded distances are the exact ones and agree with step * |twist| to the bound the shortcut assumes
        for t in tr:
            d_fast = float(t.step) * float(np.sqrt(sum(float(x) ** 2 for x in list(t.omega) + list(t.v))))
            assert abs(t.dist - d_fast) <= 1e-6 + 5e-5 * d_fast, (t.iter, t.dist, d_fast)
    assert info_a.stop_reason == _abi.STOP_MAX_ITER


def test_two_gpus_equal_one(tmp_path):
    """Sharded == single GPU: the first iterations of a source-sharded align on two
    GPUs (NCCL path and fused NVLink-mailbox path) reproduce the single-GPU records, and both
    ranks end on bit-identical poses.  Skipped on a one-GPU box."""
    n_dev = int(_abi.load_library().cvo_b200_device_count())
    if n_dev < 2:
        pytest.skip("needs two GPUs")
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29641",
                          os.path.join(root, "tools", "mgpu_check.py")], capture_output=True, text=True,
                         timeout=600, cwd=root)
    print(out.stdout[-3000:], out.stderr[-2000:])
    assert out.returncode == 0
    assert "MGPU_CHECK_OK" in out.stdout
