"""GPU parity tests: the CUDA path (through the C-ABI) against the oracle on the same inputs.

Parity protocol: CVO trajectories are chaotic under 1e-7 perturbations
(tests/test_oracle.py::test_trajectories_are_chaotic...), so per-iteration parity is
TEACHER-FORCED: the state (pose, ell, row cap) of iteration k is taken from the oracle's own
trajectory and both sides compute that one iteration.  Tolerances follow north_star:
1e-4 relative on the normalised se(3) twist per iteration, 1e-3 on the final pose.
Integer outputs (nnz, max row count, the association's index pattern) must be exact.
"""
import json
import os

import numpy as np
import pytest

import oracle
import unified_cvo_b200 as u
from helpers import (DATA, compare_traces, demo_clouds, demo_params, geometric_params, rel_err,
                     synthetic_pair, to_oracle_cloud)

pytestmark = pytest.mark.gpu

TWIST_TOL = 1e-4   # north_star: relative error of the se(3) twist per iteration
POSE_TOL = 1e-3    # north_star: final pose


def _col9(R):
    return np.asarray(R, np.float32).reshape(9)


def _teacher_forced(gpu, p, src, tgt, ks, max_iter=None):
    """Run the oracle's align, then replay iterations `ks` on both sides from the oracle's states."""
    cs, ct = to_oracle_cloud(src), to_oracle_cloud(tgt)
    if max_iter is not None:
        p = p.copy()
        p.MAX_ITER = max_iter
    _, _, info, tr = oracle.align(p, cs, ct, None, trace_cap=max(ks) + 1)
    gpu.set_cloud(0, src)
    gpu.set_cloud(1, tgt)
    failures = []
    for k in ks:
        if k >= len(tr):
            continue
        if k == 0:
            R9, T3 = np.eye(3, dtype=np.float32).reshape(9), np.zeros(3, np.float32)
        else:
            R9, T3 = _col9(list(tr[k - 1].R)), np.array(list(tr[k - 1].T), np.float32)
        ell, cap = float(tr[k].ell), int(tr[k].num_neighbors)
        ref = oracle.iterate(p, cs, ct, R9, T3, ell, cap)
        got = gpu.iterate(R9.reshape(3, 3).T, T3, ell, cap)
        bad = compare_traces(got, ref, twist_tol=TWIST_TOL)
        # pose after the update: float rounding of one Exp map
        if np.abs(np.array(list(got.R)) - np.array(list(ref.R))).max() > 1e-6:
            bad.append("R after update")
        if np.abs(np.array(list(got.T)) - np.array(list(ref.T))).max() > 1e-5:
            bad.append("T after update")
        if bad:
            failures.append((k, bad))
    return failures


@pytest.fixture(scope="module", params=["dense", "grid", "grid-launches", "tile", "tile-launches", "brute", "auto"], autouse=True)
def candidate_mode(request):
    """Every test runs with the dense N x M scan (pair_kernel), with cell queries in the
    persistent cooperative kernel (align_grid_kernel), with cell queries as one launch per phase
    (flow_kernel_t<1> / step_kernel_t<true>), with tile cells inside the persistent kernel and as
    one launch per phase (tile_kernel + flow_kernel_t<2>; wherever the Morton view applies), with
    every row walked exactly by one warp inside the persistent kernel ("brute": what the policy
    picks for a few hundred rows against a small target) and with the automatic choice: the candidate
    generator and the launch structure must never change a result.  Read by cvo_b200_create."""
    old = {k: os.environ.get(k) for k in ("CVO_B200_MODE", "CVO_B200_PERSIST")}
    os.environ["CVO_B200_MODE"] = request.param.split("-")[0]
    os.environ["CVO_B200_PERSIST"] = "0" if request.param.endswith("-launches") else "1"
    yield request.param
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.fixture(scope="module")
def gpu_geo(candidate_mode):
    g = u.CvoGPU(geometric_params())
    yield g
    g.close()


def test_teacher_forced_iterations_synthetic_geometric(gpu_geo):
    src, tgt, _ = synthetic_pair(2500, 2000, 2000, 20002)
    p = geometric_params()
    gpu_geo.write_params(p)
    fails = _teacher_forced(gpu_geo, p, src, tgt, [0, 1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233, 377, 500])
    assert not fails, fails


def test_teacher_forced_iterations_c2_full_size(gpu_geo):
    """BASELINE config 2 (N=M=10k, geometric) at full size, first iterations."""
    src, tgt, _ = synthetic_pair(12_500, 10_000, 10_000, 20_002)
    p = geometric_params()
    gpu_geo.write_params(p)
    fails = _teacher_forced(gpu_geo, p, src, tgt, [0, 1, 2, 5, 9], max_iter=10)
    assert not fails, fails


@pytest.mark.parametrize("color", [True, False])
def test_teacher_forced_iterations_demo_pcds(color):
    """BASELINE config 1: the README demo clouds, colour and geometric-only flavours.  ell_init =
    5.76 m makes 16 % of all pairs pass the geometric cut, so the row cap (256, then 1.2*max)
    truncates rows in target order here."""
    src, tgt = demo_clouds(color)
    p = demo_params(src, tgt, color)
    g = u.CvoGPU(p)
    fails = _teacher_forced(g, p, src, tgt, [0, 1, 2, 10, 50, 100, 300, 1000, 2000, 4000, 6500])
    g.close()
    assert not fails, fails


def test_teacher_forced_colour_and_semantics_with_geotype():
    """Config-5 flavour: 5-dim colour + 20-class semantic kernel + geometric type, runtime C=20."""
    src, tgt, _ = synthetic_pair(4000, 3000, 3200, 20006, F=5, C=20, geotype=True)
    p = u.read_params_yaml(os.path.join(DATA, "cvo_semantic_params_img_gpu0.yaml"))
    p.is_using_geometric_type = 1
    p.c_ell = 0.3
    p.ell_init = 0.9
    g = u.CvoGPU(p)
    fails = _teacher_forced(g, p, src, tgt, [0, 1, 2, 7, 20, 60], max_iter=61)
    g.close()
    assert not fails, fails


@pytest.mark.parametrize("name", ["demo_color_iters.json", "demo_geometric_iters.json",
                                  "synthetic_2k_iters.json"])
def test_committed_golden_iterations(name):
    """Golden fixtures (tests/golden, generated by make_golden.py from the oracle): no oracle call."""
    from golden.make_golden import CASES, case_inputs
    with open(os.path.join(os.path.dirname(__file__), "golden", name)) as fh:
        gold = json.load(fh)
    src, tgt, p = case_inputs(CASES[gold["case"]])
    g = u.CvoGPU(p)
    g.set_cloud(0, src)
    g.set_cloud(1, tgt)
    for rec in gold["iters"]:
        got = g.iterate(np.array(rec["R"], np.float32).reshape(3, 3).T, rec["T"], rec["ell"], rec["cap"])
        assert got.nnz == rec["nnz"] and got.max_row_nnz == rec["max_row_nnz"], rec["k"]
        assert rel_err(list(got.omega) + list(got.v), rec["twist"]) <= TWIST_TOL, rec["k"]
        np.testing.assert_allclose([got.B, got.C, got.D, got.E], rec["BCDE"], rtol=1e-4,
                                   atol=1e-9 * max(abs(x) for x in rec["BCDE"]))
        assert got.step == pytest.approx(rec["step"], rel=1e-4)
        np.testing.assert_allclose(list(got.R), rec["R_next"], atol=1e-6)
        np.testing.assert_allclose(list(got.T), rec["T_next"], atol=1e-5)
    g.close()


@pytest.fixture(scope="session")
def oracle_final_pose_reference():
    """The oracle's final pose of the 2k synthetic registration and ITS OWN sensitivity: the same
    oracle started from initial poses moved by 1e-7 m stops after 354..613 iterations and its
    final poses differ by up to 2.7e-3 (z, +1e-7) - the loop is chaotic (test_oracle.py), so
    north_star's 1e-3 on the final pose is only meaningful where the oracle itself is that
    stable.  The bar used below is max(1e-3, this spread)."""
    src, tgt, Tgt = synthetic_pair(2500, 2000, 2000, 20002)
    p = geometric_params()
    cs, ct = to_oracle_cloud(src), to_oracle_cloud(tgt)
    base = oracle.align(p, cs, ct, None, trace_cap=16)
    spread = 0.0
    for ax, eps in ((0, 1e-7), (1, 3e-7), (2, 1e-7)):
        Ti = np.eye(4, dtype=np.float32)
        Ti[ax, 3] = eps
        _, Tp, _, _ = oracle.align(p, cs, ct, Ti)
        spread = max(spread, float(np.abs(Tp - base[1]).max()))
    return src, tgt, Tgt, p, base, spread


def test_align_final_pose_and_short_trajectory(gpu_geo, oracle_final_pose_reference):
    """End to end through cvo_b200_align: the first iterations track the oracle within tolerance
    (before chaos sets in) and the final pose agrees to 1e-3 or, where the oracle's own
    sensitivity to a 1e-7 m change of the initial pose is larger than that, to that spread."""
    src, tgt, Tgt, p, (r2, T2, i2, tr2), spread = oracle_final_pose_reference
    gpu_geo.write_params(p)
    ret, T, info, tr = gpu_geo.align(src, tgt, None, trace_cap=16)
    assert ret == r2 == 0
    assert info.stop_reason == i2.stop_reason == u._abi.STOP_DIST_SMALL
    for k in range(6):
        assert not compare_traces(tr[k], tr2[k], twist_tol=TWIST_TOL), k
        assert tr[k].num_neighbors == tr2[k].num_neighbors and tr[k].ell == tr2[k].ell
    assert np.abs(T - T2).max() <= max(POSE_TOL, spread), (np.abs(T - T2).max(), spread)
    assert np.abs(T - Tgt).max() < 0.02
    assert info.pairs_tested == 2000 * 2000 * (info.iterations + 1)
    assert info.registration_seconds > 0


def test_align_host_equals_resident_align(gpu_geo):
    src, tgt, _ = synthetic_pair(1500, 1000, 1200, 77)
    p = geometric_params()
    p.MAX_ITER = 25
    gpu_geo.write_params(p)
    r1, T1, i1 = gpu_geo.align(src, tgt)
    r2, T2, i2 = gpu_geo.align_host(src, tgt)
    r3, T3, i3 = gpu_geo.align(src, tgt, resident=True)
    assert r1 == r2 == r3 and i1.iterations == i2.iterations == i3.iterations == 25
    assert np.array_equal(T1, T2) and np.array_equal(T1, T3)  # deterministic reductions
    assert i2.upload_seconds > 0


def test_align_with_initial_guess_and_demo_final_pose():
    src, tgt = demo_clouds(True)
    p = demo_params(src, tgt, True)
    g = u.CvoGPU(p)
    ret, T, info = g.align(src, tgt)
    with open(os.path.join(os.path.dirname(__file__), "golden", "demo_color_align.json")) as fh:
        gold = json.load(fh)
    assert ret == gold["ret"]
    # ~7000 chaotic iterations on a flat optimum (523 vs 1080 points of a partial scene).  The
    # oracle's OWN final pose moves by `spread` = 2.8e-3 when its initial pose moves by 1e-7 m
    # (six starts, tests/golden/make_golden.py::demo_final_pose_spread), so north_star's 1e-3 is
    # not defined on this problem; the GPU run is held to three times that spread (the maximum of
    # six samples is not a bound), 5x tighter than round 1's 4e-2.  The tight per-iteration bar is
    # teacher-forced.
    spread = float(gold["oracle_spread"]["spread"])
    dev = float(np.abs(T - np.array(gold["transform"])).max())
    print(f"demo final pose: |T_gpu - T_oracle| = {dev:.2e} after {info.iterations} iterations "
          f"(oracle: {gold['iterations']}; its own spread {spread:.2e})")
    assert 1e-3 < spread < 5e-3
    assert dev <= 3.0 * spread, (dev, spread)
    # restart from the solution (the ell schedule starts over, so it runs as long): it stays in the
    # same flat optimum - a macroscopically different start, so not the 1e-7 spread above
    # (measured 5e-4 .. 1.2e-2 over the six modes)
    ret2, T2, info2 = g.align(src, tgt, np.linalg.inv(T))
    dev2 = float(np.abs(T2 - T).max())
    print(f"restart from the solution: moved by {dev2:.2e} in {info2.iterations} iterations")
    assert dev2 <= 4e-2
    g.close()


def test_edge_cases_empty_identical_and_zero_cap(gpu_geo):
    src, tgt, _ = synthetic_pair(300, 200, 200, 4)
    empty = u.CvoPointCloud(np.zeros((0, 3), np.float32))
    p = geometric_params(0.3)
    gpu_geo.write_params(p)
    ret, T, info = gpu_geo.align(empty, tgt)        # CvoGPU.cu:1614-1617
    assert ret == 0 and info.iterations == 0
    assert gpu_geo.function_angle(empty, tgt, np.eye(4), 0.5) == 0.0
    g = np.stack(np.meshgrid(*[np.arange(6, dtype=np.float32) * 2.0 + 1.0] * 3, indexing="ij"), -1)
    grid = u.CvoPointCloud(g.reshape(-1, 3))
    ret, T, info, tr = gpu_geo.align(grid, grid, None, trace_cap=2)   # gradient vanishes -> -1
    assert ret == -1 and info.iterations == 0
    assert info.stop_reason == (u._abi.STOP_GRAD_SMALL | u._abi.STOP_GRAD_ZERO)
    assert tr[0].nnz == 216 and np.array_equal(T, np.eye(4, dtype=np.float32))
    gpu_geo.set_cloud(0, src)
    gpu_geo.set_cloud(1, tgt)
    t0 = gpu_geo.iterate(np.eye(3), np.zeros(3), 0.95, 0)          # cap can reach 0 (int(1.2*0))
    assert t0.nnz == 0 and list(t0.omega) == [0, 0, 0] and t0.step == pytest.approx(p.max_step)
    pz = geometric_params()
    pz.is_using_geometric_type = 1                                   # all-zero types: NaN similarity
    gz = u.CvoGPU(pz)
    gz.set_cloud(0, u.CvoPointCloud(src.positions_))
    gz.set_cloud(1, u.CvoPointCloud(tgt.positions_))
    assert gz.iterate(np.eye(3), np.zeros(3), 0.95, 64).nnz == 0
    gz.close()


def test_ragged_sizes_and_row_cap_sweep(gpu_geo):
    """Sizes that are not multiples of the 64-row tile / 256-target block, caps from 1 up."""
    p = geometric_params()
    gpu_geo.write_params(p)
    for (P, N, M, seed) in [(90, 1, 77, 1), (400, 65, 257, 2), (900, 63, 511, 3), (1500, 1000, 1, 4),
                            (2100, 777, 1301, 5)]:
        src, tgt, _ = synthetic_pair(P, N, M, seed)
        gpu_geo.set_cloud(0, src)
        gpu_geo.set_cloud(1, tgt)
        for ell, cap in [(0.95, 256), (3.0, 1), (3.0, 5), (6.0, 33)]:
            ref = oracle.iterate(p, to_oracle_cloud(src), to_oracle_cloud(tgt), np.eye(3).reshape(9),
                                 np.zeros(3), ell, cap)
            got = gpu_geo.iterate(np.eye(3), np.zeros(3), ell, cap)
            assert not compare_traces(got, ref, twist_tol=TWIST_TOL), (N, M, ell, cap)


def test_inner_product_function_angle_and_association(gpu_geo):
    src, tgt, Tgt = synthetic_pair(1800, 1200, 1500, 9, F=5)
    p = u.read_params_yaml(os.path.join(DATA, "cvo_intensity_params_img_gpu0.yaml"))
    p.c_ell = 0.4
    g = u.CvoGPU(p)
    cs, ct = to_oracle_cloud(src), to_oracle_cloud(tgt)
    X = np.linalg.inv(Tgt).astype(np.float32)
    for T in (np.eye(4, dtype=np.float32), X):
        ip_ref, sp = oracle.inner_product(p, cs, ct, T, 0.7, want_matrix=True)
        assert g.inner_product_gpu(src, tgt, T, 0.7) == pytest.approx(ip_ref, rel=2e-5)
        assert g.function_angle(src, tgt, T, 0.7, True) == pytest.approx(
            oracle.function_angle(p, cs, ct, T, 0.7, True), rel=2e-5)
        assoc = g.compute_association_gpu(src, tgt, T, 0.7)
        row_ptr, cols, vals = oracle.sparse_to_csr(sp)
        assert np.array_equal(assoc.row_ptr, row_ptr) and np.array_equal(assoc.cols, cols)
        np.testing.assert_allclose(assoc.vals, vals, rtol=1e-6)
        assert assoc.source_inliers == [int(i) for i in np.nonzero(np.diff(row_ptr))[0]]
    assert g.function_angle(src, tgt, X, 0.7, False) == pytest.approx(
        oracle.function_angle(p, cs, ct, X, 0.7, False), rel=5e-5)
    # anisotropic (Mahalanobis) kernel, CvoGPU.cu:1975-1995
    K = np.diag([0.3, 0.5, 0.8]).astype(np.float32)
    _, spk = oracle.inner_product(p, cs, ct, X, 0.0, kernel3x3=K, want_matrix=True)
    assoc = g.compute_association_gpu(src, tgt, X, K)
    row_ptr, cols, vals = oracle.sparse_to_csr(spk)
    assert np.array_equal(assoc.row_ptr, row_ptr) and np.array_equal(assoc.cols, cols)
    np.testing.assert_allclose(assoc.vals, vals, rtol=1e-6)
    g.close()


def test_align_exports_the_last_iterations_association():
    """is_exporting_association (CvoGPU.cu:1552-1556): align leaves the kernel matrix of its LAST
    iteration; compared with the oracle's matrix at that iteration's state (teacher-forced from
    the GPU's own trace, so no trajectory divergence is involved)."""
    src, tgt, _ = synthetic_pair(2500, 1500, 1800, 77, F=5)
    p = u.read_params_yaml(os.path.join(DATA, "cvo_intensity_params_img_gpu0.yaml"))
    p.ell_init, p.c_ell, p.is_exporting_association = 0.9, 0.3, 1
    cs, ct = to_oracle_cloud(src), to_oracle_cloud(tgt)
    for max_iter in (1, 4):
        p.MAX_ITER = max_iter
        g = u.CvoGPU(p)
        assoc = u.Association()
        ret, T, info, tr = g.align(src, tgt, None, association=assoc, trace_cap=max_iter)
        assert ret == 0 and len(tr) == max_iter
        last = tr[-1]
        if max_iter == 1:
            R, t = np.eye(3, dtype=np.float32).reshape(9), np.zeros(3, np.float32)
        else:
            R, t = np.array(list(tr[-2].R), np.float32), np.array(list(tr[-2].T), np.float32)
        ref, sp = oracle.iterate(p, cs, ct, R, t, last.ell, last.num_neighbors, want_matrix=True)
        assert ref.nnz == last.nnz > 0
        row_ptr, cols, vals = oracle.sparse_to_csr(sp)
        assert assoc.shape == (1500, 1800)
        assert np.array_equal(assoc.row_ptr, row_ptr) and np.array_equal(assoc.cols, cols)
        np.testing.assert_allclose(assoc.vals, vals, rtol=1e-6)
        # any other call overwrites the matrix: the export must then refuse, not return garbage
        g.inner_product_gpu(src, tgt, np.eye(4, dtype=np.float32), 0.5)
        with pytest.raises(u.CvoError):
            g._fill_association(u.Association(), 1500, 1800,
                                lambda *a: g._lib.cvo_b200_align_association(g._h, *a))
        g.close()


@pytest.mark.parametrize("name,ell", [("KITTI05", 1.5), ("KITTI05", 0.15), ("C4", 0.5)])
def test_full_size_single_iterations_against_the_oracle(name, ell):
    """BASELINE's full sizes (KITTI-05-sized 16 384 x 16 384 and C4 = 200 000 x 200 000, both
    with 5-dim colour), one iteration at the identity and one at a displaced pose: exact nnz and
    max row count, twist to 1e-4.  The oracle finishes in seconds at these sizes because its rows
    enumerate candidates through a grid (bit-identical to its dense loop, test_oracle.py)."""
    from unified_cvo_b200 import synthetic
    d = synthetic.make_config(name)
    src = u.CvoPointCloud(d["source"]["xyz"], d["source"]["features"], None, None)
    tgt = u.CvoPointCloud(d["target"]["xyz"], d["target"]["features"], None, None)
    p = u.read_params_yaml(os.path.join(DATA, "cvo_intensity_params_img_gpu0.yaml"))
    g = u.CvoGPU(p)
    g.set_cloud(0, src)
    g.set_cloud(1, tgt)
    cs, ct = to_oracle_cloud(src), to_oracle_cloud(tgt)
    # a pose 5 % off the true one: T_target_to_source = inverse of the true motion (synthetic.py)
    a = np.deg2rad(2.0 * 0.95)
    Gn = np.eye(4)
    Gn[:3, :3] = [[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]
    Tr = np.eye(4)
    Tr[:3, 3] = np.array([0.05, 0.02, 0.50]) * 0.95
    near = np.linalg.inv(Gn @ Tr).astype(np.float32)
    total = 0
    for R, T in ((np.eye(3, dtype=np.float32), np.zeros(3, np.float32)),
                 (near[:3, :3].copy(), near[:3, 3].copy())):
        ref = oracle.iterate(p, cs, ct, R.T.reshape(9).copy(), T, ell, 256)  # column-major R
        got = g.iterate(R, T, ell, 256)
        total += ref.nnz
        if ref.nnz == 0:  # nothing within reach at this pose (ell = 0.15 against a 0.5 m offset)
            assert got.nnz == 0 and got.max_row_nnz == 0
            continue
        bad = compare_traces(got, ref, twist_tol=TWIST_TOL)
        assert not bad, (name, ell, bad)
    assert total > 1000
    g.close()


def test_full_size_c5_colour_and_20_class_semantics():
    """BASELINE configs[4] size (TUM-fr1-sized N=M=12 800, 5-dim colour + 20-class semantic kernel,
    C a run-time value): single iterations at the identity and near the true pose against the
    oracle.  Parameters: cvo_semantic_params_img_gpu0.yaml — with cvo_rgbd_params.yaml's
    sigma/c_sigma and semantics switched on, sigma^2 c_sigma^2 s_sigma^2 barely exceeds sp_thres
    and the oracle itself keeps no pair on these clouds."""
    from unified_cvo_b200 import synthetic
    d = synthetic.make_config("C5")
    src = u.CvoPointCloud(d["source"]["xyz"], d["source"]["features"], d["source"]["labels"], None)
    tgt = u.CvoPointCloud(d["target"]["xyz"], d["target"]["features"], d["target"]["labels"], None)
    assert src.num_classes() == 20 and src.feature_dimensions() == 5 and src.num_points() == 12_800
    p = u.read_params_yaml(os.path.join(DATA, "cvo_semantic_params_img_gpu0.yaml"))
    g = u.CvoGPU(p)
    g.set_cloud(0, src)
    g.set_cloud(1, tgt)
    cs, ct = to_oracle_cloud(src), to_oracle_cloud(tgt)
    a = np.deg2rad(2.0 * 0.95)
    Gn = np.eye(4)
    Gn[:3, :3] = [[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]
    Tr = np.eye(4)
    Tr[:3, 3] = np.array([0.05, 0.02, 0.50]) * 0.95
    near = np.linalg.inv(Gn @ Tr).astype(np.float32)
    cap = int(p.nearest_neighbors_max)
    for R, T, ell in ((np.eye(3, dtype=np.float32), np.zeros(3, np.float32), 1.0),
                      (near[:3, :3].copy(), near[:3, 3].copy(), 0.15),
                      (near[:3, :3].copy(), near[:3, 3].copy(), 0.5)):
        ref = oracle.iterate(p, cs, ct, R.T.reshape(9).copy(), T, ell, cap)
        got = g.iterate(R, T, ell, cap)
        assert ref.nnz > 4000
        bad = compare_traces(got, ref, twist_tol=TWIST_TOL)
        assert not bad, (ell, bad)
    g.close()


def test_full_size_properties_kitti_sized_colour():
    """KITTI-05-sized clouds (N=M=16384, F=5): properties that need no oracle at this size —
    source-row shards sum to the whole (the multi-GPU decomposition), determinism, and the
    association's rows are strictly ascending and capped."""
    src, tgt, _ = synthetic_pair(20_480, 16_384, 16_384, 20_005, F=5)
    p = u.read_params_yaml(os.path.join(DATA, "cvo_intensity_params_img_gpu0.yaml"))
    p.c_ell = 0.3
    g = u.CvoGPU(p)
    g.set_cloud(0, src)
    g.set_cloud(1, tgt)
    R, T, ell, cap = np.eye(3), np.zeros(3), 0.6, 40
    full = g.iterate(R, T, ell, cap)
    again = g.iterate(R, T, ell, cap)
    assert bytes(full) == bytes(again)  # bit-deterministic
    parts = []
    for (b, e) in [(0, 5000), (5000, 5001), (5001, 16_384)]:
        g.set_row_range(b, e)
        parts.append(g.iterate(R, T, ell, cap))
    g.set_row_range(0, -1)
    assert sum(t.nnz for t in parts) == full.nnz and full.nnz > 1_000
    assert max(t.max_row_nnz for t in parts) == full.max_row_nnz <= cap
    for name in ("omega_sum", "v_sum"):
        s = np.sum([list(getattr(t, name)) for t in parts], axis=0)
        np.testing.assert_allclose(s, list(getattr(full, name)), rtol=1e-6, atol=1e-9)  # shards with unaligned row ranges use the original target order: float sums in another order
    assert sum(t.a_sum for t in parts) == pytest.approx(full.a_sum, rel=1e-9)
    assoc = g.compute_association_gpu(src, tgt, np.eye(4), ell)
    counts = np.diff(assoc.row_ptr)
    assert counts.max() <= p.nearest_neighbors_max
    for i in np.argsort(counts)[-50:]:
        js = assoc.cols[assoc.row_ptr[i]:assoc.row_ptr[i + 1]]
        assert np.all(np.diff(js) > 0)
    assert np.all(assoc.vals > p.sp_thres)
    g.close()
