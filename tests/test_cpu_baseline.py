"""The restated reference CPU path (oracle/cvo_cpu_baseline.c = cvo::cvo::align, Cvo.cpp:885-1089),
the CPU arm bench.py times beside the GPU path.  It is a timing baseline, not a parity oracle
(different algorithm from CvoGPU: no row cap, no gradient normalisation); these tests check that
it is a working registration and that its kd-tree search is exact."""
import numpy as np

import oracle
from helpers import demo_clouds, demo_params, geometric_params, synthetic_pair, to_oracle_cloud


def test_kdtree_radius_search_equals_brute_force():
    rng = np.random.default_rng(5)
    pts = rng.uniform(-5, 5, (3000, 3)).astype(np.float32)
    pts[100:140] = pts[100]  # duplicates: degenerate splits
    q = rng.uniform(-5, 5, (200, 3)).astype(np.float32)
    for r2 in (0.25, 1.0, 9.0):
        got = oracle.cpu_baseline_radius_counts(pts, q, r2)
        d = q[:, None, :] - pts[None, :, :]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        assert np.array_equal(got, (d2 < np.float32(r2)).sum(1))


def test_baseline_registers_a_synthetic_pair_to_ground_truth():
    src, tgt, Tgt = synthetic_pair(2500, 2000, 2000, 20002)
    p = geometric_params()
    ret, T, info = oracle.cpu_baseline_align(p, to_oracle_cloud(src), to_oracle_cloud(tgt))
    assert ret == 0 and 10 < info["iterations"] < p.MAX_ITER
    assert np.abs(T - Tgt).max() < 0.02
    assert info["pairs"] == 2000 * 2000 * info["executed"]
    assert info["seconds"] >= info["t_se_kernel"] > 0
    # and agrees with the GPU-path semantics (the oracle) on where the optimum is
    _, T2, _, _ = oracle.align(p, to_oracle_cloud(src), to_oracle_cloud(tgt))
    assert np.abs(T - T2).max() < 0.02


def test_baseline_edge_cases():
    src, tgt, _ = synthetic_pair(300, 200, 200, 4)
    p = geometric_params()
    empty = oracle.Cloud(np.zeros((0, 3), np.float32))
    ret, T, info = oracle.cpu_baseline_align(p, empty, to_oracle_cloud(tgt))
    assert ret == 0 and np.array_equal(T, np.eye(4, dtype=np.float32)) and info["executed"] == 0
    cs = to_oracle_cloud(src)
    ret, T, info = oracle.cpu_baseline_align(p, cs, cs)  # identical clouds: the gradient vanishes
    assert ret == -1 and info["iterations"] == 0


def test_baseline_runs_the_demo_with_colour():
    src, tgt = demo_clouds(True)
    p = demo_params(src, tgt, True)
    p.MAX_ITER = 300
    ret, T, info = oracle.cpu_baseline_align(p, to_oracle_cloud(src), to_oracle_cloud(tgt))
    assert ret == 0 and info["executed"] == 300 and info["nnz_last"] > 0
    R = T[:3, :3].astype(np.float64)
    assert np.abs(R @ R.T - np.eye(3)).max() < 1e-4
