"""The oracle pinned against the REFERENCE'S OWN kernels (CPU half).

oracle/make_ref.py compiles the reference's text of fill_in_A_mat_gpu (CvoGPU.cu:477-593) and its
helpers — no Eigen, no PCL, no thrust in it — with g++ into oracle/_ref/libcvo_ref_host.so (the
grid becomes a host loop).  Tier 1 below asserts oracle_fill_A ≡ that kernel BIT FOR BIT (stored
values, column indices, per-row counts) in the oracle's "as written" arithmetic.  Tier 2 does the
same for K1b / K2 / K3+K4, whose Eigen 3-vector primitives come from oracle/ref_mini_eigen.h (ours).
The B200 half (nvcc builds with the reference's own flags, the CUDA product against them) is
tests/test_ref_pin_gpu.py.
"""
import os

import numpy as np
import pytest

import oracle
from oracle import ref
import unified_cvo_b200 as u
from helpers import DATA, demo_clouds, demo_params, geometric_params, synthetic_pair, to_oracle_cloud

pytestmark = pytest.mark.skipif(not ref.available("host"),
                                reason="oracle/_ref not built (python oracle/make_ref.py needs /root/reference)")


@pytest.fixture(autouse=True)
def as_written_arithmetic():
    """g++ does not contract on x86-64: compare with the oracle's uncontracted mode."""
    oracle.set_device_arith(False)
    yield
    oracle.set_device_arith(True)


def assert_same_matrix(got, want, what=""):
    """bit for bit: counts, indices (incl. the -1 fill), values (incl. the zero fill)"""
    assert np.array_equal(got["nonzeros"], want["nonzeros"]), f"{what}: per-row counts differ"
    assert np.array_equal(got["ind"], want["ind"]), f"{what}: column indices differ"
    assert np.array_equal(got["mat"].view(np.uint32), want["mat"].view(np.uint32)), f"{what}: values differ"


def rot_z(deg):
    a = np.deg2rad(deg)
    return np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], np.float32)


@pytest.mark.parametrize("color", [True, False])
def test_demo_pcds_at_the_drivers_ell_init_saturating_rows(color):
    """config 1: demo_data at ell = |mean(src) - mean(tgt)| = 5.76 (two_color_pcd.cpp:56-60),
    cap 256: rows saturate (the densest has 507 geometric survivors)."""
    src, tgt = demo_clouds(color=color)
    p = demo_params(src, tgt, color=color)
    cs, ct = to_oracle_cloud(src), to_oracle_cloud(tgt)
    cap = int(p.nearest_neighbors_max)
    want = ref.fill_A(p, cs, ct, ct.xyz, cap, float(p.ell_init))
    got = oracle.fill_A(p, cs, ct, ct.xyz, cap, float(p.ell_init))
    assert_same_matrix(got, want, "demo")
    assert want["nonzero_sum"] > 5000
    if not color:
        assert int(want["nonzeros"].max()) == cap  # the truncation rule is exercised
    # the literal dense loop of the oracle as well (the grid-accelerated enumeration is the default)
    oracle.set_accel(False)
    try:
        assert_same_matrix(oracle.fill_A(p, cs, ct, ct.xyz, cap, float(p.ell_init)), want, "demo/dense")
    finally:
        oracle.set_accel(True)


@pytest.mark.parametrize("ell,cap", [(0.95, 256), (0.3, 256), (0.95, 5), (2.0, 1)])
def test_c2_sized_geometric(ell, cap):
    """config 2: N = M = 10 000 synthetic, geometric kernel, a moved target."""
    src, tgt, _ = synthetic_pair(12500, 10000, 10000, 20002)
    p = geometric_params()
    cs, ct = to_oracle_cloud(src), to_oracle_cloud(tgt)
    ym = oracle.transform(rot_z(1.0), [0.02, -0.01, 0.3], ct.xyz)
    want = ref.fill_A(p, cs, ct, ym, cap, ell)
    got = oracle.fill_A(p, cs, ct, ym, cap, ell)
    assert_same_matrix(got, want, f"C2 ell={ell} cap={cap}")
    assert want["nonzero_sum"] > 1000


def test_colour_semantics_geotype_moved_target():
    """all four factors of the kernel: 5-dim colour, 19-class semantics (the width the reference is
    compiled for), geometric types; the shipped semantic KITTI parameter set."""
    src, tgt, _ = synthetic_pair(4000, 3000, 3300, 31, F=5, C=19, geotype=True)
    p = u.read_params_yaml(os.path.join(DATA, "cvo_semantic_params_img_gpu0.yaml"))
    p.is_using_geometric_type = 1
    p.c_ell = 1.0      # random colours and labels: loosen the kernels so that pairs survive them
    p.sp_thres = 0.001
    cs, ct = to_oracle_cloud(src), to_oracle_cloud(tgt)
    ym = oracle.transform(rot_z(-2.0), [0.05, 0.02, 0.5], ct.xyz)
    for ell, cap in ((1.2, 512), (0.6, 9)):
        want = ref.fill_A(p, cs, ct, ym, cap, ell)
        assert_same_matrix(oracle.fill_A(p, cs, ct, ym, cap, ell), want, f"full kernel ell={ell}")
        assert want["nonzero_sum"] > 200
        assert int(want["nonzeros"].max()) <= cap


def test_rgbd_parameter_set_with_semantics_stores_nothing():
    """BASELINE config 5 verbatim: cvo_rgbd_params.yaml + is_using_semantics=1.  sigma^2 * c_sigma^2
    * s_sigma^2 = 0.01 * 0.36 * 0.64 = 0.0023 < sp_thres = 0.003: no pair can pass a > sp_thres."""
    src, tgt, _ = synthetic_pair(3000, 2400, 2400, 20006, F=5, C=19)
    p = u.read_params_yaml(os.path.join(DATA, "cvo_rgbd_params.yaml"))
    p.is_using_semantics = 1
    cs, ct = to_oracle_cloud(src), to_oracle_cloud(tgt)
    want = ref.fill_A(p, cs, ct, ct.xyz, int(p.nearest_neighbors_max), float(p.ell_init))
    got = oracle.fill_A(p, cs, ct, ct.xyz, int(p.nearest_neighbors_max), float(p.ell_init))
    assert_same_matrix(got, want, "config 5")
    assert want["nonzero_sum"] == 0


def test_all_zero_geometric_types_are_nan_and_dropped():
    """clouds built by reserve/add_point have zero geometric types: geo_sim = 0/0 = NaN, the pair is
    not skipped (NaN < 0.01 is false) and not stored (a = NaN > sp_thres is false)."""
    src, tgt, _ = synthetic_pair(700, 500, 600, 3)
    p = geometric_params()
    p.is_using_geometric_type = 1
    cs = oracle.Cloud(src.positions_, geotype=np.zeros((500, 2), np.float32))
    gt = np.zeros((600, 2), np.float32)
    gt[::2] = (0.0, 1.0)  # every other target has a type: still NaN through the source's zeros
    ct = oracle.Cloud(tgt.positions_, geotype=gt)
    want = ref.fill_A(p, cs, ct, ct.xyz, 64, 0.95)
    assert_same_matrix(oracle.fill_A(p, cs, ct, ct.xyz, 64, 0.95), want, "NaN geo type")
    assert want["nonzero_sum"] == 0
    # and a mixed case where only some rows are affected
    gs = np.tile(np.array([[1.0, 0.0]], np.float32), (500, 1))
    gs[100:200] = 0.0
    gs[300:] = (0.0, 1.0)
    cs2 = oracle.Cloud(src.positions_, geotype=gs)
    want = ref.fill_A(p, cs2, ct, ct.xyz, 64, 0.95)
    assert_same_matrix(oracle.fill_A(p, cs2, ct, ct.xyz, 64, 0.95), want, "mixed geo types")
    assert 0 < want["nonzero_sum"]
    assert want["nonzeros"][100:200].sum() == 0


def test_ragged_and_degenerate_shapes():
    src, tgt, _ = synthetic_pair(700, 513, 1, 9)  # 513 rows: one thread past a 512 block; one target
    p = geometric_params()
    cs, ct = to_oracle_cloud(src), to_oracle_cloud(tgt)
    for cap in (1, 4):
        want = ref.fill_A(p, cs, ct, ct.xyz, cap, 50.0)  # ell so large that every row keeps the target
        assert_same_matrix(oracle.fill_A(p, cs, ct, ct.xyz, cap, 50.0), want, "one target")
        assert want["nonzero_sum"] == 513
    # geometry switched off: every pair passes, rows hold the first `cap` targets with a = sigma^2... = 1
    src, tgt, _ = synthetic_pair(200, 64, 90, 10)
    p.is_using_geometry = 0
    p.sp_thres = 0.5
    cs, ct = to_oracle_cloud(src), to_oracle_cloud(tgt)
    want = ref.fill_A(p, cs, ct, ct.xyz, 7, 0.5)
    assert_same_matrix(oracle.fill_A(p, cs, ct, ct.xyz, 7, 0.5), want, "no geometry")
    assert np.array_equal(want["ind"][0], np.arange(7))


# ------------------------------------------------------------------ tier 2 (Eigen stand-in)
def test_dense_kernel_variant_k1b():
    """fill_in_A_mat_gpu_dense_mat_kernel (CvoGPU.cu:217-327) with an anisotropic kernel."""
    src, tgt, _ = synthetic_pair(1500, 1000, 1200, 77, F=5)
    p = u.read_params_yaml(os.path.join(DATA, "cvo_intensity_params_img_gpu0.yaml"))
    p.c_ell = 0.6
    cs, ct = to_oracle_cloud(src), to_oracle_cloud(tgt)
    K = np.array([[0.30, 0.02, 0.00], [0.02, 0.20, 0.01], [0.00, 0.01, 0.50]], np.float32)
    Kinv = np.linalg.inv(K.astype(np.float64)).astype(np.float32)
    for cap in (256, 6):
        want = ref.fill_A(p, cs, ct, ct.xyz, cap, 0.0, kernel_inv=Kinv)
        got = oracle.fill_A(p, cs, ct, ct.xyz, cap, 0.0, kernel_inv=Kinv)
        assert_same_matrix(got, want, f"K1b cap={cap}")
        assert want["nonzero_sum"] > 100


@pytest.mark.parametrize("range_ell", [0, 1])
def test_flow_and_step_rows_k2_k3_k4(range_ell):
    """compute_flow_gpu_no_eigen and compute_step_size_xi/_poly_coeff on a matrix produced by K1:
    per-row outputs equal bit for bit (same formulas, same float/double mix, same order)."""
    src, tgt, _ = synthetic_pair(3000, 2000, 2500, 123)
    p = geometric_params()
    p.is_using_range_ell = range_ell
    cs, ct = to_oracle_cloud(src), to_oracle_cloud(tgt)
    ym = oracle.transform(rot_z(0.7), [0.01, 0.03, 0.2], ct.xyz)
    ell, cap = 0.8, 40
    A = ref.fill_A(p, cs, ct, ym, cap, ell)
    om_w, v_w = ref.flow_rows(p, cs.xyz, ym, A)
    om_g, v_g = oracle.flow_rows(p, cs, ym, A)
    assert np.array_equal(om_g, om_w) and np.array_equal(v_g, v_w)
    assert np.abs(om_w).sum() > 0
    # a unit twist, as compute_flow hands it on
    tw = np.concatenate([om_w.sum(0), v_w.sum(0)]).astype(np.float32)
    tw /= np.linalg.norm(tw)
    want = ref.step_rows(tw[:3], tw[3:], ell, float(p.ell_init), range_ell, cs.xyz, ym, A)
    got = oracle.step_rows(p, cs, ym, A, tw[:3], tw[3:], ell)
    assert np.array_equal(got, want)
    assert np.abs(want).sum() > 0


def test_exp_sek3_pose_increment_matches_the_reference_text():
    """a11: oracle_exp_sek3 == the reference's Exp_SEK3 (LieGroup.cpp:245-274, called at
    CvoGPU.cu:1462) compiled from its own text over the mini-Eigen, bit for bit: the small-angle
    branch (theta < 1e-6), unit twists at the step sizes the loop takes, large rotations."""
    rng = np.random.default_rng(20011)
    cases = []
    for _ in range(300):  # what align_impl passes: the jointly normalised twist, a small step
        xi = rng.normal(size=6).astype(np.float32)
        xi /= np.float32(np.linalg.norm(xi))
        cases.append((xi, float(np.float32(10.0 ** rng.uniform(-6, -0.1)))))
    for _ in range(100):  # arbitrary twists and steps, up to several turns
        cases.append((rng.normal(scale=3.0, size=6).astype(np.float32), float(np.float32(rng.uniform(0, 2.5)))))
    for scale in (0.0, 1e-9, 5e-7, 0.99e-6, 1.01e-6, 1e-5):  # around TOLERANCE
        xi = np.array([scale, 0, 0, 0.3, -0.2, 0.9], np.float32)
        cases.append((xi, 0.8))
        xi = np.concatenate([(rng.normal(size=3) * scale).astype(np.float32), rng.normal(size=3).astype(np.float32)])
        cases.append((xi, 0.37))
    worst = 0
    for xi, dt in cases:
        got = oracle.exp_sek3(xi, dt)
        want = ref.exp_sek3(xi, dt)
        assert got.shape == want.shape == (3, 4)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (xi, dt, got, want)
        # and it is a rigid motion to float accuracy (not a tautology of two equal mistakes)
        R = want[:, :3].astype(np.float64)
        worst = max(worst, float(np.abs(R @ R.T - np.eye(3)).max()))
    assert worst < 5e-6


@pytest.mark.parametrize("window", [1, 2, 3, 10, 15, 50])
def test_indicator_queues_match_the_reference_text(window):
    """a12: the oracle's indicator queues == the reference's A_sparsity_indicator_ell_update
    (CvoGPU.cu:1167-1285) compiled from its own text (std::queue only, tier 1) - decisions and both
    running float sums after every call, bit for bit, over sequences that plateau (decays fire and
    clear the queues), drift, jump and contain zeros."""
    rng = np.random.default_rng(7000 + window)
    p = geometric_params()
    p.indicator_window_size = window
    fired = 0
    for thr in (0.001, 0.01, 0.02, 0.2):
        p.indicator_stable_threshold = thr
        n = 40 * window + 60
        plateau = np.full(n, 3.25, np.float32) * (1 + rng.normal(scale=thr / 4, size=n)).astype(np.float32)
        drift = (5.0 * np.exp(-np.arange(n) / (3.0 * window + 5))).astype(np.float32) + np.float32(0.5)
        jumps = plateau.copy()
        jumps[:: 2 * window + 3] *= np.float32(1.0 + 4 * thr)
        sparse = np.where(rng.random(n) < 0.3, 0.0, rng.random(n)).astype(np.float32)
        for seq in (plateau, drift, jumps, sparse, np.concatenate([drift, plateau])):
            got = oracle.indicator_sequence(p, seq)
            want = ref.indicator_sequence(p, seq)
            assert np.array_equal(got[0], want[0])
            for a, b in zip(got[1:], want[1:]):
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
            fired += int(want[0].sum())
    assert fired > 0  # the clear-on-decay branch ran


def test_inverse_pose_and_point_transform_match_the_reference_text():
    """a3 / a4: oracle_update_tf + oracle_transform == the reference's update_tf (CvoGPU.cu:94-112)
    and transform_point_R_T (CvoGPU_impl.cu:31-82) compiled from their own text over the
    mini-Eigen, bit for bit: the inverse pose it uploads, the 4x4 it returns, and the moved target."""
    import ctypes as C

    rng = np.random.default_rng(94112)
    L = oracle.lib()
    f32p = C.POINTER(C.c_float)
    y = np.concatenate([rng.normal(scale=10.0, size=(2000, 3)), rng.normal(scale=1e-3, size=(50, 3)),
                        np.zeros((1, 3))]).astype(np.float32)
    for k in range(40):
        # rotations as the loop leaves them: products of float matrices, not re-orthogonalised
        R = np.eye(3, dtype=np.float32)
        for _ in range(1 + k % 5):
            w = rng.normal(scale=0.3, size=3)
            th = np.linalg.norm(w)
            K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]) / th
            R = (R.astype(np.float64) @ (np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K)).astype(np.float32)
        T = rng.normal(scale=[0.1, 1.0, 20.0][k % 3], size=3).astype(np.float32)
        rinv_w, tinv_w, tf_w, moved_w = ref.update_tf_and_transform(R, T, y)
        Rc = np.ascontiguousarray(R.T).reshape(9)
        rinv, tinv, tf = np.zeros(9, np.float32), np.zeros(3, np.float32), np.zeros(16, np.float32)
        L.oracle_update_tf(Rc.ctypes.data_as(f32p), T.ctypes.data_as(f32p), rinv.ctypes.data_as(f32p),
                           tinv.ctypes.data_as(f32p), tf.ctypes.data_as(f32p))
        assert np.array_equal(rinv.reshape(3, 3).T.view(np.uint32), rinv_w.view(np.uint32))
        assert np.array_equal(tinv.view(np.uint32), tinv_w.view(np.uint32))
        assert np.array_equal(tf.reshape(4, 4).T.view(np.uint32), tf_w.view(np.uint32))
        moved = oracle.transform(R, T, y)
        assert np.array_equal(moved.view(np.uint32), moved_w.view(np.uint32))
        assert np.abs(moved.astype(np.float64) - (y.astype(np.float64) - T) @ R.astype(np.float64)).max() < 1e-3


def test_frame_pose_transform_matches_the_reference_text():
    """N3: oracle_transform_pose_vec == the reference's transform_point_pose_vec
    (CvoGPU_impl.cu:84-150: a row-major 3x4 pose mapped onto [x y z 1]) compiled from its own text
    over the mini-Eigen, bit for bit."""
    rng = np.random.default_rng(84150)
    x = np.concatenate([rng.normal(scale=12.0, size=(3000, 3)), np.zeros((1, 3)),
                        rng.normal(scale=1e-4, size=(20, 3))]).astype(np.float32)
    for k in range(30):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        pose = np.concatenate([q * (1 + 1e-4 * rng.normal()), rng.normal(scale=[0.05, 2.0, 30.0][k % 3], size=(3, 1))],
                              axis=1).astype(np.float32)
        got = oracle.transform_pose_vec(pose, x)
        want = ref.transform_pose_vec(pose, x)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        exact = x.astype(np.float64) @ pose[:, :3].astype(np.float64).T + pose[:, 3].astype(np.float64)
        assert np.abs(want - exact).max() < 1e-4
