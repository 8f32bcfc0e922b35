"""The CUDA product and the oracle pinned against the REFERENCE'S OWN kernels on a B200.

oracle/make_ref.py compiles the reference's text of fill_in_A_mat_gpu (CvoGPU.cu:477-593; no
Eigen/PCL/thrust in it) with nvcc for sm_100a twice: with the reference's own flags (Release,
default --fmad=true: `cuda`) and with --fmad=false (`cuda_nofma`).  Here:
  * the product (cvo_b200_association through the C-ABI: the same pairwise pass every call of the
    hot path runs) ≡ the reference kernel, reference flags: BIT FOR BIT — per-row counts, column
    indices, stored values — at BASELINE's sizes, including C4 = 200 000 x 200 000 with colour at
    ell = 1.5, the benchmarked multi-GPU regime;
  * the oracle in its default (device) arithmetic ≡ the same kernel: counts and indices exact;
    values exact up to the last bit of exp(double) / logf, which are glibc's on the host and
    CUDA's on the device (reported: how many values differ at all);
  * the oracle "as written" ≡ the --fmad=false build likewise.
The reference kernel runs on the already-moved target; the product moves the target itself, so
the non-identity cases also pin the product's update_tf + transform against the oracle's.
"""
import os

import numpy as np
import pytest

import oracle
from oracle import ref
import unified_cvo_b200 as u
from helpers import DATA, demo_clouds, demo_params, geometric_params, synthetic_pair, to_oracle_cloud

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref.available("cuda"), reason="oracle/_ref CUDA build missing")]


def rot_z(deg):
    a = np.deg2rad(deg)
    return np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], np.float32)


def pose4(R, T):
    M = np.eye(4, dtype=np.float32)
    M[:3, :3] = R
    M[:3, 3] = T
    return M


def ell_to_csr(sp):
    nz = sp["nonzeros"].astype(np.int64)
    row_ptr = np.zeros(len(nz) + 1, np.int64)
    np.cumsum(nz, out=row_ptr[1:])
    k = sp["ind"].shape[1]
    mask = np.arange(k)[None, :] < nz[:, None]
    return row_ptr, sp["ind"][mask].astype(np.int32), sp["mat"][mask].astype(np.float32)


def product_matrix(p, src, tgt, T4, ell):
    g = u.CvoGPU(p, device=0)
    try:
        a = g.compute_association_gpu(src, tgt, T4, float(ell))
        assert g.launch_count() > 0
    finally:
        g.close()
    return a.row_ptr, a.cols, a.vals


def assert_product_equals_reference(p, src, tgt, R, T, ell, what, min_nnz=1):
    """cvo_b200_association(pose, ell) vs the reference kernel at cap = nearest_neighbors_max."""
    cs, ct = to_oracle_cloud(src), to_oracle_cloud(tgt)
    ym = oracle.transform(R, T, ct.xyz)
    cap = int(p.nearest_neighbors_max)
    want = ref.fill_A(p, cs, ct, ym, cap, ell, kind="cuda")
    rp_w, c_w, v_w = ell_to_csr(want)
    rp_g, c_g, v_g = product_matrix(p, src, tgt, pose4(R, T), ell)
    assert np.array_equal(rp_g, rp_w), f"{what}: per-row counts differ"
    assert np.array_equal(c_g, c_w), f"{what}: column indices differ"
    assert np.array_equal(v_g.view(np.uint32), v_w.view(np.uint32)), \
        f"{what}: {int((v_g.view(np.uint32) != v_w.view(np.uint32)).sum())} of {len(v_w)} values differ"
    assert want["nonzero_sum"] >= min_nnz, f"{what}: vacuous ({want['nonzero_sum']} entries)"
    return want


I3, Z3 = np.eye(3, dtype=np.float32), np.zeros(3, np.float32)


@pytest.mark.parametrize("color", [True, False])
def test_product_demo_pcds_saturating_rows(color):
    src, tgt = demo_clouds(color=color)
    p = demo_params(src, tgt, color=color)
    want = assert_product_equals_reference(p, src, tgt, I3, Z3, float(p.ell_init), "demo", 5000)
    if not color:
        assert int(want["nonzeros"].max()) == int(p.nearest_neighbors_max)
    assert_product_equals_reference(p, src, tgt, rot_z(3.0), np.array([0.3, -0.2, 0.5], np.float32), 2.0,
                                    "demo moved", 100)


@pytest.mark.parametrize("ell", [0.95, 0.3, 0.1])
def test_product_c2_full_size(ell):
    src, tgt, _ = synthetic_pair(12500, 10000, 10000, 20002)
    p = geometric_params()
    assert_product_equals_reference(p, src, tgt, I3, Z3, ell, f"C2 ell={ell}", 100)
    assert_product_equals_reference(p, src, tgt, rot_z(1.0), np.array([0.02, -0.01, 0.3], np.float32), ell,
                                    f"C2 moved ell={ell}", 100 if ell > 0.2 else 1)


@pytest.mark.parametrize("cap", [512, 9, 1])
def test_product_colour_semantics_geotype(cap):
    src, tgt, _ = synthetic_pair(4000, 3000, 3300, 31, F=5, C=19, geotype=True)
    p = u.read_params_yaml(os.path.join(DATA, "cvo_semantic_params_img_gpu0.yaml"))
    p.is_using_geometric_type = 1
    p.c_ell = 1.0
    p.sp_thres = 0.001
    p.nearest_neighbors_max = cap
    assert_product_equals_reference(p, src, tgt, rot_z(-2.0), np.array([0.05, 0.02, 0.5], np.float32), 1.2,
                                    f"full kernel cap={cap}", 200)


def test_product_nan_geometric_types_and_config5():
    src, tgt, _ = synthetic_pair(700, 500, 600, 3)
    p = geometric_params()
    p.is_using_geometric_type = 1
    gs = np.tile(np.array([[1.0, 0.0]], np.float32), (500, 1))
    gs[100:200] = 0.0
    gs[300:] = (0.0, 1.0)
    gt = np.zeros((600, 2), np.float32)
    gt[::2] = (0.0, 1.0)
    gt[1::4] = (1.0, 0.0)
    s2 = u.CvoPointCloud(src.positions_, None, None, gs)
    t2 = u.CvoPointCloud(tgt.positions_, None, None, gt)
    want = assert_product_equals_reference(p, s2, t2, I3, Z3, 0.95, "NaN geo types", 1)
    assert want["nonzeros"][100:200].sum() == 0
    # BASELINE config 5 verbatim (cvo_rgbd_params.yaml + semantics on): nothing can be stored
    src, tgt, _ = synthetic_pair(3000, 2400, 2400, 20006, F=5, C=19)
    p = u.read_params_yaml(os.path.join(DATA, "cvo_rgbd_params.yaml"))
    p.is_using_semantics = 1
    want = assert_product_equals_reference(p, src, tgt, I3, Z3, float(p.ell_init), "config 5", 0)
    assert want["nonzero_sum"] == 0
    g = u.CvoGPU(p, device=0)
    ret, _, info = g.align(src, tgt)
    g.close()
    ret_o, _, info_o, _ = oracle.align(p, to_oracle_cloud(src), to_oracle_cloud(tgt))
    # the reference's "gradient vanished" return (CvoGPU.cu:1454-1457) on both sides
    assert ret == ret_o == -1 and info.iterations == info_o.iterations == 0


def test_product_kitti_sized_colour():
    src, tgt, _ = synthetic_pair(20480, 16384, 16384, 20005, F=5)
    p = u.read_params_yaml(os.path.join(DATA, "cvo_intensity_params_img_gpu0.yaml"))
    assert_product_equals_reference(p, src, tgt, I3, Z3, float(p.ell_init_first_frame), "KITTI first frame", 1000)
    assert_product_equals_reference(p, src, tgt, rot_z(0.5), np.array([0.01, 0.0, 0.1], np.float32),
                                    0.5, "KITTI tracking", 10)


def test_product_c4_full_size_at_the_benchmarked_ell():
    """C4 = 200 000 x 200 000 with 5-dim colour at ell_init_first_frame-like ell = 1.5, identity
    pose: the regime SCALE measures (dense scan, colour cut in the emission path, saturated rows)."""
    src, tgt, _ = synthetic_pair(250000, 200000, 200000, 20004, F=5)
    p = u.read_params_yaml(os.path.join(DATA, "cvo_intensity_params_img_gpu0.yaml"))
    want = assert_product_equals_reference(p, src, tgt, I3, Z3, 1.5, "C4 ell=1.5", 10000)
    print("C4 ell=1.5: nnz", want["nonzero_sum"], "max row", int(want["nonzeros"].max()))


# ------------------------------------------------------------------ the oracle against the device builds
def _oracle_vs_device(kind, device_arith):
    oracle.set_device_arith(device_arith)
    try:
        report = {}
        cases = []
        src, tgt = demo_clouds(color=True)
        p = demo_params(src, tgt, color=True)
        cases.append(("demo colour", p, src, tgt, I3, Z3, float(p.ell_init), int(p.nearest_neighbors_max)))
        src, tgt = demo_clouds(color=False)
        p = demo_params(src, tgt, color=False)
        cases.append(("demo geometric", p, src, tgt, I3, Z3, float(p.ell_init), int(p.nearest_neighbors_max)))
        src, tgt, _ = synthetic_pair(12500, 10000, 10000, 20002)
        cases.append(("C2", geometric_params(), src, tgt, rot_z(1.0), np.array([0.02, -0.01, 0.3], np.float32), 0.95, 256))
        src, tgt, _ = synthetic_pair(4000, 3000, 3300, 31, F=5, C=19, geotype=True)
        p = u.read_params_yaml(os.path.join(DATA, "cvo_semantic_params_img_gpu0.yaml"))
        p.is_using_geometric_type = 1
        p.c_ell = 1.0
        p.sp_thres = 0.001
        cases.append(("full kernel", p, src, tgt, rot_z(-2.0), np.array([0.05, 0.02, 0.5], np.float32), 1.2, 512))
        for what, p, src, tgt, R, T, ell, cap in cases:
            cs, ct = to_oracle_cloud(src), to_oracle_cloud(tgt)
            ym = oracle.transform(R, T, ct.xyz)
            want = ref.fill_A(p, cs, ct, ym, cap, ell, kind=kind)
            got = oracle.fill_A(p, cs, ct, ym, cap, ell)
            assert np.array_equal(got["nonzeros"], want["nonzeros"]), f"{what}: counts differ"
            assert np.array_equal(got["ind"], want["ind"]), f"{what}: indices differ"
            gi, wi = got["mat"].view(np.int32).astype(np.int64), want["mat"].view(np.int32).astype(np.int64)
            ulps = np.abs(gi - wi)
            assert ulps.max() <= 1, f"{what}: values differ by {ulps.max()} ulp"
            report[what] = (int((ulps > 0).sum()), int(want["nonzero_sum"]))
        return report
    finally:
        oracle.set_device_arith(True)


def test_oracle_device_arithmetic_equals_the_reference_flags_build():
    rep = _oracle_vs_device("cuda", True)
    print("oracle (device arithmetic) vs reference kernel, reference flags: values off by one ulp / entries:", rep)


def test_oracle_as_written_equals_the_nofma_build():
    rep = _oracle_vs_device("cuda_nofma", False)
    print("oracle (as written) vs reference kernel --fmad=false: values off by one ulp / entries:", rep)


def test_contraction_changes_results():
    """why the arithmetic mode matters: the two device builds of the same text do not agree."""
    src, tgt, _ = synthetic_pair(12500, 10000, 10000, 20002)
    p = geometric_params()
    cs, ct = to_oracle_cloud(src), to_oracle_cloud(tgt)
    a = ref.fill_A(p, cs, ct, ct.xyz, 256, 0.95, kind="cuda")
    b = ref.fill_A(p, cs, ct, ct.xyz, 256, 0.95, kind="cuda_nofma")
    differ = int((a["mat"].view(np.uint32) != b["mat"].view(np.uint32)).sum())
    print("fmad=true vs fmad=false builds of the reference kernel: entries that differ:", differ,
          "of", a["nonzero_sum"], "; counts equal:", bool(np.array_equal(a["nonzeros"], b["nonzeros"])))
    assert differ > 0


# ------------------------------------------------------------------ tier 2 on the device
@pytest.mark.parametrize("range_ell", [0, 1])
def test_device_flow_and_step_rows_against_oracle(range_ell):
    """K2 / K3+K4 of the reference (its Eigen primitives from oracle/ref_mini_eigen.h) compiled by
    nvcc with the reference's flags vs the oracle's rows: the device build contracts inside the
    3-term products (unknowable for real Eigen), so this is a tolerance check — 1e-5 of the row
    scale, an order below north_star's 1e-4 on the twist."""
    src, tgt, _ = synthetic_pair(3000, 2000, 2500, 123)
    p = geometric_params()
    p.is_using_range_ell = range_ell
    cs, ct = to_oracle_cloud(src), to_oracle_cloud(tgt)
    ym = oracle.transform(rot_z(0.7), [0.01, 0.03, 0.2], ct.xyz)
    ell, cap = 0.8, 40
    A = ref.fill_A(p, cs, ct, ym, cap, ell, kind="cuda")
    om_w, v_w = ref.flow_rows(p, cs.xyz, ym, A, kind="cuda")
    om_g, v_g = oracle.flow_rows(p, cs, ym, A)
    scale = max(np.abs(om_w).max(), np.abs(v_w).max())
    assert np.abs(om_g - om_w).max() <= 1e-5 * scale and np.abs(v_g - v_w).max() <= 1e-5 * scale
    tw = np.concatenate([om_w.sum(0), v_w.sum(0)]).astype(np.float32)
    tw /= np.linalg.norm(tw)
    want = ref.step_rows(tw[:3], tw[3:], ell, float(p.ell_init), range_ell, cs.xyz, ym, A, kind="cuda")
    got = oracle.step_rows(p, cs, ym, A, tw[:3], tw[3:], ell)
    for k in range(4):
        s = np.abs(want[:, k]).max()
        assert np.abs(got[:, k] - want[:, k]).max() <= 1e-4 * s, "BCDE"[k]
    # and the sums the controller consumes
    np.testing.assert_allclose(got.sum(0), want.sum(0), rtol=1e-5)
