"""Pose-graph edge update: BinaryStateGPU::update_inner_product
(IRLS_State_GPU.cu:43-79) on frames moved by their own poses (CvoFrameGPU.cu:44-62).

CPU part: the oracle's restatement against an independent numpy one.  GPU part: the C-ABI
(cvo_b200_frame_set / cvo_b200_edge_update, through the BinaryStateGPU mirror) against the
oracle — structure of the matrix bit-exact, values to 1e-6 relative."""
import os

import numpy as np
import pytest

import numpy_ref
import oracle
import unified_cvo_b200 as u
from helpers import DATA, geometric_params, synthetic_pair, to_oracle_cloud

f32 = np.float32


def pose_rt(rx_deg, ry_deg, rz_deg, t):
    """Row-major 3x4 [R t] in double, as CvoFrame::pose_vec holds it."""
    ax, ay, az = np.deg2rad([rx_deg, ry_deg, rz_deg])
    Rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
    Ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
    Rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
    P = np.zeros((3, 4))
    P[:, :3] = Rz @ Ry @ Rx
    P[:, 3] = t
    return P.reshape(12)


def numpy_pose_transform(pose12, xyz):
    """T * [x y z 1]^T with Eigen's 4-term unrolled redux (c0 + c1) + (c2 + c3), float."""
    P = np.asarray(pose12, f32).reshape(3, 4)
    x = np.asarray(xyz, f32)
    out = np.empty_like(x)
    for r in range(3):
        c0, c1, c2 = (P[r, 0] * x[:, 0]).astype(f32), (P[r, 1] * x[:, 1]).astype(f32), (P[r, 2] * x[:, 2]).astype(f32)
        c3 = f32(P[r, 3] * f32(1.0))
        out[:, r] = (c0 + c1).astype(f32) + (c2 + c3).astype(f32)
    return out


def _rows(sp):
    return [(sp["ind"][i, : sp["nonzeros"][i]], sp["mat"][i, : sp["nonzeros"][i]])
            for i in range(len(sp["nonzeros"]))]


# ------------------------------------------------------------------ CPU: oracle vs numpy
def test_oracle_pose_transform_is_bit_identical_to_numpy():
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((500, 3)) * [8, 2, 15] + [0, 0, 16]).astype(f32)
    P = pose_rt(1.0, -2.5, 0.7, [0.3, -0.1, 1.2]).astype(f32)
    assert np.array_equal(oracle.transform_pose_vec(P, x), numpy_pose_transform(P, x))
    I = np.eye(4)[:3].reshape(12).astype(f32)
    assert np.array_equal(oracle.transform_pose_vec(I, x), x)  # the identity pose moves nothing


@pytest.mark.usefixtures("as_written_arithmetic")
@pytest.mark.parametrize("cap,ell,colour", [(40, 0.9, False), (5, 1.5, False), (12, 1.2, True)])
def test_oracle_edge_update_matches_numpy_restatement(cap, ell, colour):
    if colour:
        f1, f2, _ = synthetic_pair(260, 150, 200, 5, F=5, C=4, geotype=True)
        p = u.read_params_yaml(os.path.join(DATA, "cvo_semantic_params_img_gpu0.yaml"))
        p.is_using_geometric_type = 1
        p.c_ell = 0.5
    else:
        f1, f2, _ = synthetic_pair(300, 200, 240, 11)
        p = geometric_params()
    P1 = pose_rt(0.5, 1.0, -0.4, [0.1, 0.0, -0.2]).astype(f32)
    P2 = pose_rt(0.2, 3.0, -0.1, [0.15, 0.02, 0.3]).astype(f32)
    total, sp = oracle.edge_update(p, to_oracle_cloud(f1), P1, to_oracle_cloud(f2), P2, ell, cap)
    x = numpy_pose_transform(P1, f1.positions_)
    y = numpy_pose_transform(P2, f2.positions_)
    rows = numpy_ref.fill_A(p, x, y, ell, cap, f1.features_, f2.features_, f1.labels_, f2.labels_,
                            f1.geometric_types_, f2.geometric_types_)
    assert total == sum(len(r[0]) for r in rows) and total > 50
    for (gi, gv), (ri, rv) in zip(_rows(sp), rows):
        assert np.array_equal(gi, ri)
        np.testing.assert_allclose(gv, rv, rtol=5e-7)
    assert sp["nonzeros"].max() <= cap


def test_oracle_edge_update_depends_on_both_poses_not_only_on_the_relative_one():
    """The length-scale of a row grows with the MOVED point's range (CvoGPU.cu:506-507), so the
    same relative pose applied at another place of the world gives another matrix — which is why
    the edge call takes two poses and not one relative transform."""
    f1, f2, _ = synthetic_pair(300, 200, 240, 11)
    p = geometric_params()
    I = np.eye(4)[:3].reshape(12)
    far = I.copy()
    far[[3, 7, 11]] = [300.0, 0.0, 0.0]
    a, _ = oracle.edge_update(p, to_oracle_cloud(f1), I, to_oracle_cloud(f2), I, 0.5, 64)
    b, _ = oracle.edge_update(p, to_oracle_cloud(f1), far, to_oracle_cloud(f2), far, 0.5, 64)
    assert b > a > 0


def _golden_edges():
    import json
    with open(os.path.join(os.path.dirname(__file__), "golden", "edge_updates.json")) as fh:
        return json.load(fh)


def test_oracle_edge_update_reproduces_the_committed_golden():
    from golden.make_golden import EDGE_CASES, run_edge_case
    for c, gold in zip(EDGE_CASES, _golden_edges()):
        got = run_edge_case(c)
        assert got["nnz"] == gold["nnz"] > 300 and got["max_row_nnz"] == gold["max_row_nnz"]
        assert got["row_ptr"] == gold["row_ptr"] and got["cols"] == gold["cols"]
        np.testing.assert_allclose(got["vals"], gold["vals"], rtol=1e-6)


# ------------------------------------------------------------------ GPU: C-ABI vs oracle
@pytest.mark.gpu
def test_edge_update_against_the_committed_golden():
    """tests/golden/edge_updates.json (make_golden.py, from the oracle): no oracle call here."""
    from golden.make_golden import EDGE_CASES, edge_inputs, edge_pose
    for c, gold in zip(EDGE_CASES, _golden_edges()):
        c1, c2, p = edge_inputs(c)
        g = u.CvoGPU(p)
        f1 = u.CvoFrameGPU(g, c1, edge_pose(c["pose1"]))
        f2 = u.CvoFrameGPU(g, c2, edge_pose(c["pose2"]))
        st = u.BinaryStateGPU(f1, f2, num_neighbor=c["cap"], init_ell=c["ell"])
        assert st.update_inner_product() == gold["nnz"]
        A = st.A_result_cpu_
        assert st.last_max_row_nnz == gold["max_row_nnz"]
        assert A.row_ptr.tolist() == gold["row_ptr"] and A.cols.tolist() == gold["cols"]
        np.testing.assert_allclose(A.vals, gold["vals"], rtol=1e-6)
        g.close()



def _check_edge(state, p, f1, f2, cap_expected=None):
    nnz = state.update_inner_product()
    if cap_expected is not None:
        assert state.num_neighbors_ == cap_expected
    total, sp = oracle.edge_update(p, to_oracle_cloud(f1.points), f1.pose_float(),
                                   to_oracle_cloud(f2.points), f2.pose_float(), state.ell_,
                                   state.num_neighbors_)
    row_ptr, cols, vals = oracle.sparse_to_csr(sp)
    A = state.A_result_cpu_
    assert nnz == total == len(A.vals)
    assert np.array_equal(A.row_ptr, row_ptr) and np.array_equal(A.cols, cols)
    np.testing.assert_allclose(A.vals, vals, rtol=1e-6)
    assert state.last_max_row_nnz == int(sp["nonzeros"].max())
    return nnz


@pytest.mark.gpu
def test_edge_update_matches_oracle_geometric_and_cap_schedule():
    """Three outer iterations of one edge with the frames' poses changing in between: every
    matrix equals the oracle's, and the cap follows min(init, 1.1 * last max row)."""
    c1, c2, _ = synthetic_pair(2500, 1500, 1800, 21)
    p = geometric_params()
    g = u.CvoGPU(p)
    f1 = u.CvoFrameGPU(g, c1, pose_rt(0.3, 0.5, -0.2, [0.05, 0.0, -0.1]))
    f2 = u.CvoFrameGPU(g, c2, pose_rt(0.1, 2.2, 0.0, [0.1, 0.02, 0.35]))
    st = u.BinaryStateGPU(f1, f2, num_neighbor=48, init_ell=0.8)
    n0 = _check_edge(st, p, f1, f2, cap_expected=48)
    assert n0 > 1000
    f2.pose_vec[:] = pose_rt(0.05, 2.05, 0.0, [0.06, 0.02, 0.45])  # the solver moved frame 2
    expect = min(48, int(st.last_max_row_nnz * 1.1))
    _check_edge(st, p, f1, f2, cap_expected=expect)
    st.update_ell()
    assert st.ell_ == pytest.approx(0.8 * p.multiframe_ell_decay_rate)
    _check_edge(st, p, f1, f2)
    assert st.iter_ == 3
    g.close()


@pytest.mark.gpu
def test_edge_update_colour_semantics_geotype_and_cap_above_nearest_neighbors_max():
    c1, c2, _ = synthetic_pair(1200, 700, 900, 5, F=5, C=20, geotype=True)
    p = u.read_params_yaml(os.path.join(DATA, "cvo_semantic_params_img_gpu0.yaml"))
    p.is_using_geometric_type = 1
    p.c_ell = 0.5
    p.nearest_neighbors_max = 16  # the edge's own cap may exceed the align() cap
    g = u.CvoGPU(p)
    f1 = u.CvoFrameGPU(g, c1, pose_rt(0.0, 0.4, 0.0, [0.0, 0.0, 0.05]))
    f2 = u.CvoFrameGPU(g, c2, pose_rt(0.0, 2.4, 0.0, [0.05, 0.02, 0.55]))
    st = u.BinaryStateGPU(f1, f2, num_neighbor=40, init_ell=1.2)
    assert _check_edge(st, p, f1, f2, cap_expected=40) > 200
    g.close()
    # geometry only: rows fill well beyond nearest_neighbors_max, up to the edge's own cap
    q = geometric_params()
    q.nearest_neighbors_max = 4
    g = u.CvoGPU(q)
    d1, d2, _ = synthetic_pair(1200, 700, 900, 6)
    f1 = u.CvoFrameGPU(g, d1, pose_rt(0.0, 0.4, 0.0, [0.0, 0.0, 0.05]))
    f2 = u.CvoFrameGPU(g, d2, pose_rt(0.0, 2.4, 0.0, [0.05, 0.02, 0.55]))
    st = u.BinaryStateGPU(f1, f2, num_neighbor=40, init_ell=1.2)
    _check_edge(st, q, f1, f2, cap_expected=40)
    assert 4 < st.last_max_row_nnz <= 40
    g.close()


@pytest.mark.gpu
def test_edge_loop_over_a_four_frame_graph_and_edge_cases():
    """update_edges over a ring + one loop-closure edge (IRLS.cpp:111-121): per-edge parity, frames
    shared between edges stay intact, zero cap / empty frame / unknown frame behave."""
    p = geometric_params()
    g = u.CvoGPU(p)
    clouds, frames = [], []
    for k in range(4):
        a, _, _ = synthetic_pair(1600, 900 + 64 * k, 900, 40)  # the same scene, ragged sizes
        clouds.append(a)
        frames.append(u.CvoFrameGPU(g, a, pose_rt(0.0, 0.3 * k, 0.0, [0.01 * k, 0.0, 0.02 * k])))
    edges = [(0, 1), (1, 2), (2, 3), (3, 0), (0, 2)]
    states = [u.BinaryStateGPU(frames[i], frames[j], num_neighbor=24, init_ell=0.5) for i, j in edges]
    total, per_edge = u.update_edges(states)
    assert total == sum(per_edge) and min(per_edge) > 100
    for st in states:  # matrices of earlier edges were copied out before the slots were reused
        t, sp = oracle.edge_update(p, to_oracle_cloud(st.frame1.points), st.frame1.pose_float(),
                                   to_oracle_cloud(st.frame2.points), st.frame2.pose_float(), 0.5, 24)
        row_ptr, cols, _ = oracle.sparse_to_csr(sp)
        assert np.array_equal(st.A_result_cpu_.row_ptr, row_ptr)
        assert np.array_equal(st.A_result_cpu_.cols, cols)
    # the reference's two-cloud calls still work on the same handle afterwards
    assert g.inner_product_gpu(clouds[0], clouds[1], np.eye(4, dtype=np.float32), 0.5) > 0
    # zero cap: fill_in_A_mat_gpu leaves every row empty (CvoGPU.cu:522)
    z = u.BinaryStateGPU(frames[0], frames[1], num_neighbor=0, init_ell=0.5)
    assert z.update_inner_product() == 0 and z.A_result_cpu_.row_ptr[-1] == 0
    # empty frame
    e = u.CvoFrameGPU(g, u.CvoPointCloud(np.zeros((0, 3), np.float32)))
    assert u.BinaryStateGPU(frames[0], e, 24, 0.5).update_inner_product() == 0
    assert u.BinaryStateGPU(e, frames[0], 24, 0.5).update_inner_product() == 0
    # a released frame is refused, loudly
    frames[3].release()
    with pytest.raises(u.CvoError):
        states[2].update_inner_product()
    g.close()


@pytest.mark.gpu
def test_batched_edge_update_equals_the_per_edge_calls():
    """cvo_b200_edge_update_batch (all edges of IRLS.cpp:111-121's loop in one call, one
    synchronisation): every edge's CSR equals the per-edge call's bit for bit, with ragged frames, an
    empty frame, a repeated edge, per-edge ell / cap, over three rounds."""
    p = geometric_params()
    g = u.CvoGPU(p)
    frames = []
    for k in range(4):
        a, _, _ = synthetic_pair(1600, 900 + 64 * k, 900, 40)
        frames.append(u.CvoFrameGPU(g, a, pose_rt(0.0, 0.3 * k, 0.0, [0.01 * k, 0.0, 0.02 * k])))
    frames.append(u.CvoFrameGPU(g, u.CvoPointCloud(np.zeros((0, 3), np.float32))))
    edges = [(0, 1, 0.5, 24), (1, 2, 0.5, 24), (2, 3, 0.8, 7), (3, 0, 0.5, 24), (0, 2, 0.3, 24), (0, 4, 0.5, 24),
             (4, 1, 0.5, 24), (0, 1, 0.5, 24), (1, 0, 0.5, 0)]

    def make():
        return [u.BinaryStateGPU(frames[i], frames[j], num_neighbor=c, init_ell=e) for i, j, e, c in edges]

    one, many = make(), make()
    for rep in range(3):  # the caps settle after the first pass; every data call answers from the size query's device-side copy
        total_1, per_1 = u.update_edges(one, batched=False)
        total_n, per_n = u.update_edges(many, batched=True)
        assert total_n == total_1 and per_n == per_1
        for a, b in zip(one, many):
            assert np.array_equal(a.A_result_cpu_.row_ptr, b.A_result_cpu_.row_ptr)
            assert np.array_equal(a.A_result_cpu_.cols, b.A_result_cpu_.cols)
            assert np.array_equal(a.A_result_cpu_.vals.view(np.uint32), b.A_result_cpu_.vals.view(np.uint32))
            assert a.last_max_row_nnz == b.last_max_row_nnz and a.num_neighbors_ == b.num_neighbors_
    assert per_1[5] == per_1[6] == per_1[8] == 0 and min(per_1[:5]) > 100
    # a pose changed between the calls: the cache must not answer
    frames[1].set_pose_vec(pose_rt(0.0, 0.31, 0.0, [0.02, 0.0, 0.02]))
    total_1b, per_1b = u.update_edges(one, batched=False)
    total_nb, per_nb = u.update_edges(many, batched=True)
    assert per_nb == per_1b and per_1b != per_1
    for a, b in zip(one, many):
        assert np.array_equal(a.A_result_cpu_.cols, b.A_result_cpu_.cols)
    g.close()


# ------------------------------------------------------------------ CPU: host logic of the mirror
class _FakeEdgeLib:
    """Stands in for libcvo_b200's two edge entry points (same two-call CSR protocol), answering
    from the oracle, so the HOST logic of CvoFrameGPU / BinaryStateGPU (ids, pose narrowing, cap
    schedule, CSR assembly) is exercised without a GPU.  Test-only."""

    def __init__(self, params):
        self.params, self.frames, self.calls = params, {}, []

    def cvo_b200_frame_set(self, h, fid, n, xyz, F, feat, C_, lab, geo):
        self.frames[fid] = n
        return 0

    def cvo_b200_frame_clear(self, h, fid):
        self.frames.pop(fid, None)
        return 0

    def bind(self, clouds):
        self.clouds = clouds  # frame id -> CvoPointCloud

    def cvo_b200_edge_update(self, h, f1, p1, f2, p2, ell, cap, nnz, mx, row_ptr, cols, vals):
        import ctypes as C
        P1 = np.ctypeslib.as_array(p1, shape=(12,)).copy()
        P2 = np.ctypeslib.as_array(p2, shape=(12,)).copy()
        self.calls.append((f1, f2, float(ell.value), int(cap), cols is not None))
        total, sp = oracle.edge_update(self.params, to_oracle_cloud(self.clouds[f1]), P1,
                                       to_oracle_cloud(self.clouds[f2]), P2, float(ell.value), int(cap))
        rp, cs, vs = oracle.sparse_to_csr(sp)
        C.cast(nnz, C.POINTER(C.c_int64))[0] = total
        C.cast(mx, C.POINTER(C.c_int32))[0] = int(sp["nonzeros"].max()) if len(sp["nonzeros"]) else 0
        np.ctypeslib.as_array(row_ptr, shape=(len(rp),))[:] = rp
        if cols is not None and total:
            np.ctypeslib.as_array(cols, shape=(total,))[:] = cs
            np.ctypeslib.as_array(vals, shape=(total,))[:] = vs
        return 0


    def cvo_b200_edge_update_batch(self, h, n, edges, nnz, mx, row_ptr, cols, vals):
        """same protocol as include/cvo_b200.h: concatenated row pointers (each edge's own, from 0),
        concatenated entries"""
        ro = eo = 0
        for k in range(n):
            e = edges[k]
            self.calls.append((e.frame1, e.frame2, float(e.ell), int(e.num_neighbors), cols is not None))
            total, sp = oracle.edge_update(self.params, to_oracle_cloud(self.clouds[e.frame1]),
                                           np.array(list(e.pose1), np.float32), to_oracle_cloud(self.clouds[e.frame2]),
                                           np.array(list(e.pose2), np.float32), float(e.ell), int(e.num_neighbors))
            rp, cs, vs = oracle.sparse_to_csr(sp)
            nnz[k], mx[k] = total, int(sp["nonzeros"].max()) if len(sp["nonzeros"]) else 0
            np.ctypeslib.as_array(row_ptr, shape=(ro + len(rp),))[ro:] = rp
            if cols and total:
                np.ctypeslib.as_array(cols, shape=(eo + total,))[eo:] = cs
                np.ctypeslib.as_array(vals, shape=(eo + total,))[eo:] = vs
            ro += len(rp)
            eo += total
        return 0


class _FakeGPU:
    def __init__(self, params):
        self.params, self._h, self._lib = params, 1, _FakeEdgeLib(params)

    def _check(self, rc):
        assert rc == 0

    def write_params(self):
        pass

    _fill_association = u.CvoGPU._fill_association


def test_mirror_host_logic_cap_schedule_pose_narrowing_and_two_call_protocol():
    c1, c2, _ = synthetic_pair(300, 200, 240, 11)
    p = geometric_params()
    p.multiframe_num_neighbors, p.multiframe_ell_init = 9, 1.5
    p.multiframe_ell_min, p.multiframe_ell_decay_rate = 1.0, 0.7
    g = _FakeGPU(p)
    f1 = u.CvoFrameGPU(g, c1, pose_rt(0.3, 0.5, -0.2, [0.05, 0.0, -0.1]))
    f2 = u.CvoFrameGPU(g, c2, np.vstack([pose_rt(0.1, 2.2, 0.0, [0.1, 0.02, 0.35]).reshape(3, 4), [0, 0, 0, 1]]))
    assert (f1.frame_id, f2.frame_id) == (0, 1) and g._lib.frames == {0: 200, 1: 240}
    assert f2.pose_vec.shape == (12,) and f1.pose_float().dtype == np.float32
    g._lib.bind({0: c1, 1: c2})
    st = u.BinaryStateGPU(f1, f2)  # CvoGPU.cu:1663-1666: cap and ell from the multiframe params
    assert st.num_neighbors_ == 9 and st.ell_ == pytest.approx(1.5)
    n0 = st.update_inner_product()
    total, sp = oracle.edge_update(p, to_oracle_cloud(c1), f1.pose_float(), to_oracle_cloud(c2),
                                   f2.pose_float(), 1.5, 9)
    assert n0 == total > 100 and st.last_max_row_nnz == sp["nonzeros"].max()
    assert np.array_equal(st.A_result_cpu_.cols, oracle.sparse_to_csr(sp)[1])
    assert st.A_result_cpu_.shape == (200, 240) and st.A_result_cpu_.to_scipy().nnz == total
    # size query first, then the entries: exactly two calls per update
    assert [c[4] for c in g._lib.calls] == [False, True]
    st.last_max_row_nnz = 5            # the next cap: min(9, int(5 * 1.1)) = 5  (IRLS_State_GPU.cu:45-47)
    st.update_inner_product()
    assert st.num_neighbors_ == 5 and g._lib.calls[-1][3] == 5
    st.last_max_row_nnz = 100
    st.update_inner_product()
    assert st.num_neighbors_ == 9      # never above the initial cap
    st.update_ell()                    # IRLS_State_GPU.cpp:54-57
    assert st.ell_ == pytest.approx(1.05)
    st.update_ell()
    assert st.ell_ == pytest.approx(0.735)
    st.update_ell()                    # 0.735 is not above ell_min = 1.0: no further decay
    assert st.ell_ == pytest.approx(0.735)
    assert u.update_edges([st])[0] == st.A_result_cpu_.row_ptr[-1]
    # the batched edge loop assembles the same per-edge matrices as the per-edge calls
    st2 = u.BinaryStateGPU(f2, f1, 7, 1.2)
    a, b = [u.BinaryStateGPU(f1, f2, 9, 1.5), st2], [u.BinaryStateGPU(f1, f2, 9, 1.5), u.BinaryStateGPU(f2, f1, 7, 1.2)]
    for _ in range(2):  # the second pass runs with the caps of the first (1.1 * fullest row)
        ta, pa = u.update_edges(a, batched=False)
        tb, pb = u.update_edges(b, batched=True)
        assert ta == tb > 0 and pa == pb
        for x, y in zip(a, b):
            assert x.num_neighbors_ == y.num_neighbors_ and x.last_max_row_nnz == y.last_max_row_nnz
            assert np.array_equal(x.A_result_cpu_.row_ptr, y.A_result_cpu_.row_ptr)
            assert np.array_equal(x.A_result_cpu_.cols, y.A_result_cpu_.cols)
            assert np.array_equal(x.A_result_cpu_.vals, y.A_result_cpu_.vals)
            assert x.A_result_cpu_.shape == y.A_result_cpu_.shape and x.iter_ == y.iter_
    f1.release()
    assert 0 not in g._lib.frames


def test_edge_parallel_loop_over_two_ranks_gathers_every_matrix_on_rank_0(tmp_path):
    """update_edges_sharded on 2 gloo ranks (host logic; the device is the oracle-backed stand-in):
    rank r refills edges r, r + 2, ..., rank 0 ends up with every edge's matrix, equal to the
    single-process loop."""
    import subprocess
    import sys
    import textwrap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    worker = tmp_path / "worker.py"
    worker.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
        import numpy as np
        import torch.distributed as dist
        import unified_cvo_b200 as u
        from helpers import geometric_params, synthetic_pair
        from test_multiframe import _FakeGPU, pose_rt
        rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        dist.init_process_group("gloo", rank=rank, world_size=world)
        p = geometric_params()
        clouds = [synthetic_pair(300, 150 + 16 * k, 150, 11)[0] for k in range(4)]
        poses = [pose_rt(0.0, 0.4 * k, 0.0, [0.02 * k, 0.0, 0.03 * k]) for k in range(4)]
        edges = [(0, 1), (1, 2), (2, 3), (3, 0), (0, 2)]

        def graph():
            g = _FakeGPU(p)
            frames = [u.CvoFrameGPU(g, c, P) for c, P in zip(clouds, poses)]
            g._lib.bind({{f.frame_id: c for f, c in zip(frames, clouds)}})
            return [u.BinaryStateGPU(frames[a], frames[b], 12, 1.2) for a, b in edges]

        sharded, single = graph(), graph()
        for _ in range(2):  # the second round runs with the caps of the first
            total, mine = u.update_edges_sharded(sharded, rank, world, dist)
            assert mine == list(range(rank, len(edges), world))
            u.update_edges(single)
        if rank == 0:
            for a, b in zip(sharded, single):
                assert np.array_equal(a.A_result_cpu_.row_ptr, b.A_result_cpu_.row_ptr)
                assert np.array_equal(a.A_result_cpu_.cols, b.A_result_cpu_.cols)
                assert np.array_equal(a.A_result_cpu_.vals, b.A_result_cpu_.vals)
                assert a.last_max_row_nnz == b.last_max_row_nnz > 0
            print("EDGE_SHARD_OK", sum(len(s.A_result_cpu_.vals) for s in sharded))
        dist.barrier()
    """))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29677", str(worker)],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-3000:])
    assert "EDGE_SHARD_OK" in out.stdout
