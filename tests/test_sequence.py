"""Trajectory plumbing: pose accumulation, constant-velocity initial guess,
first-frame parameter swap and the KITTI / TUM trajectory writers of the reference's sequence
drivers (main_cvo_gpu_align_raw_image.cpp:36-167, main_cvo_gpu_align_rgbd.cpp:38-141)."""
import os

import numpy as np
import pytest

import unified_cvo_b200 as u
from unified_cvo_b200 import sequence, synthetic
from helpers import DATA


def test_quaternion_matches_scipy_and_writers_follow_the_drivers_formats(tmp_path):
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(5)
    for _ in range(20):
        R = Rotation.from_rotvec(rng.normal(size=3) * rng.uniform(0.01, 3.1)).as_matrix()
        q = sequence.rotation_to_quaternion(R)
        ref = Rotation.from_matrix(R).as_quat()  # x y z w
        if np.dot(q, ref) < 0:
            ref = -ref
        np.testing.assert_allclose(q, ref, atol=1e-9)
    T = np.eye(4, dtype=np.float32)
    T[:3, 3] = [1.5, -0.25, 1e-7]
    assert sequence.kitti_line(np.eye(4)) == "1 0 0 0 0 1 0 0 0 0 1 0"  # raw_image.cpp:36
    assert sequence.kitti_line(T) == "1 0 0 1.5 0 1 0 -0.25 0 0 1 1e-07"  # ostream << float: %g
    assert sequence.tum_line("1305031102.175304", T) == "1305031102.175304 1.5 -0.25 1e-07 0 0 0 1"

    class FakeCvo:  # the plumbing alone: align returns a fixed motion
        def __init__(self):
            self.p = u.read_params_yaml(os.path.join(DATA, "cvo_intensity_params_img_gpu0.yaml"))
            self.writes, self.inits = [], []

        def get_params(self):
            return self.p

        def write_params(self):
            self.writes.append((self.p.ell_init, self.p.ell_decay_rate, self.p.ell_decay_start))

        def align(self, s, t, T_init):
            self.inits.append(np.array(T_init))
            info = u.AlignInfo()
            info.registration_seconds = 0.5
            return 0, synthetic.gt_transform().astype(np.float32), info

    f = FakeCvo()
    odo = sequence.FrameToFrameOdometry(f)
    assert f.writes[-1] == pytest.approx((1.5, 0.99, 600))          # first-frame swap
    odo.run([None, None, None, None])
    assert f.writes[-1] == pytest.approx((0.15, 0.95, 60)) and len(f.writes) == 2   # restored once
    G = synthetic.gt_transform()
    np.testing.assert_allclose(f.inits[0], np.eye(4), atol=0)          # first pair: identity guess
    np.testing.assert_allclose(f.inits[1], np.linalg.inv(G), atol=1e-6)  # then the inverted result
    np.testing.assert_allclose(odo.poses[3], G @ G @ G, atol=1e-5)
    assert odo.seconds == pytest.approx(1.5)
    odo.write_kitti(tmp_path / "05.txt")
    odo.write_tum(tmp_path / "tum.txt", ["a", "b", "c"])
    lines = open(tmp_path / "05.txt").read().splitlines()
    assert len(lines) == 4 and lines[0] == "1 0 0 0 0 1 0 0 0 0 1 0"
    np.testing.assert_allclose(np.array(lines[2].split(), float).reshape(3, 4), (G @ G)[:3], atol=1e-5)
    tum = open(tmp_path / "tum.txt").read().splitlines()
    assert len(tum) == 3 and tum[0].split()[0] == "a" and len(tum[0].split()) == 8
    assert sequence.kitti_translation_error(odo.poses, [np.linalg.matrix_power(G, k) for k in range(4)]) < 1e-5


def _moving_scene(n_frames, P, N, seed, F=5):
    """Frames of one static scene seen from a sensor moving with the constant motion T_gt."""
    base = synthetic.make_pair(P, N, N, seed, F=F)["source"]  # scene points + features
    G = synthetic.gt_transform()
    frames = []
    for k in range(n_frames):
        Tk = np.linalg.inv(np.linalg.matrix_power(G, k))  # scene -> frame k
        idx = np.sort(np.argsort(synthetic.uniform01(seed, 500 + k, N), kind="stable")[: int(0.85 * N)])
        xyz = base["xyz"][idx].astype(np.float64) @ Tk[:3, :3].T + Tk[:3, 3]
        xyz += 0.005 * np.stack([synthetic.normal(seed, 600 + 3 * k + a, len(idx)) for a in range(3)], axis=1)
        frames.append(u.CvoPointCloud(xyz.astype(np.float32), base["features"][idx], None, None))
    return frames, G


@pytest.mark.gpu
def test_synthetic_sequence_tracks_the_true_trajectory(tmp_path):
    frames, G = _moving_scene(4, 6000, 5000, 4242)
    p = u.read_params_yaml(os.path.join(DATA, "cvo_intensity_params_img_gpu0.yaml"))
    cvo = u.CvoGPU(p)
    odo = u.FrameToFrameOdometry(cvo)
    poses = odo.run(frames)
    assert len(poses) == 4 and all(i.ret == 0 for i in odo.infos)
    # the first pair starts from identity with the wide first-frame kernel, the others from the
    # constant-velocity guess with the regular one: far fewer iterations
    assert odo.infos[1].iterations < odo.infos[0].iterations
    gt = [np.linalg.matrix_power(G, k) for k in range(4)]
    assert np.abs(poses[3] - gt[3]).max() < 0.05
    assert sequence.kitti_translation_error(poses, gt) < 0.03
    q = cvo.get_params()
    assert (q.ell_init, q.ell_decay_start) == (pytest.approx(0.15), 60)
    odo.write_kitti(tmp_path / "seq.txt")
    assert len(open(tmp_path / "seq.txt").read().splitlines()) == 4
    cvo.close()


# ---- KITTI odometry evaluator against the reference's OWN published numbers -----------------
GOLD04 = os.path.join(os.path.dirname(__file__), "golden", "kitti04")
REF = "/root/reference"


def test_kitti_evaluator_reproduces_the_references_errors_and_stats_for_sequence_04():
    """Golden vectors from the reference (tests/golden/kitti04/README.md): its trajectory of KITTI
    04, the ground truth, and the per-segment errors + stats its devkit evaluator wrote
    (devkit/cpp/evaluate_odometry.cpp:83-143, :381-408)."""
    gt = sequence.load_kitti_poses(os.path.join(GOLD04, "gt.txt"))
    est = sequence.load_kitti_poses(os.path.join(GOLD04, "result.txt"))
    assert gt.shape == est.shape == (271, 4, 4)
    errs = sequence.kitti_sequence_errors(gt, est)
    gold = np.loadtxt(os.path.join(GOLD04, "errors.txt"))
    assert len(errs) == len(gold) == 43
    got = np.array(errs)
    assert np.array_equal(got[:, 0], gold[:, 0]) and np.array_equal(got[:, 3], gold[:, 3])
    np.testing.assert_allclose(got[:, 1], gold[:, 1], atol=1.5e-6)   # "%f": six decimals
    np.testing.assert_allclose(got[:, 2], gold[:, 2], atol=1.5e-6)
    np.testing.assert_allclose(got[:, 4], gold[:, 4], atol=1.5e-6)
    t_err, r_err = sequence.kitti_stats(errs)
    ref_t, ref_r = (float(x) for x in open(os.path.join(GOLD04, "stats.txt")).read().split())
    assert ref_t == pytest.approx(0.038597) and ref_r == pytest.approx(0.000401)
    assert abs(t_err - ref_t) < 1e-6 and abs(r_err - ref_r) < 1e-6
    # a perfect trajectory has no error; a short one has no 100 m segment
    assert sequence.kitti_stats(sequence.kitti_sequence_errors(gt, gt))[0] < 1e-6
    assert sequence.kitti_sequence_errors(gt[:20], est[:20]) == []


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "results")), reason="reference tree not present")
def test_kitti_evaluator_reproduces_every_published_stats_file_of_the_reference():
    """results/cvo_intensity_img_gpu0_oct25_best/{00..10}.txt against ground_truth/: the 11
    published (t_err, r_err) pairs of the reference's headline KITTI run, to the six decimals
    they are printed with.  (Runs only where /root/reference exists; never on the GPU box.)"""
    d = os.path.join(REF, "results", "cvo_intensity_img_gpu0_oct25_best")
    checked = 0
    for k in range(11):
        gt = sequence.load_kitti_poses(os.path.join(REF, "ground_truth", f"{k:02d}.txt"))
        est = sequence.load_kitti_poses(os.path.join(d, f"{k:02d}.txt"))
        ref_t, ref_r = (float(x) for x in open(os.path.join(d, "stats", f"{k:02d}_avg.txt")).read().split())
        t_err, r_err = sequence.kitti_stats(sequence.kitti_sequence_errors(gt, est))
        assert abs(t_err - ref_t) < 2e-6 and abs(r_err - ref_r) < 2e-6, (k, t_err, r_err, ref_t, ref_r)
        checked += 1
    assert checked == 11


def test_kitti_writer_round_trips_the_references_own_trajectory_file_verbatim():
    """The reference writes poses with operator<< at the default precision
    (main_cvo_gpu_align_raw_image.cpp:158-166): parsing its 04.txt and writing it back with
    kitti_line reproduces every line of the file character by character."""
    path = os.path.join(GOLD04, "result.txt")
    lines = open(path).read().splitlines()
    poses = sequence.load_kitti_poses(path)
    assert len(lines) == len(poses) == 271
    for line, T in zip(lines, poses):
        assert sequence.kitti_line(T) == line.strip()
