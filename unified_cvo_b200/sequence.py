"""Frame-to-frame odometry plumbing around CvoGPU.align: what the reference's
sequence drivers do between two align() calls, for clouds that already exist in memory.

  main_cvo_gpu_align_raw_image.cpp:36-167 (KITTI)   main_cvo_gpu_align_rgbd.cpp:38-141 (TUM)
    * first-frame parameter swap: ell_init / ell_decay_rate / ell_decay_start <- *_first_frame for
      the first pair only, restored afterwards (:43-47, :150-156);
    * initial guess of pair k+1 = result of pair k (constant velocity), passed INVERTED as
      T_target_to_source (:96-97, :110);
    * accumulated pose accum <- accum * result (:128), logged per pair as a KITTI 3x4 row
      (:149-153) or a TUM "stamp tx ty tz qx qy qz qw" line (rgbd.cpp:129-133).
The perception front-end (stereo matching, FAST / DSO point selection) stays outside (§8)."""
from __future__ import annotations

from typing import Iterable, List, Optional

import numpy as np

from .cvo import CvoGPU, CvoPointCloud


def rotation_to_quaternion(R) -> np.ndarray:
    """(x, y, z, w) of a rotation matrix, Eigen::Quaternionf(Matrix3f) branch structure."""
    R = np.asarray(R, np.float64)
    t = R[0, 0] + R[1, 1] + R[2, 2]
    q = np.zeros(4)  # x y z w
    if t > 0.0:
        s = np.sqrt(t + 1.0)
        q[3] = 0.5 * s
        s = 0.5 / s
        q[0] = (R[2, 1] - R[1, 2]) * s
        q[1] = (R[0, 2] - R[2, 0]) * s
        q[2] = (R[1, 0] - R[0, 1]) * s
    else:
        i = 0
        if R[1, 1] > R[0, 0]:
            i = 1
        if R[2, 2] > R[i, i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
        q[i] = 0.5 * s
        s = 0.5 / s
        q[3] = (R[k, j] - R[j, k]) * s
        q[j] = (R[j, i] + R[i, j]) * s
        q[k] = (R[k, i] + R[i, k]) * s
    return q


def kitti_line(T) -> str:
    """One KITTI odometry pose: the first three rows of the 4x4, row-major, space separated."""
    T = np.asarray(T)
    return " ".join(f"{float(T[r, c]):g}" for r in range(3) for c in range(4))


def tum_line(stamp: str, T) -> str:
    T = np.asarray(T)
    q = rotation_to_quaternion(T[:3, :3])
    return (f"{stamp} {float(T[0, 3]):g} {float(T[1, 3]):g} {float(T[2, 3]):g} "
            f"{q[0]:g} {q[1]:g} {q[2]:g} {q[3]:g}")


class FrameToFrameOdometry:
    """The loop body of the sequence drivers.  `track(source, target)` registers one pair and
    returns the pair's transform; `poses` holds the accumulated 4x4 poses (identity first)."""

    def __init__(self, cvo: CvoGPU, first_frame_swap: bool = True):
        self.cvo = cvo
        p = cvo.get_params()
        self._regular = (p.ell_init, p.ell_decay_rate, p.ell_decay_start)
        self._first = first_frame_swap
        if first_frame_swap:  # raw_image.cpp:43-47
            p.ell_init = p.ell_init_first_frame
            p.ell_decay_rate = p.ell_decay_rate_first_frame
            p.ell_decay_start = p.ell_decay_start_first_frame
            cvo.write_params()
        self.init_guess = np.eye(4, dtype=np.float32)  # source frame -> target frame motion
        self.accum = np.eye(4, dtype=np.float32)
        self.poses: List[np.ndarray] = [self.accum.copy()]
        self.infos = []
        self.seconds = 0.0

    def track(self, source: CvoPointCloud, target: CvoPointCloud):
        init_guess_inv = np.linalg.inv(self.init_guess).astype(np.float32)  # :96-97
        ret, result, info = self.cvo.align(source, target, init_guess_inv)
        self.seconds += info.registration_seconds
        self.infos.append(info)
        self.init_guess = result.astype(np.float32)           # constant velocity, :127
        self.accum = (self.accum @ result).astype(np.float32)  # :128
        self.poses.append(self.accum.copy())
        if self._first:  # :150-156: back to the regular schedule after the first pair
            p = self.cvo.get_params()
            p.ell_init, p.ell_decay_rate, p.ell_decay_start = self._regular
            self.cvo.write_params()
            self._first = False
        return ret, result, info

    def run(self, clouds: Iterable[CvoPointCloud]):
        it = iter(clouds)
        source = next(it)
        for target in it:
            self.track(source, target)
            source = target  # :148
        return self.poses

    # ---- trajectory files
    def write_kitti(self, path: str):
        """results/<method>/NN.txt: one 3x4 row-major pose per frame, identity first (:36)."""
        with open(path, "w") as fh:
            for T in self.poses:
                fh.write(kitti_line(T) + "\n")

    def write_tum(self, path: str, stamps: Optional[Iterable[str]] = None):
        """'stamp tx ty tz qx qy qz qw' per registered frame (rgbd.cpp:129-133; no identity line)."""
        stamps = list(stamps) if stamps is not None else [str(i) for i in range(1, len(self.poses))]
        with open(path, "w") as fh:
            for s, T in zip(stamps, self.poses[1:]):
                fh.write(tum_line(s, T) + "\n")


def kitti_translation_error(poses_est, poses_gt) -> float:
    """Mean relative translation error of consecutive-frame motions (a one-segment-length
    simplification of devkit/cpp/evaluate_odometry.cpp:385-491, for synthetic sequences)."""
    errs = []
    for k in range(1, min(len(poses_est), len(poses_gt))):
        d_est = np.linalg.inv(poses_est[k - 1]) @ poses_est[k]
        d_gt = np.linalg.inv(poses_gt[k - 1]) @ poses_gt[k]
        e = np.linalg.inv(d_gt) @ d_est
        length = np.linalg.norm(d_gt[:3, 3])
        errs.append(np.linalg.norm(e[:3, 3]) / max(length, 1e-9))
    return float(np.mean(errs)) if errs else 0.0


# ---- KITTI odometry evaluation (devkit/cpp/evaluate_odometry.cpp) ------------------------------
KITTI_LENGTHS = (100, 200, 300, 400, 500, 600, 700, 800)  # evaluate_odometry.cpp:12


def load_kitti_poses(path: str) -> np.ndarray:
    """loadPoses (evaluate_odometry.cpp:25-43): one row-major 3x4 pose per line -> (n, 4, 4)."""
    rows = np.loadtxt(path, dtype=np.float64, ndmin=2)
    poses = np.tile(np.eye(4), (len(rows), 1, 1))
    poses[:, :3, :] = rows[:, :12].reshape(-1, 3, 4)
    return poses


def kitti_sequence_errors(poses_gt, poses_est, step_size: int = 10):
    """calcSequenceErrors (evaluate_odometry.cpp:83-127): for every 10th start frame and every
    segment length of 100..800 m (measured along the GROUND-TRUTH path), the rotation [rad/m] and
    translation [m/m] error of the estimated relative motion.  Distances, errors and the speed are
    float like the devkit's.  Returns rows (first_frame, r_err, t_err, len, speed)."""
    gt = np.asarray(poses_gt, np.float64)
    est = np.asarray(poses_est, np.float64)
    f32 = np.float32
    # trajectoryDistances (:45-57): float accumulation of float step lengths
    dist = np.zeros(len(gt), f32)
    for i in range(1, len(gt)):
        d = (gt[i - 1, :3, 3] - gt[i, :3, 3]).astype(f32)
        dist[i] = dist[i - 1] + f32(np.sqrt(f32(d[0] * d[0] + d[1] * d[1]) + f32(d[2] * d[2])))
    out = []
    for first in range(0, len(gt), step_size):
        for length in KITTI_LENGTHS:
            # lastFrameFromSegmentLength (:59-64)
            later = np.nonzero(dist[first:] > dist[first] + f32(length))[0]
            if len(later) == 0:
                continue
            last = first + int(later[0])
            d_gt = np.linalg.inv(gt[first]) @ gt[last]
            d_est = np.linalg.inv(est[first]) @ est[last]
            err = np.linalg.inv(d_est) @ d_gt
            # rotationError / translationError (:66-81)
            c = f32(0.5 * (f32(err[0, 0]) + f32(err[1, 1]) + f32(err[2, 2]) - 1.0))
            r_err = f32(np.arccos(max(min(c, f32(1.0)), f32(-1.0))))
            t = err[:3, 3].astype(f32)
            t_err = f32(np.sqrt(f32(t[0] * t[0] + t[1] * t[1]) + f32(t[2] * t[2])))
            num_frames = f32(last - first + 1)
            speed = f32(length / (0.1 * num_frames))
            out.append((first, float(r_err / f32(length)), float(t_err / f32(length)), float(length), float(speed)))
    return out


def kitti_stats(errors):
    """saveStats (evaluate_odometry.cpp:381-408): (mean t_err, mean r_err) as written to
    stats/NN_avg.txt — translation as a fraction (x100 = %), rotation in rad/m."""
    if not errors:
        return 0.0, 0.0
    t = np.float32(0)
    r = np.float32(0)
    for e in errors:  # float running sums like the devkit
        t = np.float32(t + np.float32(e[2]))
        r = np.float32(r + np.float32(e[1]))
    n = np.float32(len(errors))
    return float(t / n), float(r / n)
