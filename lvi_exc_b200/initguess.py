"""Initial-guess stage of the calibrator: IMU pre-integration between sensor frames, rotation extrinsic by the
quaternion hand-eye equation, then gravity / velocities / translation extrinsic (and, for the up-to-scale camera track, the scale) by
linear alignment.  Host code, as in the reference — small dense least squares, nothing data-parallel:

  IntegrationBase                  L/include/vi_init/integration_base.h:17-323   (mid-point integration; delta_p / delta_q / delta_v / sum_dt)
  InitialEXRotation                L/src/vi_init/initial_ex_rotation.cpp:23-77   (CalibrationExRotationLiDAR)
  linear_alignment, refine_gravity L/src/vi_init/initial_aligment.cpp:48-66,130-171,175-423
  compute_integration_for_frames   L/test/lvi_initialize_surfel_orb.cpp:989-1042 (ComputeIntegrationForFrames)
  estimate_init_extrinsic          L/test/lvi_initialize_surfel_orb.cpp:1044-1148 (EstimateInitExtrinsicLI / CI)

  key-pose selection               T:457-509                                     (ReadPoseGT: 5 deg / 0.1 m gate on the LOAM poses)

Reference behaviour that is kept: every integrator's first IMU sample only seeds acc_0 / gyr_0 (push_back with dt < 0); the time of the
previous sample is carried ACROSS frames and the very first step uses 0.001 s; the alignment loops stop two frames before the end
(`i < size - 2`); the gyroscope-bias step is disabled (VisualIMUAlignment has it commented out), so the bias Jacobians and the covariance
of the pre-integration are never read and are not propagated here; gravity magnitude 9.7964; and (Q16) the velocity blocks of the
gravity systems are written at columns 0..5 for EVERY frame pair (`A.block<3, 6>(i * 3, 0)`, initial_aligment.cpp:167,301) instead of
columns 3i..3i+5, so only v_0 and v_1 are ever observed and `x.segment<3>(3 i)` is zero for i >= 2 (Eigen's LDLT answers the zero pivots
with 0).  That is kept on purpose: the velocity-only system has 3 (n - 2) rows, so with per-frame columns (3 n + 3 unknowns) it would be
under-determined, while the reference's 9 unknowns are determined — its gravity is the least-squares fit of THAT model, and everything
downstream (the calibration stages) starts from it.  Quaternions are (x, y, z, w)."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

G_NORM = 9.7964


def _qmul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz])


def _qrot(q, v):
    return quat_to_matrix(q) @ np.asarray(v, dtype=np.float64)


def quat_to_matrix(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def matrix_to_quat(R):
    """Eigen::Quaterniond(Matrix3d): w >= 0 branch first"""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[3] = (R[k, j] - R[j, k]) / s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
    return q


class IntegrationBase:
    """mid-point IMU pre-integration between two frames (integration_base.h:101-175,258-284), biases held at their linearisation point"""

    def __init__(self, ba=(0, 0, 0), bg=(0, 0, 0)):
        self.ba, self.bg = np.asarray(ba, dtype=np.float64), np.asarray(bg, dtype=np.float64)
        self.dt = -1.0
        self.sum_dt = 0.0
        self.delta_p, self.delta_v = np.zeros(3), np.zeros(3)
        self.delta_q = np.array([0, 0, 0, 1.0])
        self.acc_0 = self.gyr_0 = None

    def push_back(self, dt, acc, gyr):
        acc, gyr = np.asarray(acc, dtype=np.float64), np.asarray(gyr, dtype=np.float64)
        if self.dt < 0.0:            # the first sample of an integrator only seeds the mid-point rule (:103-112)
            self.dt = 1e-6
            self.acc_0, self.gyr_0 = acc, gyr
            return
        self.dt = dt
        un_acc_0 = _qrot(self.delta_q, self.acc_0 - self.ba)
        un_gyr = 0.5 * (self.gyr_0 + gyr) - self.bg
        q1 = _qmul(self.delta_q, np.array([un_gyr[0] * dt / 2, un_gyr[1] * dt / 2, un_gyr[2] * dt / 2, 1.0]))   # (:160-162) first-order, normalised below
        un_acc_1 = quat_to_matrix_unnormalised(q1) @ (acc - self.ba)   # Eigen rotates with the raw (not yet normalised) coefficients
        un_acc = 0.5 * (un_acc_0 + un_acc_1)
        self.delta_p = self.delta_p + self.delta_v * dt + 0.5 * un_acc * dt * dt
        self.delta_v = self.delta_v + un_acc * dt
        self.delta_q = q1 / np.linalg.norm(q1)                                                                   # propagate(): delta_q.normalize()
        self.sum_dt += dt
        self.acc_0, self.gyr_0 = acc, gyr


def quat_to_matrix_unnormalised(q):
    """Eigen's `q * v` on a non-unit quaternion: v + 2 w (u x v) + 2 u x (u x v) with the raw coefficients (what :163 evaluates before normalize())"""
    u, w = np.asarray(q[:3]), q[3]

    def rot(v):
        uv = np.cross(u, v) * 2.0
        return v + w * uv + np.cross(u, uv)
    return np.column_stack([rot(e) for e in np.eye(3)])


@dataclass
class IntegrationFrame:
    timestamp: float            # seconds
    T: np.ndarray               # 4x4 sensor pose in its odometry frame (the reference's `Tcw`: LOAM / ORB-SLAM pose of the sensor)
    integrator: IntegrationBase | None = None


def compute_integration_for_frames(frames: list, imu_t, gyro, accel) -> None:
    """ComputeIntegrationForFrames (T:989-1042): frame i gets the pre-integration of the IMU samples in [t_{i-1}, t_i)"""
    imu_t = np.asarray(imu_t)
    while frames and frames[0].timestamp < imu_t[0]:      # RemoveOverTimeFrames (T:965-986)
        frames.pop(0)
    while frames and frames[-1].timestamp > imu_t[-1]:
        frames.pop()
    k = 0
    last_imu_time = -1.0
    n = len(imu_t)
    for i in range(1, len(frames)):
        t_prev, t_cur = frames[i - 1].timestamp, frames[i].timestamp
        while k < n and imu_t[k] < t_prev:     # PopOldIMU
            k += 1
        integ = IntegrationBase()
        while k < n and imu_t[k] < t_cur:
            if last_imu_time < 0:
                integ.push_back(0.001, accel[k], gyro[k])
            else:
                integ.push_back(imu_t[k] - last_imu_time, accel[k], gyro[k])
            last_imu_time = imu_t[k]
            k += 1
        frames[i].integrator = integ


class InitialEXRotation:
    """rotation between a sensor and the IMU from pairs of relative rotations (initial_ex_rotation.cpp:4-77)"""

    def __init__(self):
        self.reset()

    def reset(self):
        self.frame_count = 0
        self.Rc, self.Rimu, self.Rc_g = [np.eye(3)], [np.eye(3)], [np.eye(3)]
        self.ric = np.eye(3)

    def calibrate(self, delta_R_sensor, delta_q_imu):
        """CalibrationExRotationLiDAR: returns (converged, ric) — converged when the second smallest singular value exceeds 0.25"""
        self.frame_count += 1
        Rimu = quat_to_matrix(delta_q_imu)
        self.Rc.append(np.asarray(delta_R_sensor, dtype=np.float64))
        self.Rimu.append(Rimu)
        self.Rc_g.append(self.ric.T @ Rimu @ self.ric)
        A = np.zeros((self.frame_count * 4, 4))
        for i in range(1, self.frame_count + 1):
            r1, r2 = matrix_to_quat(self.Rc[i]), matrix_to_quat(self.Rc_g[i])
            d = _qmul(r1, np.array([-r2[0], -r2[1], -r2[2], r2[3]]))
            ang = np.degrees(2.0 * np.arctan2(np.linalg.norm(d[:3]), abs(d[3])))
            huber = 5.0 / ang if ang > 5.0 else 1.0
            L, R = np.zeros((4, 4)), np.zeros((4, 4))
            q, w = r1[:3], r1[3]
            L[:3, :3] = w * np.eye(3) + _skew(q); L[:3, 3] = q; L[3, :3] = -q; L[3, 3] = w
            rij = matrix_to_quat(self.Rimu[i])
            q, w = rij[:3], rij[3]
            R[:3, :3] = w * np.eye(3) - _skew(q); R[:3, 3] = q; R[3, :3] = -q; R[3, 3] = w
            A[(i - 1) * 4:(i - 1) * 4 + 4] = huber * (L - R)
        _, s, Vt = np.linalg.svd(A, full_matrices=True)
        x = Vt[3]                                # singular vector of the smallest singular value, coefficients (x, y, z, w)
        est = quat_to_matrix(x / np.linalg.norm(x))
        self.ric = est.T
        ok = len(s) >= 3 and s[-2] > 0.25        # ric_cov = singularValues().tail<3>(); ric_cov(1) > 0.25
        return bool(ok), self.ric.copy()


def _skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


@dataclass
class AlignFrame:               # ImageFrame (initial_alignment.h:13-26): pose of the IMU-aligned sensor in the window's first frame
    R: np.ndarray
    T: np.ndarray
    pre: IntegrationBase


def _tangent_basis(g0):         # initial_aligment.cpp:48-61
    a = g0 / np.linalg.norm(g0)
    tmp = np.array([0, 0, 1.0])
    if np.array_equal(a, tmp):
        tmp = np.array([1.0, 0, 0])
    b = tmp - a * (a @ tmp)
    b /= np.linalg.norm(b)
    return np.column_stack([b, np.cross(a, b)])


def refine_gravity(frames: list, g):
    """RefineGravity (:130-171): four passes on the 2-dof tangent parameterisation with |g| = 9.7964; returns (g, x)"""
    g0 = g / np.linalg.norm(g) * G_NORM
    n = len(frames)
    x = None
    for _ in range(4):
        lxly = _tangent_basis(g0)
        A = np.zeros(((n - 1) * 3, n * 3 + 2))
        b = np.zeros((n - 1) * 3)
        for i in range(n - 2):
            fi, fj = frames[i], frames[i + 1]
            dt = fj.pre.sum_dt
            c = 0                    # Q16: `A.block<3, 6>(i * 3, 0)`
            A[3 * i:3 * i + 3, c:c + 3] += -np.eye(3)
            A[3 * i:3 * i + 3, c + 3:c + 6] += fi.R.T @ fj.R
            A[3 * i:3 * i + 3, 3 * n:3 * n + 2] += fi.R.T @ (dt * lxly)
            b[3 * i:3 * i + 3] += fj.pre.delta_v - fi.R.T @ (dt * g0)
        x = _ldlt_solve(A.T @ A * 1000.0, A.T @ b * 1000.0)
        g0 = g0 + lxly @ x[3 * n:3 * n + 2]
        g0 = g0 / np.linalg.norm(g0) * G_NORM
    return g0, x


def _ldlt_solve(A, b):
    """Eigen's ldlt().solve on a semi-definite normal matrix: unknowns that no row observes have an exactly zero pivot (and zero
    couplings), which LDLT answers with 0; the observed block is solved in the least-squares sense"""
    live = np.flatnonzero(np.diag(A) != 0.0)
    x = np.zeros(len(b))
    if len(live):
        x[live] = np.linalg.lstsq(A[np.ix_(live, live)], b[live], rcond=1e-14)[0]
    return x


def linear_alignment(frames: list, fix_scale: bool):
    """LinearAlignment (:175-423).  Returns (ok, g, T_ext, x): gravity in the window's first frame, the translation extrinsic, and
    x = velocities (fix_scale) or [T_ext, scale] (camera track)."""
    n = len(frames)
    A = np.zeros(((n - 1) * 3, n * 3 + 3))
    b = np.zeros((n - 1) * 3)
    for i in range(n - 2):
        fi, fj = frames[i], frames[i + 1]
        dt = fj.pre.sum_dt
        c = 0                        # Q16
        A[3 * i:3 * i + 3, c:c + 3] += -np.eye(3)
        A[3 * i:3 * i + 3, c + 3:c + 6] += fi.R.T @ fj.R
        A[3 * i:3 * i + 3, 3 * n:3 * n + 3] += fi.R.T * dt
        b[3 * i:3 * i + 3] += fj.pre.delta_v
    x = _ldlt_solve(A.T @ A * 1000.0, A.T @ b * 1000.0)
    g = x[3 * n:3 * n + 3]
    g, x = refine_gravity(frames, g)
    if not abs(np.linalg.norm(g) - G_NORM) < 0.5:
        return False, g, np.zeros(3), x
    m = 3 if fix_scale else 4
    A2 = np.zeros(((n - 1) * 3, m))
    b2 = np.zeros((n - 1) * 3)
    for i in range(n - 2):
        fi, fj = frames[i], frames[i + 1]
        dt = fj.pre.sum_dt
        A2[3 * i:3 * i + 3, :3] += np.eye(3) - fi.R.T @ fj.R
        rhs = fj.pre.delta_p + dt * x[3 * i:3 * i + 3] - (fi.R.T * (dt * dt / 2)) @ g
        if fix_scale:
            rhs = rhs - fi.R.T @ (fj.T - fi.T)
        else:
            A2[3 * i:3 * i + 3, 3] += fi.R.T @ (fj.T - fi.T)
        b2[3 * i:3 * i + 3] += rhs
    t = _ldlt_solve(A2.T @ A2, A2.T @ b2)
    if fix_scale:
        return True, g, t[:3], x
    return bool(t[3] > 0), g, t[:3], t


@dataclass
class InitGuess:
    ok_rotation: bool = False
    ok_translation: bool = False
    R_I_S: np.ndarray = field(default_factory=lambda: np.eye(3))      # rotation sensor -> IMU (R_I_L_ / R_I_C_)
    T_I_S: np.ndarray = field(default_factory=lambda: np.zeros(3))    # translation extrinsic (T_I_L_ / T_I_C_)
    gravity: np.ndarray = field(default_factory=lambda: np.zeros(4))  # g in the IMU frame of `gravity[3]` (g_I_t_*)
    scale: float = 1.0
    frames_used_rotation: int = 0

    @property
    def q_StoI(self):
        return matrix_to_quat(self.R_I_S)


def estimate_init_extrinsic(frames: list, fix_scale: bool, winsize: int = 10) -> InitGuess:
    """EstimateInitExtrinsicLI (fix_scale = True, T:1101-1148) / EstimateInitExtrinsicCI (False, T:1044-1099) on frames that already carry
    their pre-integrations: rotation from consecutive relative rotations until the estimator converges, then the first window of `winsize`
    frames whose linear alignment succeeds."""
    out = InitGuess()
    est = InitialEXRotation()
    for i in range(len(frames) - 1):
        if frames[i + 1].integrator is None:
            continue
        rel = frames[i].T[:3, :3].T @ frames[i + 1].T[:3, :3]
        ok, ric = est.calibrate(rel, frames[i + 1].integrator.delta_q)
        if ok:
            out.ok_rotation, out.R_I_S, out.frames_used_rotation = True, ric, i + 1
            break
    if not out.ok_rotation:
        return out
    for i in range(winsize, len(frames)):
        T_inv = np.linalg.inv(frames[i - winsize].T)
        win = []
        for k in range(i - winsize, i):
            Tk = T_inv @ frames[k].T
            win.append(AlignFrame(Tk[:3, :3] @ out.R_I_S.T, Tk[:3, 3].copy(), frames[k].integrator))
        if any(f.pre is None for f in win[1:]):
            continue
        if win[0].pre is None:
            win[0].pre = IntegrationBase()
        ok, g, T_ext, x = linear_alignment(win, fix_scale)
        if ok:
            out.ok_translation, out.T_I_S = True, T_ext
            out.gravity = np.array([*(out.R_I_S @ g), frames[i - winsize].timestamp])
            out.scale = 1.0 if fix_scale else float(x[-1])
            return out
    return out


def select_key_poses(stamps, poses) -> list:
    """ReadPoseGT's gate (T:496-509): a LOAM pose becomes an integration frame when it turned >= 5 deg or moved >= 0.1 m from the LAST READ
    pose -- `last_pos / last_ori` are refreshed only when a frame is kept (T:506-507)."""
    frames, last_q, last_p = [], None, None
    for t, T in zip(stamps, poses):
        q, p = matrix_to_quat(T[:3, :3]), T[:3, 3]
        if frames:
            d = _qmul(np.array([-last_q[0], -last_q[1], -last_q[2], last_q[3]]), q)
            ang = np.degrees(2.0 * np.arctan2(np.linalg.norm(d[:3]), abs(d[3])))
            if ang < 5.0 and np.linalg.norm(last_p - p) < 0.1:
                continue
        frames.append(IntegrationFrame(float(t), np.array(T, dtype=np.float64)))
        last_q, last_p = q, p.copy()
    return frames


def initial_extrinsics(lidar_stamps, lidar_poses, cam_stamps, cam_poses, imu_t, gyro, accel) -> dict:
    """The LiDAR-IMU and camera-IMU guesses of the reference's initialisation (ComputeIntegrationForFrames + EstimateInitExtrinsicLI / CI)
    in the form `pipeline.run_calibration` takes.  A failed estimate leaves its entry None."""
    out = dict(q_LtoI=None, p_LinI=None, q_CtoI=None, p_CinI=None, scale=1.0, gravity_lidar=None, gravity_cam=None)
    fl = select_key_poses(lidar_stamps, lidar_poses)
    compute_integration_for_frames(fl, imu_t, gyro, accel)
    gl = estimate_init_extrinsic(fl, True)
    if gl.ok_rotation:
        out["q_LtoI"] = gl.q_StoI
    if gl.ok_translation:
        out["p_LinI"], out["gravity_lidar"] = gl.T_I_S, gl.gravity
    if cam_stamps is not None and len(cam_stamps):
        fc = [IntegrationFrame(float(t), np.array(T, dtype=np.float64)) for t, T in zip(cam_stamps, cam_poses)]
        compute_integration_for_frames(fc, imu_t, gyro, accel)
        gc = estimate_init_extrinsic(fc, False)
        if gc.ok_rotation:
            out["q_CtoI"] = gc.q_StoI
        if gc.ok_translation:
            out["p_CinI"], out["gravity_cam"], out["scale"] = gc.T_I_S, gc.gravity, gc.scale
    out["lidar"], out["camera"] = gl, (gc if cam_stamps is not None and len(cam_stamps) else None)
    return out
