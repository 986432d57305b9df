"""Host-side problem record: what the Kontiki-shaped facade collects from AddMeasurement<M>() calls
(K/trajectory_estimator.h:71-74) before lowering to the C-ABI `lvi_problem_desc`.

`ProblemData` owns the numpy buffers (parameters are updated in place by a solve, like Ceres updates
DynamicParameterStore memory) and hands out a ctypes `ProblemDesc` that points into them.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from ._capi import ProblemDesc, ptr


def _f8(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _i4(a):
    return np.ascontiguousarray(a, dtype=np.int32)


@dataclass
class CameraIntrinsics:
    """PinholeCamera(rows, cols, readout, k1, k2, p1, p2, k3, fx, fy, cx, cy) — L/cfg/lvi.yaml:54-78 (distortion all zero there)"""
    fx: float = 530.175
    fy: float = 530.095
    cx: float = 635.12
    cy: float = 356.522
    readout: float = 0.0666
    rows: int = 720
    cols: int = 1280
    distortion: tuple = (0.0, 0.0, 0.0, 0.0, 0.0)   # k1 k2 p1 p2 k3 (K/sensors/pinhole_camera.h:72-80)


@dataclass
class ProblemData:
    t0: float
    dt: float
    n_knots: int
    r3_knots: np.ndarray | None          # [n,3] or None for the SO3-only estimator (S0)
    so3_knots: np.ndarray                # [n,4] x,y,z,w
    lidar_q: np.ndarray = field(default_factory=lambda: np.array([0, 0, 0, 1.0]))
    lidar_p: np.ndarray = field(default_factory=lambda: np.zeros(3))
    cam_q: np.ndarray = field(default_factory=lambda: np.array([0, 0, 0, 1.0]))
    cam_p: np.ndarray = field(default_factory=lambda: np.zeros(3))
    gravity: np.ndarray = field(default_factory=lambda: np.array([0.01, 0.01]))   # K/sensors/imu.h:127
    acc_bias: np.ndarray = field(default_factory=lambda: np.zeros(3))
    gyr_bias: np.ndarray = field(default_factory=lambda: np.zeros(3))
    cam: CameraIntrinsics = field(default_factory=CameraIntrinsics)
    rho: np.ndarray = field(default_factory=lambda: np.zeros(0))
    rho_locked: np.ndarray | None = None
    planes: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))
    locks: dict = field(default_factory=dict)   # lock_r3, lock_so3, lock_lidar_q, ... (default: sensors locked, K/sensors/sensors.h:95-97)
    tables: dict = field(default_factory=dict)  # residual tables, see set_* below
    lidar_toff: float = 0.0
    cam_toff: float = 0.0
    imu_toff: float = 0.0

    def __post_init__(self):
        self.so3_knots = _f8(self.so3_knots, (self.n_knots, 4))
        if self.r3_knots is not None:
            self.r3_knots = _f8(self.r3_knots, (self.n_knots, 3))
        for k in ("lidar_q", "lidar_p", "cam_q", "cam_p", "gravity", "acc_bias", "gyr_bias", "rho"):
            setattr(self, k, _f8(getattr(self, k)).copy())
        self.planes = _f8(self.planes).reshape(-1, 3)
        d = dict(lock_r3=0, lock_so3=0, lock_lidar_q=1, lock_lidar_p=1, lock_cam_q=1, lock_cam_p=1,
                 lock_acc_bias=1, lock_gyr_bias=1)   # K/sensors/constant_bias_imu.h:70-71
        d.update(self.locks)
        self.locks = d

    # ---- measurement tables (constructor argument order of the reference's measurement classes) -------------
    def set_gyro(self, t, w, weight):
        n = len(t)
        self.tables["gyro"] = (_f8(t), _f8(w, (n, 3)), _f8(np.broadcast_to(weight, (n,))))

    def set_accel(self, t, a, weight):
        n = len(t)
        self.tables["accel"] = (_f8(t), _f8(a, (n, 3)), _f8(np.broadcast_to(weight, (n,))))

    def set_surfel(self, t, t_map, point, plane_id, weight, huber=5.0):
        n = len(t)
        self.tables["surfel"] = (_f8(t), _f8(np.broadcast_to(t_map, (n,))), _f8(point, (n, 3)), _i4(plane_id),
                                 _f8(np.broadcast_to(weight, (n,))), _f8(np.broadcast_to(huber, (n,))))

    def set_camera(self, t0_ref, t0_obs, uv_ref, uv_obs, landmark, weight, huber):
        n = len(t0_ref)
        self.tables["cam"] = (_f8(t0_ref), _f8(t0_obs), _f8(uv_ref, (n, 2)), _f8(uv_obs, (n, 2)), _i4(landmark),
                              _f8(np.broadcast_to(weight, (n,))), _f8(np.broadcast_to(huber, (n,))))

    def set_camsurf(self, t, t_map, uv, landmark, plane_id, weight, huber=5.0):
        n = len(t)
        self.tables["camsurf"] = (_f8(t), _f8(np.broadcast_to(t_map, (n,))), _f8(uv, (n, 2)), _i4(landmark), _i4(plane_id),
                                  _f8(np.broadcast_to(weight, (n,))), _f8(np.broadcast_to(huber, (n,))))

    def set_orientation(self, t, q, weight):
        n = len(t)
        self.tables["orient"] = (_f8(t), _f8(q, (n, 4)), _f8(np.broadcast_to(weight, (n,))))

    @property
    def min_time(self) -> float:
        return self.t0

    @property
    def max_time(self) -> float:
        return self.t0 + (self.n_knots - 3) * self.dt   # K/trajectories/spline_base.h:53-56

    def desc(self) -> ProblemDesc:
        d = ProblemDesc()
        d.t0, d.dt, d.n_knots = self.t0, self.dt, self.n_knots
        d.r3_knots = ptr(self.r3_knots)
        d.so3_knots = ptr(self.so3_knots)
        for k in ("lidar_q", "lidar_p", "cam_q", "cam_p", "gravity", "acc_bias", "gyr_bias"):
            setattr(d, k, ptr(getattr(self, k)))
        d.lidar_toff, d.cam_toff, d.imu_toff = self.lidar_toff, self.cam_toff, self.imu_toff
        c = self.cam
        d.fx, d.fy, d.cx, d.cy, d.readout, d.cam_rows, d.cam_cols = c.fx, c.fy, c.cx, c.cy, c.readout, c.rows, c.cols
        for k in range(5):
            d.distortion[k] = float(c.distortion[k])
        d.n_landmarks = len(self.rho)
        d.n_planes = len(self.planes)
        d.rho = ptr(self.rho) if len(self.rho) else None
        if self.rho_locked is not None:
            self.rho_locked = np.ascontiguousarray(self.rho_locked, dtype=np.uint8)
            d.rho_locked = ptr(self.rho_locked)
        d.planes = ptr(self.planes) if len(self.planes) else None
        for k, v in self.locks.items():
            setattr(d, k, int(v))
        T = self.tables
        if "gyro" in T:
            d.n_gyro = len(T["gyro"][0]); d.gyro_t, d.gyro_w, d.gyro_weight = (ptr(a) for a in T["gyro"])
        if "accel" in T:
            d.n_accel = len(T["accel"][0]); d.accel_t, d.accel_a, d.accel_weight = (ptr(a) for a in T["accel"])
        if "surfel" in T:
            d.n_surfel = len(T["surfel"][0])
            (d.surfel_t, d.surfel_tmap, d.surfel_point, d.surfel_plane, d.surfel_weight, d.surfel_huber) = (ptr(a) for a in T["surfel"])
        if "cam" in T:
            d.n_cam = len(T["cam"][0])
            (d.cam_t0_ref, d.cam_t0_obs, d.cam_uv_ref, d.cam_uv_obs, d.cam_landmark, d.cam_weight, d.cam_huber) = (ptr(a) for a in T["cam"])
        if "camsurf" in T:
            d.n_camsurf = len(T["camsurf"][0])
            (d.cs_t, d.cs_tmap, d.cs_uv, d.cs_landmark, d.cs_plane, d.cs_weight, d.cs_huber) = (ptr(a) for a in T["camsurf"])
        if "orient" in T:
            d.n_orient = len(T["orient"][0]); d.orient_t, d.orient_q, d.orient_weight = (ptr(a) for a in T["orient"])
        return d

    def clone_params(self) -> dict:
        keys = ("r3_knots", "so3_knots", "lidar_q", "lidar_p", "cam_q", "cam_p", "gravity", "acc_bias", "gyr_bias", "rho")
        return {k: (None if getattr(self, k) is None else getattr(self, k).copy()) for k in keys}

    def restore_params(self, saved: dict) -> None:
        for k, v in saved.items():
            if v is not None:
                getattr(self, k)[...] = v


def num_knots_for(t_start: float, t_end: float, dt: float, padding: float) -> tuple[float, int]:
    """SplitTrajectory(dt, dt, t0, t0) + initialTrajTo(end) (L/include/core/trajectory_manager_lvi.h:120-124;
    ExtendTo: append knots while NumKnots < 4 or MaxTime < t, K/trajectories/spline_base.h:374-378)."""
    t0 = t_start - padding
    t_max = t_end + padding
    n = 4
    while t0 + (n - 3) * dt < t_max:
        n += 1
    return t0, n


def quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz])


def quat_conj(q):
    return np.array([-q[0], -q[1], -q[2], q[3]])


def quat_rot(q, v):
    u = np.asarray(q[:3]); w = q[3]
    uv = 2.0 * np.cross(u, v)
    return v + w * uv + np.cross(u, uv)


def quat_from_axis_angle(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    s = math.sin(angle / 2)
    return np.array([axis[0] * s, axis[1] * s, axis[2] * s, math.cos(angle / 2)])


def quat_angle(a, b) -> float:
    """|log(a^-1 b)| in radians"""
    d = quat_mul(quat_conj(a), b)
    return 2.0 * math.atan2(np.linalg.norm(d[:3]), abs(d[3]))


def quat_to_matrix(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
