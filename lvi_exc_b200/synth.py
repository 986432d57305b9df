"""Binding to tools/synth/liblvi_synth.so — the deterministic synthetic VLP-16 + IMU + mono-camera generator
(SURVEY.md §8d).  Inputs come out in the reference's own layouts (PointXYZIT scans, LOAM poses, IMU samples,
ORB views/observations/landmarks)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from pathlib import Path

import numpy as np

from ._capi import RAW_POINT_DTYPE, ptr

_LIB = Path(__file__).resolve().parent.parent / "tools" / "synth" / "liblvi_synth.so"

SEED_RANGE, SEED_LOAM, SEED_IMU, SEED_CAM, SEED_LATTICE = 0xC0FFEE01, 0xC0FFEE02, 0xC0FFEE03, 0xC0FFEE04, 0xC0FFEE05


class SynthConfig(C.Structure):
    _fields_ = [("t_start", C.c_double), ("duration", C.c_double), ("rings", C.c_int32), ("az_steps", C.c_int32),
                ("scan_rate", C.c_double), ("imu_rate", C.c_double), ("cam_rate", C.c_double),
                ("keyframe_every", C.c_int32), ("n_landmarks", C.c_int32), ("max_track_views", C.c_int32),
                ("degenerate", C.c_int32),
                ("range_noise", C.c_double), ("loam_pos_noise", C.c_double), ("loam_rot_noise", C.c_double),
                ("gyro_noise", C.c_double), ("accel_noise", C.c_double), ("pixel_noise", C.c_double),
                ("rho_rel_noise", C.c_double), ("pad_time", C.c_double)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not _LIB.exists():
            raise RuntimeError(f"{_LIB} missing: run __graft_entry__.build()")
        L = C.CDLL(str(_LIB))
        L.synth_scan_time.restype = C.c_double
        L.synth_scan_time.argtypes = [C.POINTER(SynthConfig), C.c_int32]
        L.synth_num_scans.argtypes = [C.POINTER(SynthConfig)]
        L.synth_num_imu.argtypes = [C.POINTER(SynthConfig)]
        L.synth_num_views.argtypes = [C.POINTER(SynthConfig)]
        L.synth_scans.argtypes = [C.POINTER(SynthConfig), C.c_int32, C.c_int32, C.c_uint64, C.c_void_p]
        L.synth_loam_poses.argtypes = [C.POINTER(SynthConfig), C.c_int32, C.c_uint64, C.c_void_p]
        L.synth_imu.argtypes = [C.POINTER(SynthConfig), C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.synth_camera.argtypes = [C.POINTER(SynthConfig), C.c_uint64] + [C.c_void_p] * 6 + [C.c_int64]
        L.synth_camera.restype = C.c_int64
        L.synth_gt_state.argtypes = [C.POINTER(SynthConfig), C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.synth_gt_extrinsics.argtypes = [C.c_void_p] * 6
        L.synth_lattice_scans.argtypes = [C.POINTER(SynthConfig), C.c_int64, C.c_uint64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def default_config(**kw) -> SynthConfig:
    c = SynthConfig()
    lib().synth_default_config(C.byref(c))
    for k, v in kw.items():
        setattr(c, k, v)
    return c


@dataclass
class Sequence:
    cfg: SynthConfig
    scan_times: np.ndarray       # [S]
    scans_raw: np.ndarray        # [S, H, W] RAW_POINT_DTYPE
    loam_poses: np.ndarray       # [S, 4, 4] lidar pose in the first-scan lidar frame
    imu_t: np.ndarray
    gyro: np.ndarray
    accel: np.ndarray
    view_t0: np.ndarray          # [V]
    obs_view: np.ndarray         # [O]
    obs_landmark: np.ndarray     # [O]
    obs_uv: np.ndarray           # [O, 2]
    lm_ref_obs: np.ndarray       # [L] index into obs (or -1)
    lm_rho: np.ndarray           # [L]
    gt: dict

    @property
    def map_time(self) -> float:
        return float(self.scan_times[0])

    @property
    def end_time(self) -> float:
        return float(self.scan_times[-1] + 1.0 / self.cfg.scan_rate)

    def visual_odometry(self, scale: float = 2.5):
        """ORB-SLAM stand-in (`FramePose` of the reference's input): camera pose of every view in the first view's frame, translation
        UP TO SCALE (metric / `scale`), from the ground-truth kinematics -> (stamps [V], poses [V, 4, 4])"""
        if len(self.view_t0) == 0:
            return None, None
        def T_of(q, p):
            x, y, z, w = q
            T = np.eye(4)
            T[:3, :3] = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
            T[:3, 3] = p
            return T
        T_IC = T_of(self.gt["q_CtoI"], self.gt["p_CinI"])
        t0 = float(self.view_t0[0])
        Ts = []
        for t in self.view_t0:
            st = gt_state(self.cfg, t0, float(t))
            Ts.append(T_of(st["q"], st["p"]) @ T_IC)
        inv0 = np.linalg.inv(Ts[0])
        out = np.stack([inv0 @ T for T in Ts])
        out[:, :3, 3] /= scale
        return np.asarray(self.view_t0, dtype=np.float64), out


def gt_extrinsics() -> dict:
    q_l, p_l, q_c, p_c, bg, ba = (np.zeros(4), np.zeros(3), np.zeros(4), np.zeros(3), np.zeros(3), np.zeros(3))
    lib().synth_gt_extrinsics(*(a.ctypes.data for a in (q_l, p_l, q_c, p_c, bg, ba)))
    return dict(q_LtoI=q_l, p_LinI=p_l, q_CtoI=q_c, p_CinI=p_c, gyr_bias=bg, acc_bias=ba)


def gt_state(cfg: SynthConfig, t_ref: float, t: float):
    out = np.zeros(16)
    g = np.zeros(3)
    lib().synth_gt_state(C.byref(cfg), t_ref, t, out.ctypes.data, g.ctypes.data)
    return dict(p=out[0:3], q=out[3:7], v=out[7:10], a=out[10:13], w_body=out[13:16], gravity=g)


def make_scans(cfg: SynthConfig, first: int, n: int) -> np.ndarray:
    out = np.zeros((n, cfg.rings, cfg.az_steps), dtype=RAW_POINT_DTYPE)
    lib().synth_scans(C.byref(cfg), first, n, SEED_RANGE, out.ctypes.data)
    return out


def make_lattice_scans(cfg: SynthConfig, n_leaves: int, first: int, n: int, with_map: bool = False):
    """C3: scans over a map synthesised in voxel space (`n_leaves` planar leaves on a sparse lattice) -> raw scans
    [n, H, W] (and, optionally, the same points in the map frame [n, H, W, 8] float32)"""
    raw = np.zeros((n, cfg.rings, cfg.az_steps), dtype=RAW_POINT_DTYPE)
    m = np.zeros((n, cfg.rings, cfg.az_steps, 8), dtype=np.float32) if with_map else None
    lib().synth_lattice_scans(C.byref(cfg), n_leaves, SEED_LATTICE, first, n, raw.ctypes.data, m.ctypes.data if with_map else None)
    return raw, m


def make_imu(cfg: SynthConfig):
    n_imu = lib().synth_num_imu(C.byref(cfg))
    imu_t, gyro, accel = np.zeros(n_imu), np.zeros((n_imu, 3)), np.zeros((n_imu, 3))
    lib().synth_imu(C.byref(cfg), SEED_IMU, imu_t.ctypes.data, gyro.ctypes.data, accel.ctypes.data)
    return imu_t, gyro, accel


def scan_times(cfg: SynthConfig) -> np.ndarray:
    L = lib()
    return np.array([L.synth_scan_time(C.byref(cfg), i) for i in range(L.synth_num_scans(C.byref(cfg)))])


def make_sequence(cfg: SynthConfig, with_camera: bool = True, lattice_leaves: int = 0) -> Sequence:
    L = lib()
    S = L.synth_num_scans(C.byref(cfg))
    scan_times = np.array([L.synth_scan_time(C.byref(cfg), i) for i in range(S)])
    scans = make_lattice_scans(cfg, lattice_leaves, 0, S)[0] if lattice_leaves else make_scans(cfg, 0, S)
    poses = np.zeros((S, 4, 4))
    L.synth_loam_poses(C.byref(cfg), S, SEED_LOAM, poses.ctypes.data)
    n_imu = L.synth_num_imu(C.byref(cfg))
    imu_t, gyro, accel = np.zeros(n_imu), np.zeros((n_imu, 3)), np.zeros((n_imu, 3))
    L.synth_imu(C.byref(cfg), SEED_IMU, imu_t.ctypes.data, gyro.ctypes.data, accel.ctypes.data)
    nv = L.synth_num_views(C.byref(cfg))
    view_t0 = np.zeros(nv)
    nl = cfg.n_landmarks
    lm_ref, lm_rho = np.full(nl, -1, dtype=np.int32), np.zeros(nl)
    if with_camera and nl > 0:
        n_obs = L.synth_camera(C.byref(cfg), SEED_CAM, view_t0.ctypes.data, None, None, None, None, None, 0)
        ov, ol, ouv = np.zeros(n_obs, dtype=np.int32), np.zeros(n_obs, dtype=np.int32), np.zeros((n_obs, 2))
        L.synth_camera(C.byref(cfg), SEED_CAM, view_t0.ctypes.data, ov.ctypes.data, ol.ctypes.data, ouv.ctypes.data,
                       lm_ref.ctypes.data, lm_rho.ctypes.data, n_obs)
    else:
        ov, ol, ouv = np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32), np.zeros((0, 2))
    return Sequence(cfg, scan_times, scans, poses, imu_t, gyro, accel, view_t0, ov, ol, ouv, lm_ref, lm_rho, gt_extrinsics())


def transform_scans_float(scans_raw: np.ndarray, poses: np.ndarray) -> np.ndarray:
    """pcl::transformPointCloud with a float 4x4 (L/include/core/scan_undistortion.h:111): returns PointXYZI-shaped
    [S, H, W, 8] float32 (x,y,z,1,intensity,0,0,0); NaN points stay NaN."""
    S = scans_raw.shape[0]
    out = np.zeros(scans_raw.shape + (8,), dtype=np.float32)
    for s in range(S):
        T = poses[s].astype(np.float32)
        x, y, z = scans_raw[s]["x"], scans_raw[s]["y"], scans_raw[s]["z"]
        # Eigen: (T * [x,y,z,1]) evaluated in float, row by row
        for r in range(3):
            out[s, ..., r] = T[r, 0] * x + T[r, 1] * y + T[r, 2] * z + T[r, 3]
        out[s, ..., 3] = 1.0
        out[s, ..., 4] = scans_raw[s]["intensity"]
    return out
