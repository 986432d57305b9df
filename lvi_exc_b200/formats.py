"""File formats on either side of the hot path: what upstream producers hand to `lvi_init_orb_surfel` and what it
writes.  Host-side text I/O only.

  ORB/<bag>.txt   written by L/test/write_orb_slam_results.cpp:138-185, parsed by LIinitializer::LoadOrbResults (T:337-455):
                    FramePose <stamp_ns> tx ty tz qx qy qz qw
                    UV <keyframe_stamp_ns> (u v mappoint_id)*
                    MapPoint <id> x y z <ref_keyframe_stamp_ns>        (xyz in the reference keyframe's camera frame)
  LOAM/<bag>.txt  written by aloam/src/laserMapping.cpp:890-900, parsed by LIinitializer::ReadPoseGT (T:458-516):
                    <stamp_ns> tx ty tz qw qx qy qz
  <bag>.yaml      cv::FileStorage with Initial_T_*, T_lidar_imu_1st/2nd/3rd, T_cam_imu_2nd/3rd (T:564-707)
  result CSV      CalibParamManager::save_result (L/include/core/calibration.hpp:140-153)
(T = L/test/lvi_initialize_surfel_orb.cpp)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from .problem import quat_conj, quat_rot, quat_to_matrix


def _quat_from_matrix(R):
    t = np.trace(R)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2; q = [(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s]
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = math.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2; q = [0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s, (R[2, 1] - R[1, 2]) / s]
    elif R[1, 1] > R[2, 2]:
        s = math.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2; q = [(R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s, (R[0, 2] - R[2, 0]) / s]
    else:
        s = math.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2; q = [(R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s, (R[1, 0] - R[0, 1]) / s]
    q = np.array(q)
    return q / np.linalg.norm(q)


# ---- LOAM poses -----------------------------------------------------------------------------------------------------------
def write_loam_poses(path, stamps_s, poses):
    """poses [S,4,4]; line format of aloam/src/laserMapping.cpp:890-900 (note: qw first)"""
    with open(path, "w") as f:
        for t, T in zip(stamps_s, poses):
            q = _quat_from_matrix(T[:3, :3])
            f.write(f"{int(round(t * 1e9))} {T[0, 3]:.9f} {T[1, 3]:.9f} {T[2, 3]:.9f} {q[3]:.12f} {q[0]:.12f} {q[1]:.12f} {q[2]:.12f}\n")


def load_loam_poses(path):
    """ReadPoseGT (T:458-516): -> stamps [S] (s), poses [S,4,4], key-frame mask (first pose, then >= 5 deg or >= 0.1 m from the last key)"""
    stamps, poses, keys = [], [], []
    last_p, last_q = None, None
    for line in open(path):
        w = line.split()
        if len(w) != 8:
            break
        stamp = int(w[0])
        p = np.array([float(x) for x in w[1:4]])
        qw, qx, qy, qz = (float(x) for x in w[4:8])
        q = np.array([qx, qy, qz, qw])
        T = np.eye(4); T[:3, :3] = quat_to_matrix(q / np.linalg.norm(q)); T[:3, 3] = p
        stamps.append(stamp * 1e-9); poses.append(T)
        key = True
        if last_q is not None:
            d = abs(float(np.dot(last_q, q)))
            ang = 2.0 * math.acos(min(1.0, d))              # Eigen angularDistance
            if math.degrees(ang) < 5.0 and np.linalg.norm(last_p - p) < 0.1:
                key = False
        keys.append(key)
        if key:
            last_p, last_q = p, q
    return np.array(stamps), np.array(poses).reshape(-1, 4, 4), np.array(keys, dtype=bool)


# ---- ORB-SLAM2 results -----------------------------------------------------------------------------------------------------
@dataclass
class OrbResults:
    frame_stamps: np.ndarray = field(default_factory=lambda: np.zeros(0))          # [F] s
    frame_Tcw: np.ndarray = field(default_factory=lambda: np.zeros((0, 4, 4)))     # [F,4,4]
    view_t0: np.ndarray = field(default_factory=lambda: np.zeros(0))               # [V] s (views_db_, ascending stamp)
    obs_view: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    obs_landmark: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    obs_uv: np.ndarray = field(default_factory=lambda: np.zeros((0, 2)))
    lm_ids: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))      # original MapPoint ids
    lm_ref_obs: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))  # index into obs_*
    lm_rho: np.ndarray = field(default_factory=lambda: np.zeros(0))                # 1 / (z + 1e-15)  (T:425)


def write_orb_results(path, frame_stamps, frame_Tcw, view_stamps, view_obs, mappoints):
    """view_obs: list over views of [(u, v, mp_id)]; mappoints: list of (mp_id, x, y, z, ref_view_stamp_s)"""
    with open(path, "w") as f:
        for t, T in zip(frame_stamps, frame_Tcw):
            q = _quat_from_matrix(T[:3, :3])
            f.write(f"FramePose {int(round(t * 1e9))} {T[0, 3]:.9f} {T[1, 3]:.9f} {T[2, 3]:.9f} {q[0]:.12f} {q[1]:.12f} {q[2]:.12f} {q[3]:.12f}\n")
        for t, obs in zip(view_stamps, view_obs):
            f.write(f"UV {int(round(t * 1e9))} " + " ".join(f"{u:.6f} {v:.6f} {int(i)}" for u, v, i in obs) + "\n")
        for mp_id, x, y, z, tref in mappoints:
            f.write(f"MapPoint {int(mp_id)} {x:.9f} {y:.9f} {z:.9f} {int(round(tref * 1e9))}\n")


def load_orb_results(path, rows=720, cols=1280, border=10) -> OrbResults:
    """LoadOrbResults (T:337-455) including its filters: unknown reference keyframe, reference uv missing, 10 px border on the
    REFERENCE observation only; inverse depth = 1 / (z + 1e-15); observations attached in ascending view stamp order."""
    frames, uv_points, mps = [], {}, []
    for line in open(path):
        tok = line.split()
        if not tok:
            break
        if tok[0] == "FramePose":
            assert len(tok) == 9, "Wrong line!"
            t = np.array([float(x) for x in tok[2:5]])
            qx, qy, qz, qw = (float(x) for x in tok[5:9])
            T = np.eye(4); q = np.array([qx, qy, qz, qw]); T[:3, :3] = quat_to_matrix(q / np.linalg.norm(q)); T[:3, 3] = t
            frames.append((int(tok[1]) * 1e-9, T))
        elif tok[0] == "UV":
            assert (len(tok) - 2) % 3 == 0, "Wrong line!"
            fid = int(tok[1])
            uv_points[fid] = {int(tok[4 + 3 * i]): (float(tok[2 + 3 * i]), float(tok[3 + 3 * i])) for i in range((len(tok) - 2) // 3)}
        elif tok[0] == "MapPoint":
            assert len(tok) == 6, "Wrong line!"
            mps.append((int(tok[1]), float(tok[2]), float(tok[3]), float(tok[4]), int(tok[5])))
    view_ids = sorted(uv_points)                       # std::map<int64_t, View> iteration order
    vidx = {fid: k for k, fid in enumerate(view_ids)}
    out = OrbResults()
    out.frame_stamps = np.array([f[0] for f in frames]); out.frame_Tcw = np.array([f[1] for f in frames]).reshape(-1, 4, 4)
    out.view_t0 = np.array([fid * 1e-9 for fid in view_ids])
    ov, ol, ouv, ids, refs, rho = [], [], [], [], [], []
    seen = set()
    for lm_id, x, y, z, ref_id in mps:
        if ref_id not in uv_points or lm_id not in uv_points[ref_id] or lm_id in seen:
            continue
        u, v = uv_points[ref_id][lm_id]
        if u < border or v < border or u > cols - border or v > rows - border:
            continue
        seen.add(lm_id)
        l = len(ids)
        ids.append(lm_id); rho.append(1.0 / (z + 1e-15))
        refs.append(len(ov)); ov.append(vidx[ref_id]); ol.append(l); ouv.append((u, v))
        for fid in view_ids:
            if fid != ref_id and lm_id in uv_points[fid]:
                ov.append(vidx[fid]); ol.append(l); ouv.append(uv_points[fid][lm_id])
    out.obs_view, out.obs_landmark = np.array(ov, np.int32), np.array(ol, np.int32)
    out.obs_uv = np.array(ouv, dtype=np.float64).reshape(-1, 2)
    out.lm_ids, out.lm_ref_obs, out.lm_rho = np.array(ids, np.int64), np.array(refs, np.int32), np.array(rho)
    return out


# ---- outputs ---------------------------------------------------------------------------------------------------------------
def sensor_to_imu_matrix_inverse(q_StoI, p_SinI):
    """T_I2S as the reference stores it (T:571-575): R^T, R^T (-p)"""
    R = quat_to_matrix(q_StoI)
    T = np.eye(4); T[:3, :3] = R.T; T[:3, 3] = R.T @ (-np.asarray(p_SinI))
    return T


def write_result_yaml(path, matrices: dict, scalars: dict | None = None):
    """cv::FileStorage YAML 1.0 with 4x4 double matrices under the reference's keys (Initial_T_cam_imu, T_lidar_imu_1st, ...)"""
    with open(path, "w") as f:
        f.write("%YAML:1.0\n---\n")
        for k, M in matrices.items():
            M = np.asarray(M, dtype=np.float64)
            f.write(f"{k}: !!opencv-matrix\n   rows: {M.shape[0]}\n   cols: {M.shape[1]}\n   dt: d\n   data: [ " +
                    ", ".join(f"{x:.16e}" for x in M.ravel()) + " ]\n")
        for k, v in (scalars or {}).items():
            f.write(f"{k}: {v:.16e}\n")


def read_result_yaml(path) -> dict:
    out, lines = {}, open(path).read().splitlines()
    i = 0
    while i < len(lines):
        ln = lines[i]
        if ln.endswith("!!opencv-matrix"):
            key = ln.split(":")[0]
            r = int(lines[i + 1].split(":")[1]); c = int(lines[i + 2].split(":")[1])
            data = lines[i + 4].split("[")[1].split("]")[0]
            out[key] = np.array([float(x) for x in data.split(",")]).reshape(r, c)
            i += 5
        else:
            if ":" in ln and not ln.startswith("%") and ln != "---":
                k, v = ln.split(":", 1)
                out[k] = float(v)
            i += 1
    return out


def append_result_csv(path, info: str, calib, time_offset=0.0, gravity=(0.0, 0.0, 0.0)):
    """CalibParamManager::save_result (calibration.hpp:140-153): info, p_IinL, q_ItoL (x,y,z,w), time_offset, gravity, gyro bias, acce bias"""
    q_ItoL = quat_conj(calib.q_LtoI)
    p_IinL = quat_rot(q_ItoL, -np.asarray(calib.p_LinI))
    vals = [*p_IinL, *q_ItoL, time_offset, *gravity, *calib.gyr_bias, *calib.acc_bias]
    with open(path, "a") as f:
        f.write(info + "," + ",".join(f"{v:.10g}" for v in vals) + "\n")
