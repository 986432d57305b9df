"""Host-side mirror of the reference's calibration orchestration for the hot path:

  * `TrajectoryManager`  = licalib::TrajectoryManagerLVI  (L/include/core/trajectory_manager_lvi.h:96-290,
                            L/src/core/trajectory_manager_lvi.cpp:43-62,138-351,464-606): builds each problem,
                            picks lock flags per stage, copies results back into the calibration parameters.
  * `run_calibration`    = LIinitializer::LCIoptimize stage sequence (L/test/lvi_initialize_surfel_orb.cpp:539-708,
                            1169-1300): S0, 3 data associations, S1-S5.

All heavy work goes through a `backend` object (the CUDA library in the product, the CPU oracle in tests); this
module only holds the policy.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from .problem import (CameraIntrinsics, ProblemData, num_knots_for, quat_angle, quat_conj, quat_from_axis_angle, quat_mul,
                      quat_rot, quat_to_matrix)


@dataclass
class CalibParams:
    """licalib::CalibParamManager (L/include/core/calibration.hpp:40-71) with the YAML weights of L/cfg/lvi.yaml"""
    q_LtoI: np.ndarray = field(default_factory=lambda: np.array([0, 0, 0, 1.0]))
    p_LinI: np.ndarray = field(default_factory=lambda: np.zeros(3))
    q_CtoI: np.ndarray = field(default_factory=lambda: np.array([0, 0, 0, 1.0]))
    p_CinI: np.ndarray = field(default_factory=lambda: np.zeros(3))
    gravity_rp: np.ndarray = field(default_factory=lambda: np.array([0.01, 0.01]))
    acc_bias: np.ndarray = field(default_factory=lambda: np.zeros(3))
    gyr_bias: np.ndarray = field(default_factory=lambda: np.zeros(3))
    w_gyr: float = 28.0
    w_acc: float = 18.0
    w_lidar: float = 10.0
    w_cam: float = 5.0            # passed as the Huber delta, not a weight (Q3)
    w_visual_surfel: float = 30.0


@dataclass
class PipelineConfig:
    ndt_resolution: float = 0.5
    associated_radius: float = 0.05
    knot_distance: float = 0.02
    time_offset_padding: float = 0.2
    plane_lambda_first: float = 0.6
    plane_lambda_refine: float = 0.7
    k_per_ring: int = 2
    time_downsample: int = 10
    iters_so3: int = 30
    iters_li: int = 30
    iters_lvi: int = 80
    n_refine: int = 2
    lock_traj_lidar_2nd: bool = False
    lock_traj_lidar_3rd: bool = True
    with_camera: bool = True
    initial_guess: str = "perturbed_gt"   # or "estimate": the reference's own initial-guess stage (initguess.py)


class TrajectoryManager:
    def __init__(self, cam: CameraIntrinsics, start_time: float, end_time: float, knot_distance: float, padding: float):
        self.cam = cam
        self.t0, self.n_knots = num_knots_for(start_time, end_time, knot_distance, padding)
        self.dt = knot_distance
        self.r3 = np.zeros((self.n_knots, 3))
        self.so3 = np.tile(np.array([0, 0, 0, 1.0]), (self.n_knots, 1))
        self.calib = CalibParams()
        self.imu_t = np.zeros(0)
        self.gyro = np.zeros((0, 3))
        self.accel = np.zeros((0, 3))
        self.rho = np.zeros(0)
        self.map_time = 0.0

    # feedIMUData (L/src/core/trajectory_manager_lvi.cpp:38-40)
    def feed_imu(self, t, gyro, accel):
        self.imu_t, self.gyro, self.accel = np.asarray(t), np.asarray(gyro), np.asarray(accel)

    @property
    def min_time(self):
        return self.t0

    @property
    def max_time(self):
        return self.t0 + (self.n_knots - 3) * self.dt

    def _base(self, so3_only=False, **locks) -> ProblemData:
        c = self.calib
        # every problem owns its parameter memory (a solve updates it in place; _copy_back takes the result)
        return ProblemData(self.t0, self.dt, self.n_knots, None if so3_only else self.r3.copy(), self.so3.copy(), lidar_q=c.q_LtoI, lidar_p=c.p_LinI,
                           cam_q=c.q_CtoI, cam_p=c.p_CinI, gravity=c.gravity_rp, acc_bias=c.acc_bias, gyr_bias=c.gyr_bias, cam=self.cam,
                           locks=locks)

    def _imu_mask(self):
        return (self.imu_t >= self.min_time) & (self.imu_t < self.max_time)   # :474-477

    def _add_imu(self, pd: ProblemData, accel=True):
        m = self._imu_mask()
        pd.set_gyro(self.imu_t[m], self.gyro[m], self.calib.w_gyr)           # addGyroscopeMeasurements :464-483
        if accel:
            pd.set_accel(self.imu_t[m], self.accel[m], self.calib.w_acc)     # addAccelerometerMeasurement :485-504

    def _copy_back(self, pd: ProblemData, lidar=False, cam=False):
        c = self.calib
        if pd.r3_knots is not None:
            self.r3 = pd.r3_knots
        self.so3 = pd.so3_knots
        if lidar:
            c.q_LtoI, c.p_LinI = pd.lidar_q.copy(), pd.lidar_p.copy()
        if cam:
            c.q_CtoI, c.p_CinI = pd.cam_q.copy(), pd.cam_p.copy()
        c.gravity_rp, c.acc_bias, c.gyr_bias = pd.gravity.copy(), pd.acc_bias.copy(), pd.gyr_bias.copy()

    # initialSO3TrajWithGyro (:43-62): SO3-only estimator, gyro + one orientation anchor at MinTime
    def problem_so3(self) -> ProblemData:
        pd = self._base(so3_only=True, lock_r3=1)
        self._add_imu(pd, accel=False)
        q0 = quat_from_axis_angle([0, 0, 1], 0.0001)
        pd.set_orientation([self.min_time], [q0], self.calib.w_gyr)
        return pd

    # trajInitFromSurfel (:311-351)
    def problem_surfel(self, planes_Pi, spoints, map_time) -> ProblemData:
        pd = self._base(lock_lidar_q=0, lock_lidar_p=0, lock_cam_q=1, lock_cam_p=1, lock_acc_bias=0, lock_gyr_bias=0)
        pd.planes = np.ascontiguousarray(planes_Pi, dtype=np.float64).reshape(-1, 3)
        self._add_imu(pd)
        self.map_time = map_time
        pd.set_surfel(spoints["timestamp"], map_time, spoints["point"], spoints["plane_id"], self.calib.w_lidar, 5.0)  # :561-582
        return pd

    # trajInitFromLVIdata (:138-257): camera (+ camera-surfel) on top of IMU + surfel
    def problem_lvi(self, planes_Pi, spoints, map_time, cam_obs: dict, rho, lm_plane: dict | None, lock_traj_and_lidar: bool) -> ProblemData:
        lk = int(lock_traj_and_lidar)
        pd = self._base(lock_r3=lk, lock_so3=lk, lock_lidar_q=lk, lock_lidar_p=lk, lock_cam_q=0, lock_cam_p=0, lock_acc_bias=0,
                        lock_gyr_bias=0)
        pd.planes = np.ascontiguousarray(planes_Pi, dtype=np.float64).reshape(-1, 3)
        pd.rho = np.ascontiguousarray(rho, dtype=np.float64).copy()
        self._add_imu(pd)
        self.map_time = map_time
        pd.set_surfel(spoints["timestamp"], map_time, spoints["point"], spoints["plane_id"], self.calib.w_lidar, 5.0)
        # addVisualObservation (:506-530): weight argument lands in the Huber slot (Q3) -> weight 1.0
        pd.set_camera(cam_obs["t0_ref"], cam_obs["t0_obs"], cam_obs["uv_ref"], cam_obs["uv_obs"], cam_obs["landmark"], 1.0, self.calib.w_cam)
        if lm_plane:  # addVisualLidarMeasurement (:584-606)
            lms = np.array(sorted(lm_plane.keys()), dtype=np.int32)
            pl = np.array([lm_plane[int(l)] for l in lms], dtype=np.int32)
            pd.set_camsurf(cam_obs["lm_ref_t0"][lms], map_time, cam_obs["lm_ref_uv"][lms], lms, pl, self.calib.w_visual_surfel, 5.0)
        return pd


def select_camera_observations(seq, min_time: float, max_time: float, rho: np.ndarray) -> dict:
    """addVisualObservation's filters (L/src/core/trajectory_manager_lvi.cpp:512-527): views inside the trajectory,
    landmarks with MORE THAN 5 observations (Q11) and positive inverse depth; ref = landmark->reference()."""
    nl = len(seq.lm_ref_obs)
    counts = np.bincount(seq.obs_landmark, minlength=nl)
    t_obs = seq.view_t0[seq.obs_view]
    keep = (t_obs >= min_time) & (t_obs < max_time) & (counts[seq.obs_landmark] > 5) & (rho[seq.obs_landmark] > 0)
    idx = np.nonzero(keep)[0]
    ref = seq.lm_ref_obs[seq.obs_landmark[idx]]
    lm_ref_t0 = np.zeros(nl)
    lm_ref_uv = np.zeros((nl, 2))
    has = seq.lm_ref_obs >= 0
    lm_ref_t0[has] = seq.view_t0[seq.obs_view[seq.lm_ref_obs[has]]]
    lm_ref_uv[has] = seq.obs_uv[seq.lm_ref_obs[has]]
    return dict(t0_ref=seq.view_t0[seq.obs_view[ref]], t0_obs=t_obs[idx], uv_ref=seq.obs_uv[ref], uv_obs=seq.obs_uv[idx],
                landmark=seq.obs_landmark[idx].astype(np.int32), lm_ref_t0=lm_ref_t0, lm_ref_uv=lm_ref_uv, counts=counts)


def perturbed_initial_extrinsics(gt: dict) -> dict:
    """Stand-in for the out-of-scope EstimateInitExtrinsic* stage: GT rotated by 3 deg about
    (1,1,1)/sqrt(3) and shifted by (0.05,-0.05,0.05) m."""
    dq = quat_from_axis_angle([1, 1, 1], math.radians(3.0))
    dp = np.array([0.05, -0.05, 0.05])
    return dict(q_LtoI=quat_mul(gt["q_LtoI"], dq), p_LinI=gt["p_LinI"] + dp, q_CtoI=quat_mul(gt["q_CtoI"], dq), p_CinI=gt["p_CinI"] + dp)


def estimated_initial_extrinsics(seq) -> dict:
    """EstimateInitExtrinsicLI / CI (T:1044-1148) on the sequence's LOAM key poses and its visual-odometry poses.  Where an estimate fails
    the reference's fall-backs apply: LCIoptimize refuses to run without both (T:540-543), CIoptimize uses identity / (-0.22, 0.02, 0.22)
    (T:520-525); here a failed entry raises."""
    from . import initguess
    cam_t, cam_T = seq.visual_odometry() if hasattr(seq, "visual_odometry") else (None, None)
    g = initguess.initial_extrinsics(seq.scan_times, seq.loam_poses, cam_t, cam_T, seq.imu_t, seq.gyro, seq.accel)
    for k in ("q_LtoI", "p_LinI") + (("q_CtoI", "p_CinI") if cam_t is not None else ()):
        if g[k] is None:
            raise RuntimeError(f"initial guess: {k} could not be estimated (The Camera and LiDAR should be inited first!)")
    if cam_t is None:
        g["q_CtoI"], g["p_CinI"] = np.array([0, 0, 0, 1.0]), np.array([-0.22, 0.02, 0.22])
    return g


def extrinsic_errors(calib: CalibParams, gt: dict) -> dict:
    return dict(rot_L=quat_angle(calib.q_LtoI, gt["q_LtoI"]), pos_L=float(np.linalg.norm(calib.p_LinI - gt["p_LinI"])),
                rot_C=quat_angle(calib.q_CtoI, gt["q_CtoI"]), pos_C=float(np.linalg.norm(calib.p_CinI - gt["p_CinI"])))


def check_key_scan(poses: np.ndarray) -> np.ndarray:
    """LiDAROdometry::checkKeyScan (L/src/core/lidar_odometry.cpp:107-128): first scan, or moved > 0.2 m, or any of
    yaw/pitch/roll changed by more than 5 (degrees; mathutils::R2ypr returns degrees)."""
    keys = np.zeros(len(poses), dtype=bool)
    pos_last, ypr_last = np.zeros(3), np.zeros(3)
    first = True
    for i, T in enumerate(poses):
        R = T[:3, :3]
        n, o, a = R[:, 0], R[:, 1], R[:, 2]
        y = math.atan2(n[1], n[0])
        p = math.atan2(-n[2], n[0] * math.cos(y) + n[1] * math.sin(y))
        r = math.atan2(a[0] * math.sin(y) - a[1] * math.cos(y), -o[0] * math.sin(y) + o[1] * math.cos(y))
        ypr = np.degrees([y, p, r])
        d = ypr - ypr_last
        d = np.abs((d + 180.0) % 360.0 - 180.0)
        if first or np.linalg.norm(T[:3, 3] - pos_last) > 0.2 or (d > 5.0).any():
            keys[i] = True
            pos_last, ypr_last, first = T[:3, 3].copy(), ypr, False
    return keys


def run_calibration(seq, backend, cfg: PipelineConfig | None = None, verbose: bool = False) -> dict:
    """LCIoptimize replay.  `backend` provides:
         solve(ProblemData, max_iterations) -> SolveSummary
         map_cloud(scans_in_map, key-scan mask or None) -> the cloud the map is built from (key scans concatenated)
         build_surfel_map(cloud, leaf, lambda) -> surfel-map handle with .planes_Pi [P,3]
         associate(map, scans_in_map [S,H,W,8] f32, scans_raw, radius, k, step) -> SurfelPoint array
         undistort(ProblemData-like trajectory, scans_raw, target_time, correct_position) -> [S,H,W,8] f32
         transform(scans_xyzi, poses) -> [S,H,W,8] f32        (pcl::transformPointCloud, float 4x4)
         associate_landmarks(map, p3d_L0 [L,3], radius) -> plane id per landmark (-1 none)
    """
    cfg = cfg or PipelineConfig()
    out = {"stages": []}
    mgr = TrajectoryManager(CameraIntrinsics(), seq.map_time, seq.end_time, cfg.knot_distance, cfg.time_offset_padding)
    init = perturbed_initial_extrinsics(seq.gt) if cfg.initial_guess != "estimate" else estimated_initial_extrinsics(seq)
    out["initial_guess"] = dict(kind=cfg.initial_guess, **{k: np.asarray(init[k]).tolist() for k in ("q_LtoI", "p_LinI", "q_CtoI", "p_CinI")})
    mgr.calib.q_LtoI, mgr.calib.p_LinI, mgr.calib.q_CtoI, mgr.calib.p_CinI = init["q_LtoI"], init["p_LinI"], init["q_CtoI"], init["p_CinI"]
    mgr.feed_imu(seq.imu_t, seq.gyro, seq.accel)
    map_time = seq.map_time

    def record(name, summary, pd):
        e = extrinsic_errors(mgr.calib, seq.gt)
        out["stages"].append(dict(name=name, iterations=summary.num_iterations, initial_cost=summary.initial_cost,
                                  final_cost=summary.final_cost, termination=summary.termination_type, errors=e,
                                  time_ms=summary.time_total_ms, n_res=summary.num_residuals))
        if verbose:
            print(f"[{name}] {summary.brief()} | rotL {e['rot_L']:.2e} posL {e['pos_L']:.2e} rotC {e['rot_C']:.2e} posC {e['pos_C']:.2e}")

    # S0
    pd = mgr.problem_so3()
    s = backend.solve(pd, cfg.iters_so3)
    mgr._copy_back(pd)
    record("S0_so3", s, pd)

    def traj_pd():
        return mgr._base()

    # A: first association = Mapping() with LOAM poses (T:1262-1300): rotation-only de-skew, key scans -> map
    scans_rot = backend.undistort(traj_pd(), seq.scans_raw, None, False)
    scans_in_map = backend.transform(scans_rot, seq.loam_poses)
    keys = check_key_scan(seq.loam_poses)
    smap = backend.build_surfel_map(backend.map_cloud(scans_in_map, keys), cfg.ndt_resolution, cfg.plane_lambda_first)
    spoints = backend.associate(smap, scans_in_map, seq.scans_raw, cfg.associated_radius, cfg.k_per_ring, cfg.time_downsample)
    out["assoc_counts"] = [len(spoints)]
    # S1
    pd = mgr.problem_surfel(smap.planes_Pi, spoints, map_time)
    s = backend.solve(pd, cfg.iters_li)
    mgr._copy_back(pd, lidar=True)
    record("S1_surfel", s, pd)
    # Refinement x n
    for r in range(cfg.n_refine):
        scans_in_map = backend.undistort(traj_pd(), seq.scans_raw, map_time, True)
        smap = backend.build_surfel_map(backend.map_cloud(scans_in_map), cfg.ndt_resolution, cfg.plane_lambda_refine)
        spoints = backend.associate(smap, scans_in_map, seq.scans_raw, cfg.associated_radius, cfg.k_per_ring, cfg.time_downsample)
        out["assoc_counts"].append(len(spoints))
        pd = mgr.problem_surfel(smap.planes_Pi, spoints, map_time)
        s = backend.solve(pd, cfg.iters_li)
        mgr._copy_back(pd, lidar=True)
        record(f"S{2 + r}_refine", s, pd)
    if cfg.with_camera and len(seq.obs_view):
        rho = seq.lm_rho.copy()
        cam_obs = select_camera_observations(seq, mgr.min_time, mgr.max_time, rho)
        pd = mgr.problem_lvi(smap.planes_Pi, spoints, map_time, cam_obs, rho, None, cfg.lock_traj_lidar_2nd)
        s = backend.solve(pd, cfg.iters_lvi)
        mgr._copy_back(pd, lidar=not cfg.lock_traj_lidar_2nd, cam=True)
        rho = pd.rho.copy()
        record("S4_lvi", s, pd)
        # associateVisualPointsWithPlanes (L/src/core/surfel_association.cpp:161-214)
        lm_plane = associate_landmarks(mgr, backend, smap, seq, cam_obs, rho, map_time, cfg.associated_radius)
        out["n_lm_plane"] = len(lm_plane)
        pd = mgr.problem_lvi(smap.planes_Pi, spoints, map_time, cam_obs, rho, lm_plane, cfg.lock_traj_lidar_3rd)
        s = backend.solve(pd, cfg.iters_lvi)
        mgr._copy_back(pd, lidar=not cfg.lock_traj_lidar_3rd, cam=True)
        record("S5_lvi_surfel", s, pd)
    out["calib"] = mgr.calib
    out["manager"] = mgr
    return out


def associate_landmarks(mgr: TrajectoryManager, backend, smap, seq, cam_obs, rho, map_time, radius) -> dict:
    """associateVisualPointsWithPlanes (L/src/core/surfel_association.cpp:161-214): landmark (ref uv, rho >= 0.05) -> 3-D in
    the L0 frame via the camera pose at its reference time and the L0 pose -> bbox + dist <= 2*radius; the last matching
    plane wins.  Pose evaluations and the plane test run on the backend (batched)."""
    c = mgr.calib
    cam = mgr.cam
    q_LtoC = quat_mul(quat_conj(c.q_CtoI), c.q_LtoI)
    t_LinC = quat_rot(quat_conj(c.q_CtoI), c.p_LinI - c.p_CinI)
    lms = np.nonzero((seq.lm_ref_obs >= 0))[0]
    t_ref = cam_obs["lm_ref_t0"][lms]
    keep = (rho[lms] >= 0.05) & (t_ref >= mgr.min_time) & (t_ref < mgr.max_time)
    lms, t_ref = lms[keep], t_ref[keep]
    if len(lms) == 0:
        return {}
    pd = mgr._base()
    pos, quat, valid = backend.traj_eval_many(pd, np.concatenate([[map_time], t_ref]))
    if not valid[0]:
        return {}

    def cam_pose(i):   # evaluateCameraPose: q_CtoG = q_ItoG * q_CtoI ; p_CinG = q_ItoG * p_CinI + p_IinG
        return quat_mul(quat[i], c.q_CtoI), quat_rot(quat[i], c.p_CinI) + pos[i]

    q_CtoG, p_CinG = cam_pose(0)
    q_L0_G = quat_mul(q_CtoG, q_LtoC)
    t_L0_G = quat_rot(q_CtoG, t_LinC) + p_CinG
    pts, ids = [], []
    for k, l in enumerate(lms):
        if not valid[k + 1]:
            continue
        uv = cam_obs["lm_ref_uv"][l]
        p_c = np.array([(uv[0] - cam.cx) / cam.fx, (uv[1] - cam.cy) / cam.fy, 1.0]) / rho[l]
        q, p = cam_pose(k + 1)
        p_g = quat_rot(q, p_c) + p
        pts.append(quat_rot(quat_conj(q_L0_G), p_g - t_L0_G))
        ids.append(int(l))
    if not pts:
        return {}
    plane = backend.associate_landmarks(smap, np.array(pts), radius)
    return {l: int(p) for l, p in zip(ids, plane) if p >= 0}
