"""Benchmark workloads (BASELINE.json configs) assembled from the synthetic generator through a pipeline backend.

`lvi_stage_problem` builds the S4 problem of the reference's stage sequence (trajInitFromLVIdata, lock2 = false:
gyro + accel + surfel + camera residuals over the free Split trajectory, L/src/core/trajectory_manager_lvi.cpp:138-195)
at a state close to the optimum, WITHOUT running S0-S3 first: the control points are sampled from the analytic ground
truth and the LiDAR map / association come from one pass of the map path (undistort -> voxel covariance -> surfels ->
association) through the given backend.  One LM iteration costs the same wherever in state space it is evaluated, so
this is the workload `calibration iters/sec` is quoted on; the full stage sequence is `pipeline.run_calibration`.
"""
from __future__ import annotations

import time

import numpy as np

from . import pipeline, synth


def gt_control_points(seq, t0: float, dt: float, n_knots: int):
    """control points ~ ground-truth pose at each knot's support centre (a cubic B-spline tracks its control polygon to
    O(dt^2)); quaternion signs made continuous."""
    r3, so3 = np.zeros((n_knots, 3)), np.zeros((n_knots, 4))
    for i in range(n_knots):
        s = synth.gt_state(seq.cfg, seq.map_time, t0 + (i - 1) * dt)
        r3[i], so3[i] = s["p"], s["q"]
    for i in range(1, n_knots):
        if np.dot(so3[i], so3[i - 1]) < 0:
            so3[i] = -so3[i]
    return r3, so3


def make_manager(seq, cfg: pipeline.PipelineConfig | None = None, gt_traj: bool = True) -> pipeline.TrajectoryManager:
    cfg = cfg or pipeline.PipelineConfig()
    mgr = pipeline.TrajectoryManager(pipeline.CameraIntrinsics(), seq.map_time, seq.end_time, cfg.knot_distance, cfg.time_offset_padding)
    init = pipeline.perturbed_initial_extrinsics(seq.gt)
    mgr.calib.q_LtoI, mgr.calib.p_LinI, mgr.calib.q_CtoI, mgr.calib.p_CinI = init["q_LtoI"], init["p_LinI"], init["q_CtoI"], init["p_CinI"]
    mgr.feed_imu(seq.imu_t, seq.gyro, seq.accel)
    if gt_traj:
        mgr.r3, mgr.so3 = gt_control_points(seq, mgr.t0, mgr.dt, mgr.n_knots)
    return mgr


def lvi_stage_problem(seq, backend, cfg: pipeline.PipelineConfig | None = None, with_camera: bool = True):
    """-> (ProblemData of stage S4 (or S1 when with_camera is False), info dict with map-path timings and sizes)"""
    cfg = cfg or pipeline.PipelineConfig()
    mgr = make_manager(seq, cfg)
    # the map is assembled with the GT extrinsics so that the surfels are sharp; the solve starts from the perturbed ones
    mgr_map = make_manager(seq, cfg)
    mgr_map.calib.q_LtoI, mgr_map.calib.p_LinI = seq.gt["q_LtoI"], seq.gt["p_LinI"]
    info = {}
    t = time.perf_counter()
    scans_in_map = backend.undistort(mgr_map._base(), seq.scans_raw, seq.map_time, True)
    info["undistort_s"] = time.perf_counter() - t
    t = time.perf_counter()
    smap = backend.build_surfel_map(backend.map_cloud(scans_in_map), cfg.ndt_resolution, cfg.plane_lambda_refine)
    info["map_build_s"] = time.perf_counter() - t
    t = time.perf_counter()
    spoints = backend.associate(smap, scans_in_map, seq.scans_raw, cfg.associated_radius, cfg.k_per_ring, cfg.time_downsample)
    info["associate_s"] = time.perf_counter() - t
    info.update(n_points=int(np.prod(seq.scans_raw.shape)), n_planes=len(smap.planes_Pi), n_surfel_points=len(spoints))
    if with_camera:
        rho = seq.lm_rho.copy()
        cam_obs = pipeline.select_camera_observations(seq, mgr.min_time, mgr.max_time, rho)
        pd = mgr.problem_lvi(smap.planes_Pi, spoints, seq.map_time, cam_obs, rho, None, False)
        info["n_cam_obs"] = len(cam_obs["t0_obs"])
    else:
        pd = mgr.problem_surfel(smap.planes_Pi, spoints, seq.map_time)
    info["n_imu"] = len(pd.tables["gyro"][0])
    info["n_knots"] = mgr.n_knots
    return pd, info
