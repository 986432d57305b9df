"""Small deterministic problems shared by the CPU and GPU parity tests (test infrastructure)."""
from __future__ import annotations

import ctypes as C
import functools

import numpy as np

from lvi_exc_b200 import pipeline, synth
from lvi_exc_b200._capi import ptr


@functools.lru_cache(maxsize=4)
def _sequence(duration: float, n_landmarks: int):
    cfg = synth.default_config(duration=duration, n_landmarks=n_landmarks)
    return synth.make_sequence(cfg)


@functools.lru_cache(maxsize=4)
def _oracle_association(duration: float, n_landmarks: int):
    """surfel planes + associated points of the first data association, computed by the oracle from GT-free inputs"""
    from tests.oracle_backend import OracleBackend
    seq = _sequence(duration, n_landmarks)
    orc = OracleBackend()
    mgr = _manager(seq)
    scans_map = orc.transform(orc.undistort(mgr._base(), seq.scans_raw, None, False), seq.loam_poses)
    smap = orc.build_surfel_map(scans_map.reshape(-1, 8), 0.5, 0.6)
    sp = orc.associate(smap, scans_map, seq.scans_raw, 0.05, 2, 10)
    return smap.planes_Pi.copy(), sp


def _manager(seq):
    mgr = pipeline.TrajectoryManager(pipeline.CameraIntrinsics(), seq.map_time, seq.end_time, 0.02, 0.2)
    init = pipeline.perturbed_initial_extrinsics(seq.gt)
    mgr.calib.q_LtoI, mgr.calib.p_LinI, mgr.calib.q_CtoI, mgr.calib.p_CinI = init["q_LtoI"], init["p_LinI"], init["q_CtoI"], init["p_CinI"]
    mgr.feed_imu(seq.imu_t, seq.gyro, seq.accel)
    return mgr


def gt_trajectory(seq, mgr, noise=0.0, seed=0):
    """control points sampled from the analytic ground truth (a B-spline's control points track the curve to O(dt^2)),
    optionally perturbed: a state close to, but not at, the optimum"""
    rng = np.random.default_rng(seed)
    n = mgr.n_knots
    r3, so3 = np.zeros((n, 3)), np.zeros((n, 4))
    for i in range(n):
        s = synth.gt_state(seq.cfg, seq.map_time, mgr.t0 + (i - 1) * mgr.dt)   # knot i peaks at t0 + (i-1) dt... (cubic B-spline support centre)
        r3[i], so3[i] = s["p"], s["q"]
    for i in range(1, n):   # keep the quaternion sign continuous
        if np.dot(so3[i], so3[i - 1]) < 0:
            so3[i] = -so3[i]
    if noise:
        r3 += noise * rng.standard_normal(r3.shape)
        so3 += 0.1 * noise * rng.standard_normal(so3.shape)
        so3 /= np.linalg.norm(so3, axis=1, keepdims=True)
    return r3, so3


DISTORTION = (-0.28, 0.07, 1.0e-3, -5.0e-4, 0.01)   # k1 k2 p1 p2 k3 of the "_dist" stages (a wide-angle lens; the model is active: |k1| > 1e-5)


def make_lvi_problem(stage: str, duration: float = 2.0, n_landmarks: int = 400):
    """stage: so3 (S0) | surfel (S1) | lvi (S4) | lvi_locked (S5, trajectory + LiDAR locked, camera-surfel residuals); the suffix `_dist`
    switches the pinhole camera's radial-tangential distortion on (SURVEY §8 f-4, K/sensors/pinhole_camera.h:131-240)"""
    seq = _sequence(duration, n_landmarks)
    mgr = _manager(seq)
    if stage.endswith("_dist"):
        stage = stage[:-5]
        mgr.cam.distortion = DISTORTION
    if stage == "so3":
        return mgr.problem_so3()
    planes, sp = _oracle_association(duration, n_landmarks)
    mgr.r3, mgr.so3 = gt_trajectory(seq, mgr, noise=2e-3, seed=1)
    if stage == "surfel":
        return mgr.problem_surfel(planes, sp, seq.map_time)
    rho = seq.lm_rho.copy()
    cam_obs = pipeline.select_camera_observations(seq, mgr.min_time, mgr.max_time, rho)
    if stage == "lvi":
        return mgr.problem_lvi(planes, sp, seq.map_time, cam_obs, rho, None, False)
    if stage == "lvi_locked":
        lms = np.nonzero((seq.lm_ref_obs >= 0) & (rho > 0.05))[0][:40]
        lm_plane = {int(l): int(i % len(planes)) for i, l in enumerate(lms)}
        return mgr.problem_lvi(planes, sp, seq.map_time, cam_obs, rho, lm_plane, True)
    raise ValueError(stage)


def map_tangent(backend, gp, op, pd) -> np.ndarray:
    """perm[library tangent position] = oracle tangent offset (-1 at the padding positions of the two-sided ordering), from the two
    libraries' own layout queries"""
    from lvi_exc_b200._capi import load
    lib = load()
    nt = gp.num_tangent
    lay = np.zeros(8, np.int32)
    assert lib.lvi_problem_layout(gp.h, ptr(lay)) == 0
    pad = np.arange(lay[6] - lay[7], lay[6])
    perm = np.full(nt, -1, dtype=np.int64)
    n = pd.n_knots
    for i in range(n):
        pos = lib.lvi_problem_tangent_offset_knot(gp.h, i)
        if pos < 0:
            continue
        o_r3 = op.offset_knot(i, False) if pd.r3_knots is not None else -1
        o_so3 = op.offset_knot(i, True)
        k = pos
        if o_r3 >= 0:
            perm[k:k + 3] = np.arange(o_r3, o_r3 + 3); k += 3
        if o_so3 >= 0:
            perm[k:k + 3] = np.arange(o_so3, o_so3 + 3)
    dims = [3, 3, 3, 3, 2, 3, 3]
    for which, dm in enumerate(dims):
        pos = lib.lvi_problem_tangent_offset_block(gp.h, which)
        off = op.offset_block(which)
        assert (pos < 0) == (off < 0), (which, pos, off)
        if pos >= 0:
            perm[pos:pos + dm] = np.arange(off, off + dm)
    for l in range(len(pd.rho)):
        pos = lib.lvi_problem_tangent_offset_block(gp.h, 7 + l)
        off = op.offset_block(7 + l)
        assert (pos < 0) == (off < 0)
        if pos >= 0:
            perm[pos] = off
    real = np.setdiff1d(np.arange(nt), pad)
    assert (perm[pad] < 0).all() and (perm[real] >= 0).all() and len(set(perm[real].tolist())) == len(real)
    return perm
