"""CPU tests of the host-side mirror of the reference interface (lvi_exc_b200/pipeline.py, problem.py, synth.py): stage policy of
TrajectoryManagerLVI, key-scan selection, observation filters, knot bookkeeping — and a whole LCIoptimize replay on the oracle."""
import math

import numpy as np
import pytest

from lvi_exc_b200 import pipeline, synth
from lvi_exc_b200.problem import ProblemData, num_knots_for, quat_from_axis_angle, quat_mul, quat_rot, quat_to_matrix
from tests.problems import _manager, _sequence, make_lvi_problem


def test_num_knots_matches_spline_extend_to():
    # dt 0.02, pad 0.2 -> 73 knots for 1 s, 3,023 for 60 s, 15,023 for 300 s (K/trajectories/spline_base.h:374-378)
    for dur, n in ((1.0, 73), (60.0, 3023), (300.0, 15023)):
        t0, k = num_knots_for(10.0, 10.0 + dur, 0.02, 0.2)
        assert t0 == pytest.approx(9.8) and k == n
        assert t0 + (k - 3) * 0.02 >= 10.0 + dur + 0.2 - 1e-9


def test_check_key_scan_thresholds():
    def pose(p, yaw_deg):
        T = np.eye(4); T[:3, :3] = quat_to_matrix(quat_from_axis_angle([0, 0, 1], math.radians(yaw_deg))); T[:3, 3] = p
        return T
    poses = np.array([pose([0, 0, 0], 0), pose([0.1, 0, 0], 1), pose([0.25, 0, 0], 1), pose([0.25, 0, 0], 5.5), pose([0.25, 0, 0], 11.0),
                      pose([0.3, 0.1, 0.05], 11.5)])
    # first scan always; 0.1 m / 1 deg no; 0.25 m yes; +4.5 deg from the last KEY (1 deg) no (< 5); 11 deg yes; small no
    assert pipeline.check_key_scan(poses).tolist() == [True, False, True, False, True, False]


def test_camera_observation_filters():
    seq = _sequence(2.0, 400)
    mgr = _manager(seq)
    rho = seq.lm_rho.copy()
    rho[::9] = -1.0     # non-positive inverse depth -> landmark skipped (trajectory_manager_lvi.cpp:519-524)
    co = pipeline.select_camera_observations(seq, mgr.min_time, mgr.max_time, rho)
    counts = np.bincount(seq.obs_landmark, minlength=len(rho))
    assert len(co["landmark"]) > 0
    assert (counts[co["landmark"]] > 5).all() and (rho[co["landmark"]] > 0).all()        # Q11: MORE than 5 observations
    assert (co["t0_ref"] <= co["t0_obs"]).all()                                           # reference = first observation
    kept = set(co["landmark"].tolist())
    assert all(counts[l] <= 5 or rho[l] <= 0 for l in range(len(rho)) if l not in kept and counts[l] > 0)


def test_stage_lock_policy():
    """lock flags per stage as TrajectoryManagerLVI sets them"""
    s0, s1 = make_lvi_problem("so3"), make_lvi_problem("surfel")
    s4, s5 = make_lvi_problem("lvi"), make_lvi_problem("lvi_locked")
    assert s0.r3_knots is None and s0.locks["lock_gyr_bias"] == 1 and "orient" in s0.tables and "accel" not in s0.tables
    assert (s1.locks["lock_lidar_q"], s1.locks["lock_lidar_p"], s1.locks["lock_cam_q"], s1.locks["lock_acc_bias"]) == (0, 0, 1, 0)
    assert (s4.locks["lock_r3"], s4.locks["lock_cam_q"], s4.locks["lock_cam_p"]) == (0, 0, 0)
    assert (s5.locks["lock_r3"], s5.locks["lock_so3"], s5.locks["lock_lidar_q"], s5.locks["lock_cam_q"]) == (1, 1, 1, 0)   # Q15
    assert np.all(s4.tables["cam"][5] == 1.0) and np.all(s4.tables["cam"][6] == 5.0)    # Q3: w_cam lands in the Huber slot
    assert np.all(s1.tables["surfel"][5] == 5.0) and np.all(s5.tables["camsurf"][5] == 30.0)
    t = s1.tables["gyro"][0]
    assert t.min() >= s1.min_time and t.max() < s1.max_time                            # IMU samples inside [MinTime, MaxTime)


def test_problem_desc_round_trip():
    pd = make_lvi_problem("lvi")
    d = pd.desc()
    assert d.n_knots == pd.n_knots and d.n_gyro == len(pd.tables["gyro"][0]) and d.n_cam == len(pd.tables["cam"][0])
    assert d.n_landmarks == len(pd.rho) and d.n_planes == len(pd.planes)
    assert np.ctypeslib.as_array(d.so3_knots, (pd.n_knots * 4,))[5] == pd.so3_knots.ravel()[5]
    saved = pd.clone_params()
    pd.r3_knots += 1.0; pd.lidar_p += 1.0
    pd.restore_params(saved)
    assert np.array_equal(pd.r3_knots, saved["r3_knots"]) and np.array_equal(pd.lidar_p, saved["lidar_p"])


def test_generator_is_deterministic_and_physically_consistent():
    cfg = synth.default_config(duration=0.5, n_landmarks=50)
    a, b = synth.make_sequence(cfg), synth.make_sequence(cfg)
    assert np.array_equal(a.scans_raw["timestamp"], b.scans_raw["timestamp"]) and np.array_equal(a.gyro, b.gyro)
    x = a.scans_raw["x"]
    assert np.array_equal(np.isnan(x), np.isnan(b.scans_raw["x"])) and np.array_equal(x[~np.isnan(x)], b.scans_raw["x"][~np.isnan(x)])
    # at rest convention (Q7): |accel| ~ 9.79 on average; timestamps increase along the azimuth sweep; 16 x 1800 organised scans
    assert abs(np.linalg.norm(a.accel, axis=1).mean() - 9.79) < 0.5
    assert a.scans_raw.shape[1:] == (16, 1800)
    assert np.all(np.diff(a.scans_raw["timestamp"][0, 0]) > 0)
    # no sample sits on a knot boundary of the 0.02 s grid anchored at map_time - 0.2 (the reference's segment lookup throws there)
    t0 = a.map_time - 0.2
    for t in (a.imu_t, a.view_t0, a.scans_raw["timestamp"][:, :, ::97].ravel()):
        frac = ((t - t0) / 0.02) % 1.0
        assert np.minimum(frac, 1 - frac).min() > 1e-6


def test_pipeline_replay_on_oracle_recovers_extrinsics():
    """LCIoptimize stage sequence (3 associations + S0..S5) on 6 s: the oracle itself converges towards the ground truth"""
    from tests.oracle_backend import OracleBackend
    seq = synth.make_sequence(synth.default_config(duration=6.0, n_landmarks=800))
    out = pipeline.run_calibration(seq, OracleBackend())
    names = [s["name"] for s in out["stages"]]
    assert names == ["S0_so3", "S1_surfel", "S2_refine", "S3_refine", "S4_lvi", "S5_lvi_surfel"]
    assert len(out["assoc_counts"]) == 3 and min(out["assoc_counts"]) > 1000
    e = pipeline.extrinsic_errors(out["calib"], seq.gt)
    init = pipeline.extrinsic_errors(pipeline.CalibParams(**pipeline.perturbed_initial_extrinsics(seq.gt)), seq.gt)
    assert e["rot_L"] < 0.1 * init["rot_L"] and e["rot_C"] < 0.1 * init["rot_C"]
    assert e["pos_L"] < 0.5 * init["pos_L"]
    assert out["stages"][-1]["final_cost"] < out["stages"][-2]["initial_cost"]
