"""Round trips of the text formats on either side of the hot path with the loaders' reference filters."""
import numpy as np

from lvi_exc_b200 import formats, pipeline, synth
from lvi_exc_b200.problem import quat_from_axis_angle, quat_to_matrix


def test_loam_pose_round_trip_and_key_frames(tmp_path):
    seq = synth.make_sequence(synth.default_config(duration=2.0, n_landmarks=0), with_camera=False)
    p = tmp_path / "loam.txt"
    formats.write_loam_poses(p, seq.scan_times, seq.loam_poses)
    stamps, poses, keys = formats.load_loam_poses(p)
    assert np.allclose(stamps, seq.scan_times, atol=1e-9) and np.allclose(poses, seq.loam_poses, atol=1e-8)
    assert keys[0] and 1 < keys.sum() <= len(keys)
    first = open(p).readline().split()
    assert len(first) == 8 and abs(float(first[4])) > 0.9          # qw right after the translation (laserMapping.cpp:890-900)


def test_orb_results_round_trip_and_filters(tmp_path):
    seq = synth.make_sequence(synth.default_config(duration=2.0, n_landmarks=200))
    nv = len(seq.view_t0)
    view_obs = [[] for _ in range(nv)]
    for v, l, uv in zip(seq.obs_view, seq.obs_landmark, seq.obs_uv):
        view_obs[v].append((uv[0], uv[1], 1000 + l))
    mappoints = []
    for l, (ro, rho) in enumerate(zip(seq.lm_ref_obs, seq.lm_rho)):
        if ro < 0:
            continue
        u, v = seq.obs_uv[ro]; z = 1.0 / rho
        mappoints.append((1000 + l, (u - 635.12) / 530.175 * z, (v - 356.522) / 530.095 * z, z, seq.view_t0[seq.obs_view[ro]]))
    mappoints.append((999999, 0.0, 0.0, 2.0, 123.0))                  # unknown reference keyframe -> skipped
    frames_T = np.tile(np.eye(4), (3, 1, 1)); frames_T[1, :3, :3] = quat_to_matrix(quat_from_axis_angle([0, 1, 0], 0.3)); frames_T[2, :3, 3] = [1, 2, 3]
    p = tmp_path / "orb.txt"
    formats.write_orb_results(p, [10.0, 10.05, 10.1], frames_T, seq.view_t0, view_obs, mappoints)
    res = formats.load_orb_results(p)
    assert np.allclose(res.frame_Tcw, frames_T, atol=1e-8) and np.allclose(res.view_t0, seq.view_t0, atol=1e-9)
    ok = np.array([ro >= 0 and 10 <= seq.obs_uv[ro][0] <= 1270 and 10 <= seq.obs_uv[ro][1] <= 710 for ro in seq.lm_ref_obs])
    assert len(res.lm_ids) == ok.sum() and set(res.lm_ids - 1000) == set(np.nonzero(ok)[0])     # border filter on the reference uv
    for k, l in enumerate(res.lm_ids - 1000):
        assert abs(res.lm_rho[k] - seq.lm_rho[l]) < 1e-6 * seq.lm_rho[l]
        assert np.allclose(res.obs_uv[res.lm_ref_obs[k]], seq.obs_uv[seq.lm_ref_obs[l]], atol=1e-5)
        assert (res.obs_landmark == k).sum() == (seq.obs_landmark == l).sum()
    # the loaded tracks drive the same observation filter as the generator's arrays
    assert res.obs_view.max() < nv and (np.diff(res.view_t0) > 0).all()


def test_result_yaml_and_csv(tmp_path):
    gt = synth.gt_extrinsics()
    calib = pipeline.CalibParams(q_LtoI=gt["q_LtoI"], p_LinI=gt["p_LinI"], q_CtoI=gt["q_CtoI"], p_CinI=gt["p_CinI"])
    mats = {"Initial_T_cam_imu": formats.sensor_to_imu_matrix_inverse(calib.q_CtoI, calib.p_CinI),
            "T_lidar_imu_1st": formats.sensor_to_imu_matrix_inverse(calib.q_LtoI, calib.p_LinI)}
    p = tmp_path / "bag.yaml"
    formats.write_result_yaml(p, mats, {"Error_gyro": 0.5})
    assert open(p).readline().strip() == "%YAML:1.0"
    back = formats.read_result_yaml(p)
    for k, M in mats.items():
        assert np.array_equal(back[k], M)
        assert np.allclose(M[:3, :3] @ M[:3, :3].T, np.eye(3), atol=1e-12)
    assert back["Error_gyro"] == 0.5
    T = mats["T_lidar_imu_1st"]                     # T_I2L: maps IMU-frame points into the LiDAR frame
    assert np.allclose(T[:3, :3] @ calib.p_LinI + T[:3, 3], 0.0, atol=1e-12)
    c = tmp_path / "res.csv"
    formats.append_result_csv(c, "stage1", calib); formats.append_result_csv(c, "stage2", calib)
    rows = open(c).read().strip().splitlines()
    assert len(rows) == 2 and rows[0].startswith("stage1,") and len(rows[0].split(",")) == 1 + 3 + 4 + 1 + 3 + 3 + 3
