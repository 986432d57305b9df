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


def test_cpp_loaders_read_what_python_writes(tmp_path):
    """include/lvi_exc_b200/compat/io/lvi_files.h (ReadPoseGT, LoadOrbResults, save_result with the reference's signatures and filters) against
    lvi_exc_b200/formats.py on the same files: key-pose selection, poses, landmark set, inverse depths, observation counts, reference
    observations; and the CSV line the C++ side appends has the reference's 18 fields"""
    import json
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    compat, lib = root / "include" / "lvi_exc_b200" / "compat", root / "lvi_exc_b200" / "lib"
    exe = root / "build" / "formats_check"
    exe.parent.mkdir(exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", str(compat), "-I", str(root / "include"), str(root / "tests" / "cpp" / "formats_check.cpp"),
                    "-o", str(exe), f"-L{lib}", "-llvi_exc_b200", f"-Wl,-rpath,{lib}"], check=True)
    seq = synth.make_sequence(synth.default_config(duration=3.0, n_landmarks=200))
    loam, orb, csv = tmp_path / "loam.txt", tmp_path / "orb.txt", tmp_path / "res.csv"
    formats.write_loam_poses(loam, seq.scan_times, seq.loam_poses)
    nv = len(seq.view_t0)
    view_obs = [[] for _ in range(nv)]
    for v, l, uv in zip(seq.obs_view, seq.obs_landmark, seq.obs_uv):
        view_obs[v].append((uv[0], uv[1], 1000 + l))
    mappoints = []
    for l, (ro, rho) in enumerate(zip(seq.lm_ref_obs, seq.lm_rho)):
        if ro >= 0:
            z = 1.0 / rho
            mappoints.append((1000 + l, 0.1 * z, -0.2 * z, z, seq.view_t0[seq.obs_view[ro]]))
    mappoints.append((999999, 0.0, 0.0, 2.0, 123.0))
    cam_t, cam_T = seq.visual_odometry()
    formats.write_orb_results(orb, cam_t[:5], cam_T[:5], seq.view_t0, view_obs, mappoints)
    r = subprocess.run([str(exe), str(loam), str(orb), str(csv)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    c = json.loads(r.stdout)
    stamps, poses, keys = formats.load_loam_poses(loam)
    assert c["n_loam"] == len(stamps)
    assert c["key_stamps"] == [int(round(t * 1e9)) for t in stamps[keys]]
    assert np.allclose(np.array(c["last_pose"]).reshape(3, 4), poses[-1][:3], atol=1e-12)
    res = formats.load_orb_results(orb)
    assert c["n_frames_cam"] == 5 and c["n_views"] == len(res.view_t0)
    assert np.allclose(np.array(c["cam_pose1"]).reshape(3, 4), res.frame_Tcw[1][:3], atol=1e-12)
    assert [l["id"] for l in c["landmarks"]] == sorted(int(i) for i in res.lm_ids)     # std::map order
    by_id = {int(i): k for k, i in enumerate(res.lm_ids)}
    for l in c["landmarks"]:
        k = by_id[l["id"]]
        assert l["rho"] == res.lm_rho[k] and l["n_obs"] == int((res.obs_landmark == k).sum())
        assert np.array_equal(l["ref_uv"], res.obs_uv[res.lm_ref_obs[k]])
        assert l["ref_t0"] == res.view_t0[res.obs_view[res.lm_ref_obs[k]]]
    row = open(csv).read().strip().split(",")
    assert row[0] == "check" and len(row) == 1 + 3 + 4 + 1 + 3 + 3 + 3
    q = np.array([0.1, -0.2, 0.3, 0.9]); q /= np.linalg.norm(q)              # (x, y, z, w) of the Quaterniond(0.9, 0.1, -0.2, 0.3) the check passes
    from lvi_exc_b200.problem import quat_conj, quat_rot
    assert np.allclose([float(x) for x in row[1:4]], quat_rot(quat_conj(q), -np.array([0.05, -0.1, 0.08])), atol=1e-5)
