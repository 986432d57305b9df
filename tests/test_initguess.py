"""Initial-guess stage (lvi_exc_b200/initguess.py) — host code.  The reference ships no vectors for it, so the checks are:
closed forms of the pre-integration, recovery of a known extrinsic from exact data (the equations are right), and the reference's own
quirks (first sample of every integrator dropped, velocity blocks at columns 0..5) reproduced where they change the answer."""
import numpy as np
import pytest

from lvi_exc_b200 import initguess as ig


def _exp(w):
    th = np.linalg.norm(w)
    K = ig._skew(w)
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K


class Traj:
    """smooth IMU trajectory in a gravity-aligned world: R(t) = Exp(a(t)), p(t) sinusoids; w_body and specific force by differentiation"""
    G = np.array([0, 0, ig.G_NORM])

    def R(self, t):
        return _exp(np.array([0.5 * np.sin(1.9 * t), 0.4 * np.sin(2.3 * t + 0.3), 0.6 * np.sin(1.7 * t + 1.0)]))

    def p(self, t):
        return np.array([1.5 * np.sin(0.8 * t), 1.2 * np.sin(1.1 * t + 0.5), 0.8 * np.sin(1.4 * t + 0.2)])

    def a(self, t):
        return np.array([-1.5 * 0.64 * np.sin(0.8 * t), -1.2 * 1.21 * np.sin(1.1 * t + 0.5), -0.8 * 1.96 * np.sin(1.4 * t + 0.2)])

    def imu(self, t, h=1e-6):
        R0, R1 = self.R(t - h), self.R(t + h)
        dR = R0.T @ R1
        w = np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]]) / (4 * h)
        return w, self.R(t).T @ (self.a(t) + self.G)


def _sensor_frames(tr, R_IS, p_IS, stamps, scale=1.0):
    def T(t):
        M = np.eye(4)
        M[:3, :3] = tr.R(t) @ R_IS
        M[:3, 3] = tr.p(t) + tr.R(t) @ p_IS
        return M
    inv0 = np.linalg.inv(T(stamps[0]))
    out = []
    for t in stamps:
        M = inv0 @ T(t)
        M[:3, 3] /= scale
        out.append(ig.IntegrationFrame(float(t), M))
    return out


def test_preintegration_constant_rates():
    """constant gyro about z and constant specific force: delta_q is the product of the first-order factors, normalised per step;
    the seeding sample contributes no time"""
    integ = ig.IntegrationBase()
    w, a, dt, n = np.array([0, 0, 0.3]), np.array([0.2, 0, 9.8]), 0.005, 40
    for _ in range(n + 1):
        integ.push_back(dt, a, w)
    assert integ.sum_dt == pytest.approx(n * dt)
    half = np.arctan(0.3 * dt / 2)          # each normalised factor (0, 0, w dt / 2, 1) is a rotation by 2 atan(w dt / 2)
    assert integ.delta_q[2] == pytest.approx(np.sin(n * half), abs=1e-12)
    assert integ.delta_q[3] == pytest.approx(np.cos(n * half), abs=1e-12)
    assert integ.delta_v[2] == pytest.approx(9.8 * n * dt, rel=1e-12)
    assert integ.delta_p[2] == pytest.approx(0.5 * 9.8 * (n * dt) ** 2, rel=1e-9)


def test_preintegration_matches_kinematics():
    """against the exact relative motion of a smooth trajectory: delta_q ~ R_i^T R_j, delta_v ~ R_i^T (v_j - v_i + g dt)"""
    tr = Traj()
    t = np.arange(0.0, 0.2 + 1e-9, 0.0005)
    integ = ig.IntegrationBase()
    for k, tk in enumerate(t):
        w, a = tr.imu(tk)
        integ.push_back(0.0005, a, w)
    Rrel = tr.R(t[0]).T @ tr.R(t[-1])
    assert np.abs(ig.quat_to_matrix(integ.delta_q) - Rrel).max() < 1e-5
    h = 1e-6
    v = lambda x: (tr.p(x + h) - tr.p(x - h)) / (2 * h)
    dv = tr.R(t[0]).T @ (v(t[-1]) - v(t[0]) + Traj.G * (t[-1] - t[0]))
    assert np.abs(integ.delta_v - dv).max() < 1e-4


def _run(tr, R_IS, p_IS, fix_scale, scale, rate=1000.0, frame_dt=0.1, n_frames=40):
    stamps = 1.0 + frame_dt * np.arange(n_frames)
    frames = _sensor_frames(tr, R_IS, p_IS, stamps, scale)
    imu_t = np.arange(0.5, stamps[-1] + 0.5, 1.0 / rate)
    wa = [tr.imu(x) for x in imu_t]
    ig.compute_integration_for_frames(frames, imu_t, np.array([w for w, _ in wa]), np.array([a for _, a in wa]))
    return ig.estimate_init_extrinsic(frames, fix_scale), frames


def test_rotation_handeye_recovers_extrinsic():
    tr = Traj()
    R_IS = _exp(np.array([0.3, -0.5, 0.8]))
    out, _ = _run(tr, R_IS, np.array([0.1, -0.2, 0.15]), True, 1.0)
    assert out.ok_rotation
    err = np.arccos(np.clip((np.trace(out.R_I_S.T @ R_IS) - 1) / 2, -1, 1))
    assert err < 2e-3            # first-order mid-point integration of a 1 kHz gyro
    assert np.allclose(ig.quat_to_matrix(out.q_StoI), out.R_I_S, atol=1e-12)


def _literal_alignment(win, fix_scale):
    """LinearAlignment written the way the reference writes it (tmp_A / tmp_b per frame pair, initial_aligment.cpp:262-423, 128-171),
    independent of initguess.py's vectorised form"""
    n = len(win)

    def gravity_system(extra, rhs_of):
        A, b = np.zeros(((n - 1) * 3, n * 3 + extra.shape[1] if extra is not None else 0)), np.zeros((n - 1) * 3)
        return A, b
    def solve_velocity(lxly, g0):
        m = 3 if lxly is None else 2
        A, b = np.zeros(((n - 1) * 3, n * 3 + m)), np.zeros((n - 1) * 3)
        for i in range(n - 2):
            Ri, Rj, pre = win[i].R, win[i + 1].R, win[i + 1].pre
            tmp_A = np.zeros((3, 6 + m))
            tmp_A[:, 0:3] = -np.eye(3)
            tmp_A[:, 3:6] = Ri.T @ Rj
            tmp_A[:, 6:] = Ri.T * pre.sum_dt if lxly is None else Ri.T @ (pre.sum_dt * np.eye(3)) @ lxly
            tmp_b = pre.delta_v.copy() if lxly is None else pre.delta_v - Ri.T @ (pre.sum_dt * g0)
            A[3 * i:3 * i + 3, 0:6] += tmp_A[:, :6]
            A[3 * i:3 * i + 3, 3 * n:] += tmp_A[:, 6:]
            b[3 * i:3 * i + 3] += tmp_b
        return np.linalg.lstsq(A.T @ A * 1000.0, A.T @ b * 1000.0, rcond=1e-14)[0]
    x = solve_velocity(None, None)
    g0 = x[3 * n:] / np.linalg.norm(x[3 * n:]) * 9.7964
    for _ in range(4):
        a = g0 / np.linalg.norm(g0)
        tmp = np.array([1.0, 0, 0]) if np.array_equal(a, [0, 0, 1.0]) else np.array([0, 0, 1.0])
        bb = tmp - a * (a @ tmp)
        bb /= np.linalg.norm(bb)
        lxly = np.column_stack([bb, np.cross(a, bb)])
        x = solve_velocity(lxly, g0)
        g0 = g0 + lxly @ x[3 * n:]
        g0 = g0 / np.linalg.norm(g0) * 9.7964
    m = 3 if fix_scale else 4
    A, b = np.zeros(((n - 1) * 3, m)), np.zeros((n - 1) * 3)
    for i in range(n - 2):
        Ri, Rj, pre = win[i].R, win[i + 1].R, win[i + 1].pre
        dt = pre.sum_dt
        A[3 * i:3 * i + 3, :3] = np.eye(3) - Ri.T @ Rj
        b[3 * i:3 * i + 3] = pre.delta_p + dt * x[3 * i:3 * i + 3] - (Ri.T * dt * dt / 2) @ g0
        if fix_scale:
            b[3 * i:3 * i + 3] -= Ri.T @ (win[i + 1].T - win[i].T)
        else:
            A[3 * i:3 * i + 3, 3] = Ri.T @ (win[i + 1].T - win[i].T)
    return g0, np.linalg.solve(A.T @ A, A.T @ b), x


@pytest.mark.parametrize("fix_scale,scale", [(True, 1.0), (False, 2.5)])
def test_linear_alignment_matches_literal_restatement(fix_scale, scale):
    """Q16 kept: only v_0 and v_1 are observed, the later velocities are exactly zero, |g| is pinned to 9.7964; gravity, translation and
    scale equal a second restatement that follows the reference's tmp_A / tmp_b code shape"""
    tr = Traj()
    R_IS, p_IS = _exp(np.array([0.3, -0.5, 0.8])), np.array([0.1, -0.2, 0.15])
    out, frames = _run(tr, R_IS, p_IS, fix_scale, scale)
    assert out.ok_rotation
    win = [ig.AlignFrame(f.T[:3, :3] @ out.R_I_S.T, f.T[:3, 3].copy(), f.integrator or ig.IntegrationBase()) for f in frames[:10]]
    ok, g, T_ext, x = ig.linear_alignment(win, fix_scale)
    g2, t2, x2 = _literal_alignment(win, fix_scale)
    assert np.linalg.norm(g) == pytest.approx(ig.G_NORM, abs=1e-9)
    assert np.allclose(g, g2, atol=1e-9) and np.allclose(T_ext, t2[:3], atol=1e-8)
    if fix_scale:
        assert ok and np.abs(x[6:30]).max() == 0.0
        assert out.ok_translation and np.allclose(out.T_I_S, T_ext)
    else:
        assert x[3] == pytest.approx(t2[3], rel=1e-8) and ok == bool(t2[3] > 0)


def test_key_pose_gate_and_time_window():
    tr = Traj()
    stamps = 1.0 + 0.1 * np.arange(30)
    fr = _sensor_frames(tr, np.eye(3), np.zeros(3), stamps)
    keys = ig.select_key_poses(stamps, [f.T for f in fr])
    assert keys[0].timestamp == stamps[0] and 2 <= len(keys) <= 30
    for a, b in zip(keys[:-1], keys[1:]):       # every kept pose moved >= 0.1 m or turned >= 5 deg from the previous kept one
        dR = a.T[:3, :3].T @ b.T[:3, :3]
        ang = np.degrees(np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1)))
        assert ang >= 5.0 - 1e-9 or np.linalg.norm(a.T[:3, 3] - b.T[:3, 3]) >= 0.1 - 1e-12
    # RemoveOverTimeFrames: frames outside the IMU span are dropped
    frames = [ig.IntegrationFrame(float(t), np.eye(4)) for t in (0.0, 1.0, 2.0, 3.0, 9.0)]
    imu_t = np.arange(0.5, 3.5, 0.01)
    ig.compute_integration_for_frames(frames, imu_t, np.zeros((len(imu_t), 3)), np.tile([0, 0, 9.8], (len(imu_t), 1)))
    assert [f.timestamp for f in frames] == [1.0, 2.0, 3.0]
    assert frames[0].integrator is None and frames[1].integrator.sum_dt == pytest.approx(0.99, abs=1e-9)


def test_sequence_initial_guess_rotation_close_translation_rough():
    """on the synthetic VLP-16 + IMU + mono sequence: both hand-eye rotations come back within half a degree from noisy LOAM poses and a
    biased IMU; the translations are what the reference's alignment gives (decimetres off) — the calibration stages start from there"""
    from lvi_exc_b200 import pipeline, synth
    seq = synth.make_sequence(synth.default_config(duration=6.0, n_landmarks=50))
    seq.scans_raw = None
    g = pipeline.estimated_initial_extrinsics(seq)
    assert pipeline.quat_angle(g["q_LtoI"], seq.gt["q_LtoI"]) < np.radians(0.5)
    assert pipeline.quat_angle(g["q_CtoI"], seq.gt["q_CtoI"]) < np.radians(0.5)
    assert np.linalg.norm(g["p_LinI"] - seq.gt["p_LinI"]) < 3.0 and np.linalg.norm(g["p_CinI"] - seq.gt["p_CinI"]) < 3.0
    assert g["scale"] > 0 and g["gravity_lidar"][3] == pytest.approx(seq.scan_times[0])


def test_cpp_vi_init_headers_match(tmp_path):
    """the C++ side of the stage (include/lvi_exc_b200/compat/vi_init/*.h: IntegrationBase, InitialEXRotation, VisualIMUAlignment under the
    reference's include paths) driven through the reference's own call sequence (tests/cpp/initguess_check.cpp restates
    ComputeIntegrationForFrames / EstimateInitExtrinsicLI / CI, T:952-1148) gives the Python stage's numbers on the synthetic sequence"""
    import json
    import struct
    import subprocess
    from pathlib import Path
    from lvi_exc_b200 import synth
    root = Path(__file__).resolve().parent.parent
    compat = root / "include" / "lvi_exc_b200" / "compat"
    exe = root / "build" / "initguess_check"
    exe.parent.mkdir(exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", str(compat), str(root / "tests" / "cpp" / "initguess_check.cpp"), "-o", str(exe)], check=True)
    seq = synth.make_sequence(synth.default_config(duration=6.0, n_landmarks=50))
    keys = ig.select_key_poses(seq.scan_times, seq.loam_poses)
    cam_t, cam_T = seq.visual_odometry()
    blob = tmp_path / "init.bin"
    with open(blob, "wb") as f:
        f.write(struct.pack("<4i", 0x4C564932, len(keys), len(cam_t), len(seq.imu_t)))
        f.write(np.array([k.timestamp for k in keys]).tobytes())
        f.write(np.stack([k.T for k in keys]).tobytes())
        f.write(np.ascontiguousarray(cam_t).tobytes())
        f.write(np.ascontiguousarray(cam_T).tobytes())
        for a in (seq.imu_t, seq.gyro, seq.accel):
            f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    r = subprocess.run([str(exe), str(blob)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    c = json.loads(r.stdout)
    p = ig.initial_extrinsics(seq.scan_times, seq.loam_poses, cam_t, cam_T, seq.imu_t, seq.gyro, seq.accel)
    for name, q, t, g in (("lidar", p["q_LtoI"], p["p_LinI"], p["gravity_lidar"]), ("camera", p["q_CtoI"], p["p_CinI"], p["gravity_cam"])):
        assert c[name]["rot_ok"] == 1 and c[name]["ok"] == 1
        qc = np.array(c[name]["q"])
        assert min(np.abs(qc - q).max(), np.abs(qc + q).max()) < 1e-9
        assert np.abs(np.array(c[name]["T"]) - t).max() < 1e-7
        assert np.abs(np.array(c[name]["g"]) - g).max() < 1e-7
    assert c["camera"]["scale"] == pytest.approx(p["scale"], rel=1e-8)
