"""The C++ host facade (include/lvi_exc_b200/kontiki_facade.hpp: Kontiki's names over the C-ABI) compiles against the header,
records measurements into the flat tables the C-ABI takes, fails loudly without a device, and (GPU) solves through the library."""
import json
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "build" / "facade_check"


def _build():
    src = ROOT / "tests" / "cpp" / "facade_check.cpp"
    lib = ROOT / "lvi_exc_b200" / "lib"
    hdr = ROOT / "include" / "lvi_exc_b200" / "kontiki_facade.hpp"
    if not EXE.exists() or EXE.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime, (lib / "liblvi_exc_b200.so").stat().st_mtime):
        EXE.parent.mkdir(exist_ok=True)
        subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I", str(ROOT / "include"), str(src), "-o", str(EXE), f"-L{lib}", "-llvi_exc_b200",
                        f"-Wl,-rpath,{lib}"], check=True)
    return EXE


def test_facade_records_reference_measurements():
    out = subprocess.run([str(_build()), "describe"], capture_output=True, text=True, check=True).stdout
    d = json.loads(out.strip().splitlines()[-1])
    assert (d["n_gyro"], d["n_accel"], d["n_surfel"], d["n_cam"], d["n_camsurf"]) == (10, 10, 6, 1, 1) and d["blocks"] == 28
    assert d["n_knots"] == 73                                    # 1 s + 2 x 0.2 s padding at dt 0.02
    assert d["n_planes"] == 2 and d["surfel_plane_3"] == 1 and d["cs_plane"] == 1 and d["plane1_y"] == 2.0   # plane pointers de-duplicated
    assert d["n_landmarks"] == 1 and d["rho0"] == 0.25 and abs(d["cam_t0_ref"] - 0.00411) < 1e-9
    assert (d["lock_lidar_q"], d["lock_cam_q"], d["lock_acc_bias"], d["lock_gyr_bias"]) == (0, 1, 0, 1)         # sensors locked by default
    assert d["cam_weight"] == 1.0                                # 3-argument ctor: (camera, obs, huber) -> weight 1 (Q3)
    assert abs(d["min_time"] + 0.2) < 1e-12 and d["max_time"] >= 1.2


@pytest.mark.gpu
def test_facade_solves_on_the_gpu():
    out = subprocess.run([str(_build()), "solve"], capture_output=True, text=True, check=True).stdout
    d = json.loads(out.strip().splitlines()[-1])
    assert d["usable"] == 1 and d["range_error"] == 1
    assert d["final_cost"] < 1e-9 * max(1.0, d["initial_cost"])
    assert abs(d["yaw"] - d["expected_yaw"]) < 1e-6


def test_compat_pinhole_camera_distortion_round_trip():
    """PinholeCamera::Project / Unproject of the compat headers with the radial-tangential model (K/sensors/pinhole_camera.h:96-240): the 8-step
    inverse brings a projected point back to its ray; without coefficients it is K X / z and K^-1 (u, v, 1); and the numbers equal the oracle's
    restatement of the same functions (through a one-residual evaluation they would be buried in, so here: a numpy restatement)"""
    import json
    import subprocess
    import numpy as np
    root = Path(__file__).resolve().parent.parent
    compat = root / "include" / "lvi_exc_b200" / "compat"
    lib = root / "lvi_exc_b200" / "lib"
    exe = root / "build" / "camera_check"
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", str(compat), "-I", str(root / "include"), str(root / "tests" / "cpp" / "camera_check.cpp"),
                    "-o", str(exe), f"-L{lib}", "-llvi_exc_b200", f"-Wl,-rpath,{lib}"], check=True)
    out = json.loads(subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout)
    assert out["do_distortion"] == [1, 0]
    k1, k2, p1, p2, k3 = -0.28, 0.07, 1.0e-3, -5.0e-4, 0.01
    fx, fy, cx, cy = 530.175, 530.095, 635.12, 356.522
    pts = [(0.3, -0.2, 2.0), (-1.1, 0.4, 3.0), (0.05, 0.6, 1.5), (0.0, 0.0, 4.0)]
    for X, o in zip(pts, out["points"]):
        x, y = X[0] / X[2], X[1] / X[2]
        r2 = x * x + y * y
        rad = k1 * r2 + k2 * r2 ** 2 + k3 * r2 ** 3
        xd = x + x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
        yd = y + y * rad + 2 * p2 * x * y + p1 * (r2 + 2 * y * y)
        assert np.allclose(o["y"], [fx * xd + cx, fy * yd + cy], rtol=0, atol=1e-9)
        assert np.allclose(o["y_plain"], [fx * x + cx, fy * y + cy], rtol=0, atol=1e-9)
        assert np.allclose(o["ray"], [x, y, 1.0], atol=2e-8)            # 8 fixed-point steps: converged to ~1e-9 here, not to round-off
        assert np.allclose(o["ray_plain"], [x, y, 1.0], atol=1e-14)
