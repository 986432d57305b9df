"""The reference's own call sequences (TrajectoryManagerLVI stage functions, LiDAROdometry::feedScan, SurfelAssociation::setSurfelMap /
getAssociation, pclomp NDT target cells, Kontiki measurements / estimator, ceres summary types) compile against
include/lvi_exc_b200/compat under the reference's include paths (tests/cpp/drop_in_check.cpp) and, on a GPU, reproduce the CPU oracle's
S0 / S1 / S4 / S5 stage sequence."""
import json
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

from lvi_exc_b200 import pipeline, synth

ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "build" / "drop_in_check"
COMPAT = ROOT / "include" / "lvi_exc_b200" / "compat"


def _build():
    src = ROOT / "tests" / "cpp" / "drop_in_check.cpp"
    lib = ROOT / "lvi_exc_b200" / "lib"
    deps = [src, lib / "liblvi_exc_b200.so", ROOT / "include" / "lvi_exc_b200.h", *COMPAT.rglob("*.h"), *COMPAT.rglob("*.hpp"), COMPAT / "Eigen" / "Dense"]
    if not EXE.exists() or EXE.stat().st_mtime < max(p.stat().st_mtime for p in deps):
        EXE.parent.mkdir(exist_ok=True)
        subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", str(COMPAT), "-I", str(ROOT / "include"), str(src), "-o", str(EXE), f"-L{lib}",
                        "-llvi_exc_b200", f"-Wl,-rpath,{lib}"], check=True)
    return EXE


def _write_blob(seq, path: Path):
    init = pipeline.perturbed_initial_extrinsics(seq.gt)
    S, H, W = seq.scans_raw.shape
    with open(path, "wb") as f:
        f.write(struct.pack("<8i", 0x4C564931, S, H, W, len(seq.imu_t), len(seq.view_t0), len(seq.obs_view), len(seq.lm_rho)))
        f.write(struct.pack("<2d", seq.map_time, seq.end_time))
        for k in ("q_LtoI", "p_LinI", "q_CtoI", "p_CinI"):
            f.write(np.ascontiguousarray(init[k], dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(seq.scans_raw).tobytes())
        for a in (seq.loam_poses, seq.imu_t, seq.gyro, seq.accel, seq.view_t0):
            f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(seq.obs_view, dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(seq.obs_landmark, dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(seq.obs_uv, dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(seq.lm_ref_obs, dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(seq.lm_rho, dtype=np.float64).tobytes())


def test_reference_call_sequences_compile_and_fail_loudly_without_a_device(tmp_path):
    """-Wall -Werror build of the translation unit; without a CUDA device the first device call throws (no CPU fallback)"""
    exe = _build()
    from lvi_exc_b200 import _capi
    if _capi.load().lvi_device_count() > 0:
        pytest.skip("a CUDA device is present")
    seq = synth.make_sequence(synth.default_config(duration=0.5, n_landmarks=50))
    blob = tmp_path / "seq.bin"
    _write_blob(seq, blob)
    r = subprocess.run([str(exe), str(blob)], capture_output=True, text=True)
    assert r.returncode == 1
    err = json.loads(r.stdout.strip().splitlines()[-1])
    assert "CUDA" in err["error"] or "device" in err["error"]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["batch", "per_scan"])
def test_drop_in_stage_sequence_matches_oracle(tmp_path, mode):
    from tests.oracle_backend import OracleBackend
    exe = _build()
    seq = synth.make_sequence(synth.default_config(duration=3.0, n_landmarks=500))
    blob = tmp_path / "seq.bin"
    _write_blob(seq, blob)
    r = subprocess.run([str(exe), str(blob)] + (["per_scan"] if mode == "per_scan" else []), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    g = json.loads(r.stdout.strip().splitlines()[-1])
    pc = pipeline.PipelineConfig(n_refine=0)
    # the C++ manager carries the reference's CalibParamManager weights (calibration.hpp:64-72), the Python one the YAML's: align them
    o = _oracle_run(seq, pc)
    co = o["calib"]
    assert [s["name"] for s in g["stages"]] == [s["name"] for s in o["stages"]]
    assert [s["iterations"] for s in g["stages"]] == [s["iterations"] for s in o["stages"]]
    for a, b in zip(g["stages"], o["stages"]):
        assert a["residuals"] == b["n_res"]
        assert a["initial_cost"] == pytest.approx(b["initial_cost"], rel=1e-5)   # map points are float32: the planes differ in the last ulp
        assert a["final_cost"] == pytest.approx(b["final_cost"], rel=2e-4), a["name"]   # S1 stops on its iteration cap, not at a minimum
    assert g["n_surfel_points"] == o["assoc_counts"][0] and g["n_lm_plane"] == o["n_lm_plane"] and g["n_planes"] > 10 and g["n_leaves"] >= g["n_planes"]
    assert pipeline.quat_angle(np.array(g["q_LtoI"]), co.q_LtoI) < 1e-4 and np.linalg.norm(np.array(g["p_LinI"]) - co.p_LinI) < 1e-3
    assert pipeline.quat_angle(np.array(g["q_CtoI"]), co.q_CtoI) < 1e-4 and np.linalg.norm(np.array(g["p_CinI"]) - co.p_CinI) < 1e-3
    assert g["pose_ok"] == 1 and g["pose_out"] == 0 and abs(g["point2plane"]) < 1.0


_ORACLE = {}


def _oracle_run(seq, pc):
    from tests.oracle_backend import OracleBackend
    if "o" not in _ORACLE:
        _ORACLE["o"] = _run_with_cpp_weights(seq, OracleBackend(), pc)
    return _ORACLE["o"]


def _run_with_cpp_weights(seq, backend, pc):
    """run_calibration with CalibParamManager's compiled-in weights (gyro 28.5, accel 18.5, lidar 10, visual-surfel 200, camera Huber 1.0)"""
    orig = pipeline.CalibParams

    class CppWeights(orig):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.w_gyr, self.w_acc, self.w_lidar, self.w_cam, self.w_visual_surfel = 28.5, 18.5, 10.0, 1.0, 200.0
    pipeline.CalibParams = CppWeights
    try:
        return pipeline.run_calibration(seq, backend, pc)
    finally:
        pipeline.CalibParams = orig
