"""GPU parity tests at the sizes BASELINE.json's configs name, through the C-ABI.

  C1   10 scans + 200 Hz IMU, surfel + IMU residuals only: the whole LI stage sequence against the oracle run live
  S4   20 s of the C2 workload: the full LVI problem with the REAL half bandwidth (701) and the 702-dim separator of the two-sided
       ordering (the lowering splits the band into two chains from ~19 s of data on), evaluate + LM iteration log against the oracle run live
  C2   60 s, the configuration bench.py is quoted on: whole stage sequence against the oracle's committed result
  C5   60 s of degenerate motion (planar, low excitation): iteration counts, final costs and extrinsics against the oracle's committed
       result -- parity, not accuracy, is the criterion (p_LinI.z is weakly observable)

The oracle needs minutes for a 60 s stage sequence, so C2 / C5 compare against tests/golden/oracle_calibration_c{2,5}.json, written by
tools/oracle_calibration.py (committed next to them).  Tolerance on the extrinsics: the north star's 1e-4 rad / 1e-3 m.
"""
import copy
import ctypes as C
import json
from pathlib import Path

import numpy as np
import pytest

from lvi_exc_b200 import _capi, pipeline, synth, workload
from lvi_exc_b200.backend import CudaProblem
from tests import oracle_binding as ob
from tests.oracle_backend import OracleBackend
from tests.problems import map_tangent

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
TOL_RAD, TOL_M = 1e-4, 1e-3


def _assert_extrinsics(cg, co):
    q = lambda v: np.asarray(v, dtype=np.float64)
    assert pipeline.quat_angle(cg.q_LtoI, q(co["q_LtoI"])) < TOL_RAD and np.linalg.norm(cg.p_LinI - q(co["p_LinI"])) < TOL_M
    assert pipeline.quat_angle(cg.q_CtoI, q(co["q_CtoI"])) < TOL_RAD and np.linalg.norm(cg.p_CinI - q(co["p_CinI"])) < TOL_M


def test_c1_li_sequence_matches_oracle(cuda_backend):
    """C1 exactly as SURVEY §8(d) defines it: 1.0 s, 10 scans (288,000 raw points), 200 IMU samples, surfel + IMU residuals only"""
    cfg = synth.default_config(duration=1.0, n_landmarks=0)
    seq = synth.make_sequence(cfg, with_camera=False)
    assert seq.scans_raw.shape == (10, 16, 1800)
    assert ((seq.imu_t >= seq.map_time) & (seq.imu_t < seq.end_time)).sum() == 200   # 200 Hz over the 1.0 s of LiDAR data (+ the padding)
    pc = pipeline.PipelineConfig(with_camera=False)
    og = pipeline.run_calibration(seq, cuda_backend, pc)
    oo = pipeline.run_calibration(seq, OracleBackend(), pc)
    assert og["assoc_counts"] == oo["assoc_counts"] and min(og["assoc_counts"]) > 500
    assert [s["name"] for s in og["stages"]] == ["S0_so3", "S1_surfel", "S2_refine", "S3_refine"]
    assert [s["iterations"] for s in og["stages"]] == [s["iterations"] for s in oo["stages"]]
    assert [s["termination"] for s in og["stages"]] == [s["termination"] for s in oo["stages"]]
    for a, b in zip(og["stages"], oo["stages"]):
        assert a["n_res"] == b["n_res"]
        assert a["initial_cost"] == pytest.approx(b["initial_cost"], rel=1e-6)   # later stages start from the previous optimum
        assert a["final_cost"] == pytest.approx(b["final_cost"], rel=1e-5)
    cg, co = og["calib"], oo["calib"]
    _assert_extrinsics(cg, {k: getattr(co, k) for k in ("q_LtoI", "p_LinI", "q_CtoI", "p_CinI")})
    assert np.abs(cg.gyr_bias - co.gyr_bias).max() < 1e-6 and np.abs(cg.acc_bias - co.acc_bias).max() < 1e-5


def test_s4_real_bandwidth_matches_oracle(cuda_backend):
    """stage S4 on 20 s of the C2 workload: half bandwidth 701 + 702-dim separator (the shape of the 60 s benchmark problem)"""
    seq = synth.make_sequence(synth.default_config(duration=20.0))
    pd_g, _ = workload.lvi_stage_problem(seq, cuda_backend)
    pd_o = copy.deepcopy(pd_g)   # the oracle gets its own parameter memory (both solvers update in place)
    gp, op = CudaProblem(cuda_backend, pd_g), ob.OracleProblem(pd_o)
    lay = np.zeros(8, np.int32)
    _capi.check(cuda_backend.lib.lvi_problem_layout(gp.h, lay.ctypes.data_as(C.POINTER(C.c_int32))))
    assert lay[2] == 701, f"half bandwidth {lay[2]}"                    # reference -> last observation of a 2.2 s track, + Schur fill
    assert lay[6] < lay[0] and lay[1] >= 702                            # two chains, separator of >= 702 dims in the border
    assert gp.num_residuals == op.num_residuals
    eg, eo = gp.evaluate(jacobian=False), op.evaluate(jacobian=False)
    assert abs(eg["cost"] - eo["cost"]) <= 1e-10 * eo["cost"]
    assert np.abs(eg["residuals"] - eo["residuals"]).max() <= 1e-9 * max(1.0, np.abs(eo["residuals"]).max())
    perm = map_tangent(cuda_backend, gp, op, pd_g)
    real = perm >= 0
    assert np.abs(eg["gradient"][real] - eo["gradient"][perm[real]]).max() <= 1e-7 * np.abs(eo["gradient"]).max()
    iters = 6
    sg, so = gp.solve(iters), op.solve(iters)
    assert sg.num_iterations == so.num_iterations and sg.termination_type == so.termination_type
    n = min(sg.n_log, so.n_log)
    assert n >= 2 and list(sg.log_successful[:n]) == list(so.log_successful[:n])
    for k in range(n):   # the iteration log: cost, trust-region radius and step norm of every LM iteration
        assert sg.log_cost[k] == pytest.approx(so.log_cost[k], rel=1e-7)
        assert sg.log_radius[k] == pytest.approx(so.log_radius[k], rel=1e-5)
        assert sg.log_step_norm[k] == pytest.approx(so.log_step_norm[k], rel=1e-4, abs=1e-9)
    assert pipeline.quat_angle(pd_g.lidar_q, pd_o.lidar_q) < 1e-6 and np.abs(pd_g.lidar_p - pd_o.lidar_p).max() < 1e-6
    assert pipeline.quat_angle(pd_g.cam_q, pd_o.cam_q) < 1e-6 and np.abs(pd_g.cam_p - pd_o.cam_p).max() < 1e-6
    assert np.abs(pd_g.r3_knots - pd_o.r3_knots).max() < 1e-5 and np.abs(pd_g.rho - pd_o.rho).max() < 1e-5


def _against_fixture(cuda_backend, name: str, degenerate: bool, strict: bool = True):
    g = json.loads((GOLDEN / name).read_text())
    assert g["config"]["seconds"] == 60.0 and g["config"]["degenerate"] == degenerate
    seq = synth.make_sequence(synth.default_config(duration=60.0, degenerate=int(degenerate)))
    og = pipeline.run_calibration(seq, cuda_backend)
    # The first association sees identical inputs on both sides and must agree exactly.  The later ones de-skew 17 M points with the
    # trajectory the previous solve returned; the two solvers agree on it to ~1e-9, which moves a float coordinate by an ulp now and
    # then, and a point within an ulp of a voxel face / box face / the 0.05 m radius may then fall on the other side: a handful of the
    # ~1.1 M hits may differ (tests/test_gpu_map.py and tools/diag/map_parity_c2.py check the map path bit for bit on identical inputs).
    # One flipped point can change which points a leaf's RANSAC samples (the sample sequence is by index), hence a whole plane and a few dozen
    # hits of the NEXT pass: seen once in ~15 runs (110,854 / 110,840 hits against 110,855 / 110,882).  The hit counts are therefore held to
    # 0.1 %; what the north star asks for -- iteration counts, terminations, extrinsics within 1e-4 rad / 1e-3 m -- stays exact.
    assert og["assoc_counts"][0] == g["assoc_counts"][0]
    assert all(abs(a - b) <= max(3, 1e-3 * b) for a, b in zip(og["assoc_counts"], g["assoc_counts"])), (og["assoc_counts"], g["assoc_counts"])
    if strict:
        assert abs(og.get("n_lm_plane") - g["n_lm_plane"]) <= 2
        assert [s["iterations"] for s in og["stages"]] == [s["iterations"] for s in g["stages"]]
        assert [s["termination"] for s in og["stages"]] == [s["termination"] for s in g["stages"]]
        for a, b in zip(og["stages"], g["stages"]):
            assert abs(a["n_res"] - b["n_res"]) <= max(3, 1e-3 * b["n_res"])
            assert a["initial_cost"] == pytest.approx(b["initial_cost"], rel=1e-3), a["name"]
            assert a["final_cost"] == pytest.approx(b["final_cost"], rel=1e-3), a["name"]
        _assert_extrinsics(og["calib"], g["calib"])
    return og, g


def test_c2_stage_sequence_matches_oracle_fixture(cuda_backend):
    """the benchmark configuration itself: 60 s, 17.28 M points, 246,702 residuals in S4"""
    og, g = _against_fixture(cuda_backend, "oracle_calibration_c2.json", False)
    e = pipeline.extrinsic_errors(og["calib"], synth.gt_extrinsics())
    assert e["rot_L"] < 1e-3 and e["pos_L"] < 1e-2 and e["rot_C"] < 1e-3 and e["pos_C"] < 1e-2   # and it calibrates


def test_c5_degenerate_motion_matches_oracle_fixture(cuda_backend):
    """C5: z = 0, roll = pitch = 0, yaw 0.1 sin(0.2 t), x/y amplitudes x0.3 -- convergence behaviour must be the CPU path's.
    Planar motion leaves p_LinI.z (and with it the camera position) unobservable: the CPU path drifts 98 m / 945 m along the flat
    directions in 80 + 30 iterations that never converge, so positions along them cannot be compared; what must agree is the LI part the
    data does determine -- every stage up to S3: iteration counts, termination, costs, the LiDAR rotation -- and the S4 / S5 iteration
    counts and termination types."""
    og, g = _against_fixture(cuda_backend, "oracle_calibration_c5.json", True, strict=False)
    names = [s["name"] for s in g["stages"]]
    assert [s["name"] for s in og["stages"]] == names
    for a, b in zip(og["stages"], g["stages"]):
        assert a["iterations"] == b["iterations"], a["name"]
        assert a["termination"] == b["termination"], a["name"]
        if a["name"] in ("S0_so3", "S1_surfel", "S2_refine", "S3_refine"):
            assert a["initial_cost"] == pytest.approx(b["initial_cost"], rel=1e-3), a["name"]
            assert a["final_cost"] == pytest.approx(b["final_cost"], rel=1e-3), a["name"]
            assert a["errors"]["rot_L"] == pytest.approx(b["errors"]["rot_L"], abs=1e-3), a["name"]   # the flat directions leak into the rotation too
    # the degenerate direction really is degenerate on both sides
    assert og["stages"][3]["errors"]["pos_L"] > 1.0 and g["stages"][3]["errors"]["pos_L"] > 1.0
