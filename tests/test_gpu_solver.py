"""GPU parity tests of the least-squares path (SURVEY §8 a-4 … a-13) against the CPU oracle, through the C-ABI."""
import numpy as np
import pytest

from lvi_exc_b200 import pipeline, synth
from lvi_exc_b200.backend import CudaProblem
from tests import oracle_binding as ob
from tests.oracle_backend import OracleBackend
from tests.problems import make_lvi_problem, map_tangent

pytestmark = pytest.mark.gpu


def _spd_band(rng, nb, nbo, bw):
    n = nb + nbo
    A = np.zeros((n, n))
    for i in range(nb):
        for j in range(max(0, i - bw), i + 1):
            A[i, j] = A[j, i] = rng.standard_normal()
    A[nb:, :] = rng.standard_normal((nbo, n))
    A[:, nb:] = A[nb:, :].T
    A[nb:, nb:] = (A[nb:, nb:] + A[nb:, nb:].T) / 2
    A += np.eye(n) * (np.abs(A).sum(axis=1).max() + 1.0)   # diagonally dominant -> SPD
    return A


@pytest.mark.parametrize("nb,nbo,bw", [(1, 0, 0), (31, 0, 5), (32, 3, 31), (100, 0, 0), (257, 44, 23), (700, 20, 130), (64, 0, 63),
                                       (0, 7, 0), (1500, 116, 400), (333, 1, 332),
                                       # half bandwidths of 2, 3, 6, 7, 8 and 22 tiles: the factor kernel's pre-accumulation / flagged hand-off cases
                                       # and both sides of the back substitution's local window (6 tiles)
                                       (200, 5, 64), (400, 12, 96), (700, 9, 190), (900, 40, 200), (1200, 33, 250), (2400, 70, 700)])
def test_band_solver_matches_numpy(cuda_backend, nb, nbo, bw):
    rng = np.random.default_rng(nb * 1000 + nbo * 10 + bw)
    A = _spd_band(rng, nb, nbo, bw)
    rhs = rng.standard_normal(nb + nbo)
    x = cuda_backend.band_solve_dense(A, rhs, nb, nbo, bw)
    ref = np.linalg.solve(A, rhs)
    assert np.abs(x - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("n0,n1,nbo,n_mid,bw", [(64, 64, 10, 10, 20), (320, 288, 70, 40, 45), (96, 640, 44, 0, 23), (1024, 992, 200, 150, 130),
                                                (32, 32, 3, 3, 31), (1600, 1568, 300, 256, 260), (96, 96, 70, 64, 64)])
def test_two_sided_band_solver_matches_numpy(cuda_backend, n0, n1, nbo, n_mid, bw):
    """two chains [0, n0) and [n0, n0+n1) that only meet in the border; the first n_mid border dims go through the second-level system"""
    rng = np.random.default_rng(n0 + 7 * n1 + nbo)
    nb = n0 + n1
    A = _spd_band(rng, nb, nbo, bw)
    A[n0:nb, :n0] = 0.0
    A[:n0, n0:nb] = 0.0
    # separator rows couple only to the last bw positions of each chain (that is what makes them a separator)
    far = np.ones(nb, bool); far[max(0, n0 - bw):n0] = False; far[max(n0, nb - bw):nb] = False
    A[nb:nb + n_mid, :nb][:, far] = 0.0
    A[:nb, nb:nb + n_mid][far, :] = 0.0
    rhs = rng.standard_normal(nb + nbo)
    x = cuda_backend.band_solve_dense(A, rhs, nb, nbo, bw, chain1_start=n0, n_mid=n_mid)
    ref = np.linalg.solve(A, rhs)
    assert np.abs(x - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())


def test_band_solver_reports_breakdown(cuda_backend):
    A = np.eye(40)
    A[5, 5] = -1.0
    with pytest.raises(Exception):
        cuda_backend.band_solve_dense(A, np.ones(40), 40, 0, 0)


@pytest.mark.parametrize("stage", ["so3", "surfel", "lvi", "lvi_locked", "lvi_dist", "lvi_locked_dist"])
def test_evaluate_matches_oracle(cuda_backend, stage):
    pd_g, pd_o = make_lvi_problem(stage), make_lvi_problem(stage)
    gp, op = CudaProblem(cuda_backend, pd_g), ob.OracleProblem(pd_o)
    assert gp.num_residuals == op.num_residuals
    eg, eo = gp.evaluate(jacobian=True), op.evaluate(jacobian=True)
    assert abs(eg["cost"] - eo["cost"]) <= 1e-10 * max(1.0, eo["cost"])
    scale = max(1.0, np.abs(eo["residuals"]).max())
    assert np.abs(eg["residuals"] - eo["residuals"]).max() <= 1e-9 * scale
    perm = map_tangent(cuda_backend, gp, op, pd_g)   # library tangent position -> oracle tangent offset
    real = perm >= 0
    assert real.sum() == op.num_tangent
    Jo = eo["J"][:, perm[real]]
    assert np.abs(eg["J"][:, real] - Jo).max() <= 1e-7 * max(1.0, np.abs(Jo).max()) and not eg["J"][:, ~real].any()
    assert np.abs(eg["gradient"][real] - eo["gradient"][perm[real]]).max() <= 1e-7 * max(1.0, np.abs(eo["gradient"]).max())


@pytest.mark.parametrize("stage,iters", [("so3", 30), ("surfel", 12), ("lvi", 10), ("lvi_locked", 10), ("lvi_dist", 10)])
def test_solve_matches_oracle(cuda_backend, stage, iters):
    pd_g, pd_o = make_lvi_problem(stage), make_lvi_problem(stage)
    sg = CudaProblem(cuda_backend, pd_g).solve(iters)
    so = ob.OracleProblem(pd_o).solve(iters)
    assert sg.num_iterations == so.num_iterations
    assert sg.termination_type == so.termination_type
    assert abs(sg.initial_cost - so.initial_cost) <= 1e-9 * so.initial_cost
    assert abs(sg.final_cost - so.final_cost) <= 1e-6 * so.final_cost
    n = min(sg.n_log, so.n_log)
    assert list(sg.log_successful[:n]) == list(so.log_successful[:n])
    # parameters: the north-star tolerance is 1e-4 rad / 1e-3 m on the extrinsics; identical LM paths give far tighter agreement
    assert pipeline.quat_angle(pd_g.lidar_q, pd_o.lidar_q) < 1e-6 and np.abs(pd_g.lidar_p - pd_o.lidar_p).max() < 1e-6
    assert pipeline.quat_angle(pd_g.cam_q, pd_o.cam_q) < 1e-6 and np.abs(pd_g.cam_p - pd_o.cam_p).max() < 1e-6
    assert np.abs(pd_g.so3_knots - pd_o.so3_knots).max() < 1e-5
    if pd_g.r3_knots is not None:
        assert np.abs(pd_g.r3_knots - pd_o.r3_knots).max() < 1e-5
    assert np.abs(pd_g.gyr_bias - pd_o.gyr_bias).max() < 1e-7 and np.abs(pd_g.acc_bias - pd_o.acc_bias).max() < 1e-6
    if len(pd_g.rho):
        assert np.abs(pd_g.rho - pd_o.rho).max() < 1e-5


def test_time_out_of_range_raises(cuda_backend):
    pd = make_lvi_problem("surfel")
    t, w, wt = pd.tables["gyro"]
    t = t.copy(); t[0] = pd.max_time + 1.0
    pd.tables["gyro"] = (t, w, wt)
    with pytest.raises(IndexError):   # std::range_error in the reference (K/trajectory_estimator.h:111-122)
        CudaProblem(cuda_backend, pd)


def test_non_unit_control_point_raises(cuda_backend):
    pd = make_lvi_problem("so3")
    pd.so3_knots[3] *= 1.1
    with pytest.raises(ValueError):   # std::domain_error (K/trajectories/uniform_so3_spline_trajectory.h:23-27)
        CudaProblem(cuda_backend, pd)


def test_bench_iterations_leave_parameters_untouched(cuda_backend):
    pd = make_lvi_problem("lvi")
    before = pd.clone_params()
    prob = CudaProblem(cuda_backend, pd)
    ms = prob.bench_iterations(2)
    assert (ms >= 0).all() and ms.sum() > 0
    e1 = prob.evaluate(gradient=False)
    pd2 = make_lvi_problem("lvi")
    e2 = CudaProblem(cuda_backend, pd2).evaluate(gradient=False)
    assert e1["cost"] == pytest.approx(e2["cost"], rel=1e-12)
    for k, v in before.items():
        if v is not None:
            assert np.array_equal(getattr(pd, k), v)
