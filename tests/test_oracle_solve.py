"""CPU tests that PIN the solve half of the oracle (the reference ships no tests, so the oracle is validated by
finite differences, spline identities, known minimisers and an independent scipy optimiser) and check the product's
host-compiled lowering + analytic Jacobians (liblvi_hostcheck.so) against it."""
import numpy as np
import pytest

from lvi_exc_b200 import pipeline, synth
from lvi_exc_b200.problem import ProblemData, quat_from_axis_angle, quat_mul
from tests import hostcheck_binding as hc
from tests import oracle_binding as ob
from tests.problems import make_lvi_problem

STAGES = ["so3", "surfel", "lvi", "lvi_locked", "lvi_dist", "lvi_locked_dist"]


def _rand_traj(n=12, seed=0):
    rng = np.random.default_rng(seed)
    r3 = np.cumsum(0.05 * rng.standard_normal((n, 3)), axis=0)
    so3 = np.zeros((n, 4)); q = np.array([0, 0, 0, 1.0])
    for i in range(n):
        q = quat_mul(q, quat_from_axis_angle(rng.standard_normal(3), 0.05 * rng.standard_normal()))
        so3[i] = q
    return ProblemData(1.0, 0.02, n, r3, so3)


def test_spline_identities():
    pd = _rand_traj()
    # constant control points -> constant pose, zero velocity / acceleration / angular velocity
    c = ProblemData(1.0, 0.02, 8, np.tile([1.0, -2.0, 3.0], (8, 1)), np.tile(quat_from_axis_angle([1, 2, 3], 0.7), (8, 1)))
    for t in (1.0, 1.013, 1.0999):
        e = ob.traj_eval(c, t)
        assert np.allclose(e["p"], [1, -2, 3], atol=1e-14) and np.allclose(e["v"], 0, atol=1e-12) and np.allclose(e["a"], 0, atol=1e-9)
        assert np.allclose(e["q"], c.so3_knots[0], atol=1e-14) and np.allclose(e["w"], 0, atol=1e-12)
    # C2 continuity across a knot boundary and derivative consistency (central differences)
    tb = 1.0 + 3 * 0.02
    lo, hi = ob.traj_eval(pd, tb - 1e-9), ob.traj_eval(pd, tb + 1e-9)
    for k in ("p", "v", "q", "w"):
        assert np.allclose(lo[k], hi[k], atol=1e-6)
    t, h = 1.0517, 1e-6
    em, e0, ep = ob.traj_eval(pd, t - h), ob.traj_eval(pd, t), ob.traj_eval(pd, t + h)
    assert np.allclose((ep["p"] - em["p"]) / (2 * h), e0["v"], atol=1e-6)
    assert np.allclose((ep["v"] - em["v"]) / (2 * h), e0["a"], atol=1e-4)
    dq = quat_mul((ep["q"] - em["q"]) / (2 * h), pipeline.quat_conj(e0["q"]))
    assert np.allclose(2 * dq[:3], e0["w"], atol=1e-6)     # world angular velocity = 2 (dq/dt q^-1).vec
    ob.traj_eval(pd, pd.max_time)                          # Q6: SplineView::Evaluate retries at t - 1e-5 (K/trajectories/spline_base.h:200-203)
    with pytest.raises(IndexError):
        ob.traj_eval(pd, pd.max_time + 1e-4)               # beyond the retry window: std::range_error


def test_product_spline_matches_oracle():
    pd = _rand_traj(16, 3)
    for t in np.linspace(pd.min_time, pd.max_time - 1e-9, 41):
        eo, eh = ob.traj_eval(pd, float(t)), hc.traj_eval(pd, float(t))
        assert np.allclose(eo["p"], eh["p"], atol=1e-14) and np.allclose(eo["a"], eh["a"], atol=1e-9) and np.allclose(eo["q"], eh["q"], atol=1e-14)
        w_body = pipeline.quat_rot(pipeline.quat_conj(eo["q"]), eo["w"])
        assert np.allclose(w_body, eh["w_body"], atol=1e-11)


@pytest.mark.parametrize("stage", STAGES)
def test_oracle_jacobian_vs_finite_differences(stage):
    """forward-mode Jets of the oracle against central differences through Plus() on a random subset of tangent directions"""
    pd = make_lvi_problem(stage, 1.0, 300)
    op = ob.OracleProblem(pd)
    e0 = op.evaluate(jacobian=True)
    nt = op.num_tangent
    rng = np.random.default_rng(7)
    saved = pd.clone_params()
    # tangent offsets -> (array, index, is_quat)
    targets = []
    for which, (name, dm) in enumerate([("lidar_q", 3), ("lidar_p", 3), ("cam_q", 3), ("cam_p", 3), ("gravity", 2), ("acc_bias", 3), ("gyr_bias", 3)]):
        off = op.offset_block(which)
        if off >= 0:
            targets.append((off, getattr(pd, name), None, name.endswith("_q")))
    for i in rng.choice(pd.n_knots, 6, replace=False):
        if pd.r3_knots is not None and op.offset_knot(int(i), False) >= 0:
            targets.append((op.offset_knot(int(i), False), pd.r3_knots, int(i), False))
        if op.offset_knot(int(i), True) >= 0:
            targets.append((op.offset_knot(int(i), True), pd.so3_knots, int(i), True))
    for l in range(min(len(pd.rho), 5)):
        if op.offset_block(7 + l) >= 0:
            targets.append((op.offset_block(7 + l), pd.rho, l, False))
    assert targets
    # the Huber corrector scales rows by sqrt(rho'), itself a function of the parameters; FD sees that too, so compare on residual
    # blocks inside the quadratic region only (|r| below the Huber delta) -> rows where scaling is 1
    res0 = e0["residuals"]
    h = 1e-6
    checked = 0
    for off, arr, idx, is_q in targets:
        view = arr if idx is None else arr[idx]
        dim = 3 if is_q else (view.size if view.ndim else 1)
        for c in range(dim):
            def set_plus(step):
                pd.restore_params(saved)
                if is_q:
                    d = np.zeros(3); d[c] = step
                    n = np.linalg.norm(d)
                    dq = np.array([*(np.sin(n) / n * d), np.cos(n)])
                    new = quat_mul(dq, view.copy())
                    if idx is None: arr[...] = new
                    else: arr[idx] = new
                else:
                    if idx is None: arr[c] += step
                    elif arr.ndim == 1: arr[idx] += step
                    else: arr[idx, c] += step
            set_plus(h); rp = op.evaluate(gradient=False)["residuals"]
            set_plus(-h); rm = op.evaluate(gradient=False)["residuals"]
            fd = (rp - rm) / (2 * h)
            col = e0["J"][:, off + c]
            quad = np.abs(res0) < 4.0
            assert np.abs(fd[quad] - col[quad]).max() <= 1e-5 * max(1.0, np.abs(col).max()), (off, c)
            checked += 1
    pd.restore_params(saved)
    assert checked >= 10 and nt > 0


@pytest.mark.parametrize("stage", STAGES)
def test_product_jacobian_matches_oracle(stage):
    """the analytic Jacobians + lowering the CUDA kernels are built from (host-compiled) vs the oracle's autodiff"""
    pd = make_lvi_problem(stage, 1.0, 300)
    op = ob.OracleProblem(pd)
    eo = op.evaluate(jacobian=True)
    eh = hc.evaluate(pd)
    assert eh["layout"]["n_res"] == op.num_residuals and eh["layout"]["nt"] - eh["layout"]["n_pad"] == op.num_tangent
    assert eh["cost"] == pytest.approx(eo["cost"], rel=1e-12) and eh["fixed_cost"] == pytest.approx(eo["fixed_cost"], rel=1e-12, abs=1e-9)
    assert np.abs(eh["residuals"] - eo["residuals"]).max() <= 1e-10 * max(1.0, np.abs(eo["residuals"]).max())
    perm = hc.perm_to_oracle(pd, eh["layout"], op)
    real = perm >= 0
    Jo = eo["J"][:, perm[real]]
    assert np.abs(eh["J"][:, real] - Jo).max() <= 1e-9 * max(1.0, np.abs(Jo).max())
    assert not eh["J"][:, ~real].any()                      # padding positions carry nothing


def test_layout_is_band_plus_arrow():
    pd = make_lvi_problem("surfel", 1.0, 300)
    lay = hc.layout(pd)
    # arrow border = separator of the two-sided ordering (4 knots x 6) + 4 map-time knots x 6 + lidar q,p (6) + gravity (2) + biases (6)
    assert lay["n_mid"] == 24 and lay["nbo"] - lay["n_mid"] == 24 + 6 + 2 + 6
    assert lay["bw"] == 23                      # 4 consecutive knots x 6 dims
    assert lay["chain1_start"] % 32 == 0 and 0 < lay["chain1_start"] < lay["nb"] and lay["n_pad"] < 32
    pd4 = make_lvi_problem("lvi")
    l4 = hc.layout(pd4)
    assert l4["nbo"] - l4["n_mid"] == 24 + 12 + 2 + 6 and l4["bw"] > 100   # camera residuals couple reference and observation windows
    assert l4["n_mid"] == 0 and l4["chain1_start"] == l4["nb"]               # too short for a separator of ~100 knots: single chain
    pd5 = make_lvi_problem("lvi_locked", 1.0, 300)
    l5 = hc.layout(pd5)
    assert (l5["pos_r3"] < 0).all() and (l5["pos_so3"] < 0).all() and l5["bw"] == 0   # Q15: only camera, biases, gravity, rho move


def test_lm_recovers_known_minimiser():
    """noise-free gyro data generated FROM a spline: the S0 problem has a zero-cost solution which LM must find"""
    pd_true = _rand_traj(40, 5)
    t = np.linspace(pd_true.min_time + 1e-4, pd_true.max_time - 1e-4, 300)
    w = np.zeros((len(t), 3))
    for i, ti in enumerate(t):
        e = ob.traj_eval(pd_true, float(ti))
        w[i] = pipeline.quat_rot(pipeline.quat_conj(e["q"]), e["w"])
    q0 = ob.traj_eval(pd_true, float(t[0]))["q"]
    pd = ProblemData(pd_true.t0, pd_true.dt, pd_true.n_knots, None, np.tile([0, 0, 0, 1.0], (pd_true.n_knots, 1)), locks=dict(lock_r3=1))
    pd.set_gyro(t, w, 28.0)
    pd.set_orientation([t[0]], [q0], 28.0)
    s = ob.OracleProblem(pd).solve(50)
    assert s.termination_type == 0 and s.final_cost < 1e-12 * max(1.0, s.initial_cost)
    for ti in t[::37]:
        assert pipeline.quat_angle(ob.traj_eval(pd_true, float(ti))["q"], _so3_only_eval(pd, float(ti))) < 1e-6


def _so3_only_eval(pd, t):
    full = ProblemData(pd.t0, pd.dt, pd.n_knots, np.zeros((pd.n_knots, 3)), pd.so3_knots)
    return ob.traj_eval(full, t)["q"]


def test_lm_matches_scipy_on_small_problem():
    """independent optimiser (scipy trust-region-reflective on the oracle's residuals) reaches the same optimum"""
    scipy_opt = pytest.importorskip("scipy.optimize")
    cfg = synth.default_config(duration=0.3, n_landmarks=0, gyro_noise=1e-3)
    seq = synth.make_sequence(cfg, with_camera=False)
    mgr = pipeline.TrajectoryManager(pipeline.CameraIntrinsics(), seq.map_time, seq.end_time, 0.02, 0.2)
    mgr.feed_imu(seq.imu_t, seq.gyro, seq.accel)
    pd = mgr.problem_so3()
    op = ob.OracleProblem(pd)
    nt = op.num_tangent
    base = pd.so3_knots.copy()
    offs = [op.offset_knot(i, True) for i in range(pd.n_knots)]

    def residuals(x):
        for i, o in enumerate(offs):
            if o < 0:
                continue
            d = x[o:o + 3]; n = np.linalg.norm(d)
            dq = np.array([*(np.sinc(n / np.pi) * d), np.cos(n)])
            pd.so3_knots[i] = quat_mul(dq, base[i])
        return op.evaluate(gradient=False)["residuals"]

    sol = scipy_opt.least_squares(residuals, np.zeros(nt), method="trf", xtol=1e-14, ftol=1e-14, gtol=1e-12)
    knots_scipy = pd.so3_knots.copy()
    cost_scipy = 0.5 * np.sum(sol.fun ** 2)
    pd.so3_knots[...] = base
    s = ob.OracleProblem(pd).solve(60, function_tolerance=1e-14, parameter_tolerance=1e-14)
    assert s.final_cost == pytest.approx(cost_scipy, rel=1e-6)
    used = np.array(offs) >= 0
    # gravity blocks carry no information in S0 (zero Jacobian) and unused end knots stay put; compare the observable knots
    ang = [pipeline.quat_angle(a, b) for a, b in zip(pd.so3_knots[used][2:-2], knots_scipy[used][2:-2])]
    assert max(ang) < 1e-5


def test_time_span_errors_match_reference_semantics():
    pd = make_lvi_problem("surfel", 1.0, 300)
    t, tm, pt, pl, w, hb = pd.tables["surfel"]
    bad = t.copy(); bad[0] = pd.max_time
    pd.tables["surfel"] = (bad, tm, pt, pl, w, hb)
    with pytest.raises(IndexError):
        ob.OracleProblem(pd)
    with pytest.raises(IndexError):
        hc.layout(pd)
    pd = make_lvi_problem("surfel", 1.0, 300)
    t, tm, pt, pl, w, hb = pd.tables["surfel"]
    tm2 = tm.copy(); tm2[3] = t[3] + 0.1          # spans not ordered (Q12)
    pd.tables["surfel"] = (t, tm2, pt, pl, w, hb)
    with pytest.raises(IndexError):
        ob.OracleProblem(pd)
    with pytest.raises(IndexError):
        hc.layout(pd)


def test_product_jacobian_matches_oracle_on_a_larger_problem():
    """5 s of data: more than 4096 surfel rows, so the surfel table is lowered in two shares (helper thread + calling thread, lowering.hpp),
    and the half bandwidth comes from the per-window accumulation of the Schur rows"""
    pd = make_lvi_problem("lvi", 5.0, 3000)
    assert len(pd.tables["surfel"][0]) > 4096
    op = ob.OracleProblem(pd)
    eo, eh = op.evaluate(jacobian=True), hc.evaluate(pd)
    assert eh["layout"]["n_res"] == op.num_residuals
    perm = hc.perm_to_oracle(pd, eh["layout"], op)
    real = perm >= 0
    Jo = eo["J"][:, perm[real]]
    assert np.abs(eh["residuals"] - eo["residuals"]).max() <= 1e-10 * max(1.0, np.abs(eo["residuals"]).max())
    assert np.abs(eh["J"][:, real] - Jo).max() <= 1e-9 * max(1.0, np.abs(Jo).max())
    # the half bandwidth equals the brute-force one: widest span of band columns (< nb) over the rows of one residual block, and -- after
    # the inverse depths are eliminated -- over all rows of one landmark
    lay = eh["layout"]
    J, nb = eh["J"], lay["nb"]
    n_g, n_a, n_s, n_c = (len(pd.tables[k][0]) for k in ("gyro", "accel", "surfel", "cam"))
    spans = []
    def span(rows):
        cols = np.nonzero(np.abs(J[rows, :nb]).sum(axis=0))[0]
        return (cols.max() - cols.min()) if len(cols) else 0
    r0 = 0
    for cnt, rows_per in ((n_g, 3), (n_a, 3), (n_s, 1)):
        for b in range(0, cnt, max(1, cnt // 200)):        # a sample of the blocks (they are time-ordered and alike)
            spans.append(span(slice(r0 + rows_per * b, r0 + rows_per * (b + 1))))
        r0 += rows_per * cnt
    lm = np.asarray(pd.tables["cam"][4])   # (t0_ref, t0_obs, uv_ref, uv_obs, landmark, weight, huber)
    for l in np.unique(lm):
        idx = np.nonzero(lm == l)[0]
        rows = np.concatenate([[r0 + 2 * i, r0 + 2 * i + 1] for i in idx])
        spans.append(span(rows))
    assert lay["bw"] == max(spans)
