"""The CPU oracle's LM iteration log against tests/lm_twin.py, an independent dense restatement of SURVEY Appendix C: same accept / reject
sequence, costs and trust-region radii on problems with quaternion blocks (S0), Huber-active residuals (S1 with outliers) and the camera
tables with free inverse depths (S4, bounds inactive)."""
import numpy as np
import pytest

from tests import lm_twin
from tests import oracle_binding as ob
from tests.problems import make_lvi_problem


def _logs(pd, iters):
    saved = pd.clone_params()
    s = ob.OracleProblem(pd).solve(iters)
    olog = [(s.log_cost[k] - s.fixed_cost, s.log_radius[k], bool(s.log_successful[k])) for k in range(s.n_log)]
    pd.restore_params(saved)
    tlog = lm_twin.solve(pd, ob.OracleProblem, max_iterations=iters)
    pd.restore_params(saved)
    return olog, tlog


def _compare(olog, tlog, rel):
    assert len(olog) == len(tlog), (len(olog), len(tlog))
    assert [o[2] for o in olog] == [t[2] for t in tlog]                       # accept / reject sequence
    for (oc, orad, _), (tc, trad, _) in zip(olog, tlog):
        assert oc == pytest.approx(tc, rel=rel, abs=1e-12)
        assert orad == pytest.approx(trad, rel=1e-6)


def test_twin_so3_quaternion_blocks():
    pd = make_lvi_problem("so3", 1.0, 0)
    olog, tlog = _logs(pd, 12)
    assert len(olog) >= 4
    _compare(olog, tlog, 1e-7)


def test_twin_surfel_with_huber_active_residuals():
    pd = make_lvi_problem("surfel", 1.0, 300)
    # every 7th surfel point is pushed ~1 m away: weight 10 x 1 m = 10 > Huber 5 -> the corrector is active on those rows
    n_imu_rows = 3 * (len(pd.tables["gyro"][0]) + len(pd.tables["accel"][0]))
    big0 = int(np.sum(np.abs(ob.OracleProblem(pd).evaluate(gradient=False)["residuals"][n_imu_rows:]) > 5.0))
    pd.tables["surfel"][2][::7] += np.array([0.6, 0.6, 0.6])
    big1 = int(np.sum(np.abs(ob.OracleProblem(pd).evaluate(gradient=False)["residuals"][n_imu_rows:]) > 5.0))
    assert big1 > big0 + 50                                                    # the outliers are in, beyond the Huber threshold
    olog, tlog = _logs(pd, 8)
    assert sum(1 for o in olog if not o[2]) >= 0 and len(olog) >= 4
    _compare(olog, tlog, 1e-7)


def test_twin_lvi_camera_tables():
    pd = make_lvi_problem("lvi")
    olog, tlog = _logs(pd, 6)
    assert len(olog) >= 3
    _compare(olog, tlog, 1e-6)
