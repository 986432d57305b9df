"""A second, independently written restatement of the Ceres trust-region loop, in dense numpy.

The CPU oracle (oracle/orc_solve.cpp) restates TrustRegionMinimizer + LevenbergMarquardtStrategy in C++ over the band+arrow structure;
this twin restates the same Appendix C steps from scratch on DENSE matrices (numpy Cholesky), taking from the oracle only the evaluation of
the loss-corrected residual vector and tangent Jacobian.  Agreement of the two iteration logs (cost, trust-region radius, accept / reject,
iteration count) shrinks the unpinned surface to: the evaluation itself (checked against finite differences elsewhere) and the reading of
Appendix C that both authors share.  Places where an interpretation had to be fixed are marked (I1)...(I5) and listed in DESIGN.md §7."""
from __future__ import annotations

import numpy as np


def _quat_mul(a, b):
    ax, ay, az, aw = a; bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz])


def _quat_plus(q, d):
    n = np.linalg.norm(d)   # EigenQuaternionParameterization: q+ = [sin|d|/|d| d, cos|d|] * q, d the HALF-angle vector (C.3)
    if n == 0.0:
        return q.copy()
    return _quat_mul(np.array([*(np.sin(n) / n * d), np.cos(n)]), q)


class Blocks:
    """free parameter blocks of a ProblemData and where they sit in the oracle's tangent vector"""

    def __init__(self, pd, op):
        self.pd, self.items = pd, []
        for i in range(pd.n_knots):
            if pd.r3_knots is not None:
                o = op.offset_knot(i, False)
                if o >= 0: self.items.append(("r3", i, o, 3))
            o = op.offset_knot(i, True)
            if o >= 0: self.items.append(("so3", i, o, 3))
        for which, (name, dim) in enumerate([("lidar_q", 3), ("lidar_p", 3), ("cam_q", 3), ("cam_p", 3), ("gravity", 2), ("acc_bias", 3), ("gyr_bias", 3)]):
            o = op.offset_block(which)
            if o >= 0: self.items.append((name, None, o, dim))
        for l in range(len(pd.rho)):
            o = op.offset_block(7 + l)
            if o >= 0: self.items.append(("rho", l, o, 1))

    def snapshot(self):
        pd = self.pd
        return {k: (None if getattr(pd, k) is None else getattr(pd, k).copy()) for k in ("r3_knots", "so3_knots", "lidar_q", "lidar_p", "cam_q", "cam_p", "gravity", "acc_bias", "gyr_bias", "rho")}

    def restore(self, s):
        for k, v in s.items():
            if v is not None:
                getattr(self.pd, k)[...] = v

    def plus(self, base, delta):
        """pd <- Plus(base, delta)"""
        self.restore(base)
        pd = self.pd
        for name, idx, o, dim in self.items:
            d = delta[o:o + dim]
            if name == "r3": pd.r3_knots[idx] = base["r3_knots"][idx] + d
            elif name == "so3": pd.so3_knots[idx] = _quat_plus(base["so3_knots"][idx], d)
            elif name in ("lidar_q", "cam_q"): getattr(pd, name)[...] = _quat_plus(base[name], d)
            elif name == "rho": pd.rho[idx] = base["rho"][idx] + d[0]
            else: getattr(pd, name)[...] = base[name] + d

    def ambient(self):
        """the free parameters in AMBIENT coordinates (quaternion blocks with their 4 coefficients): Ceres measures ||x|| and the step norm there (I4)"""
        pd, out = self.pd, []
        for name, idx, o, dim in self.items:
            v = pd.r3_knots[idx] if name == "r3" else pd.so3_knots[idx] if name == "so3" else pd.rho[idx:idx + 1] if name == "rho" else getattr(pd, name)
            out.append(np.asarray(v, dtype=np.float64).ravel().copy())
        return np.concatenate(out) if out else np.zeros(0)


def solve(pd, make_problem, max_iterations=30, radius0=1e4, min_relative_decrease=1e-3, function_tolerance=1e-6, gradient_tolerance=1e-10,
          parameter_tolerance=1e-8, min_lm_diagonal=1e-6, max_lm_diagonal=1e32, max_radius=1e16, min_radius=1e-32):
    """unconstrained problems only (no active bound).  make_problem(pd) -> object with evaluate(jacobian=True) returning cost (incl. loss),
    loss-corrected residuals and the tangent Jacobian.  Returns the iteration log [(cost, radius, successful)], cost excludes the fixed cost."""
    op = make_problem(pd)
    B = Blocks(pd, op)
    ev = op.evaluate(jacobian=True)
    J, r, cost = ev["J"], ev["residuals"], ev["cost"]
    scale = 1.0 / (1.0 + np.sqrt(np.sum(J * J, axis=0)))          # Jacobi scaling, ONCE, at iteration 0 (C.4)
    radius, decrease = radius0, 2.0
    log = [(cost, radius, True)]
    g = J.T @ r
    if np.max(np.abs(g)) <= gradient_tolerance:
        return log
    reuse_diag, diag = False, None
    for _ in range(max_iterations):
        Js = J * scale
        H = Js.T @ Js
        gs = Js.T @ r
        if not reuse_diag:                                         # (I1) the LM diagonal is refreshed after a successful step only
            diag = np.clip(np.diag(H), min_lm_diagonal, max_lm_diagonal)
        A = H + np.diag(diag / radius)
        y = np.linalg.solve(A, -gs)
        delta = scale * y
        Jd = J @ delta
        model_change = -float(Jd @ (r + 0.5 * Jd))                 # C.6
        if not (model_change > 0):
            radius /= decrease; decrease *= 2; reuse_diag = True   # invalid step: treated like a rejected one (I2)
            log.append((cost, radius, False))
            continue
        base = B.snapshot()
        x0 = B.ambient()
        B.plus(base, delta)
        ev2 = make_problem(pd).evaluate(jacobian=True)
        # (I3) TrustRegionMinimizer::Minimize tests the parameter and function tolerances on the CANDIDATE, before deciding whether the step
        # is accepted; when one of them fires the candidate is not taken and the iteration is logged as unsuccessful
        step_norm = float(np.linalg.norm(B.ambient() - x0))
        if step_norm <= parameter_tolerance * (float(np.linalg.norm(x0)) + parameter_tolerance):
            B.restore(base)
            log.append((cost, radius, False))
            break
        if abs(cost - ev2["cost"]) <= function_tolerance * cost:          # (I5) relative to the cost at the CURRENT point
            B.restore(base)
            log.append((cost, radius, False))
            break
        rho_step = (cost - ev2["cost"]) / model_change
        if rho_step > min_relative_decrease:
            J, r, cost = ev2["J"], ev2["residuals"], ev2["cost"]
            radius = min(max_radius, radius / max(1.0 / 3.0, 1.0 - (2.0 * rho_step - 1.0) ** 3))
            decrease, reuse_diag = 2.0, False
            log.append((cost, radius, True))
            g = J.T @ r
            if np.max(np.abs(g)) <= gradient_tolerance:
                break
        else:
            B.restore(base)
            radius /= decrease; decrease *= 2; reuse_diag = True
            log.append((cost, radius, False))
            if radius < min_radius:
                break
    return log
