"""ctypes binding of lvi_exc_b200/lib/liblvi_hostcheck.so: the product's lowering + analytic Jacobian headers compiled for the
host (lvi_exc_b200/csrc/hostcheck.cpp).  Used only by the CPU tests."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from lvi_exc_b200._capi import ProblemDesc, c_double_p, c_int32_p, ptr

_PATH = Path(__file__).resolve().parent.parent / "lvi_exc_b200" / "lib" / "liblvi_hostcheck.so"
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(str(_PATH))
        L.lvi_hostcheck_last_error.restype = C.c_char_p
        L.lvi_hostcheck_layout.argtypes = [C.POINTER(ProblemDesc), c_int32_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p]
        L.lvi_hostcheck_evaluate.argtypes = [C.POINTER(ProblemDesc), c_double_p, c_double_p, c_double_p, c_double_p]
        L.lvi_hostcheck_traj_eval.argtypes = [C.POINTER(ProblemDesc), C.c_double, c_double_p]
        _lib = L
    return _lib


def layout(pd) -> dict:
    d = pd.desc()
    out = np.zeros(8, np.int32)
    n, nl = pd.n_knots, max(len(pd.rho), 1)
    pr3, pso3, psens, prho = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(7, np.int32), np.zeros(nl, np.int32)
    rc = lib().lvi_hostcheck_layout(C.byref(d), ptr(out), ptr(pr3), ptr(pso3), ptr(psens), ptr(prho))
    if rc:
        raise (IndexError if rc == -4 else RuntimeError)(lib().lvi_hostcheck_last_error().decode())
    return dict(n_res=int(out[0]), nt=int(out[1]), nb=int(out[2]), nbo=int(out[3]), bw=int(out[4]), chain1_start=int(out[5]), n_mid=int(out[6]), n_pad=int(out[7]), pos_r3=pr3, pos_so3=pso3, pos_sens=psens,
                pos_rho=prho[:len(pd.rho)])


def evaluate(pd, jacobian=True) -> dict:
    lay = layout(pd)
    d = pd.desc()
    cost, fixed = np.zeros(1), np.zeros(1)
    res = np.zeros(lay["n_res"])
    J = np.zeros((lay["n_res"], lay["nt"])) if jacobian else None
    rc = lib().lvi_hostcheck_evaluate(C.byref(d), ptr(cost), ptr(fixed), ptr(res), ptr(J))
    if rc:
        raise (IndexError if rc == -4 else RuntimeError)(lib().lvi_hostcheck_last_error().decode())
    return dict(cost=float(cost[0]), fixed_cost=float(fixed[0]), residuals=res, J=J, layout=lay)


def traj_eval(pd, t: float) -> dict:
    out = np.zeros(13)
    d = pd.desc()
    rc = lib().lvi_hostcheck_traj_eval(C.byref(d), t, ptr(out))
    if rc:
        raise IndexError("t out of range")
    return dict(p=out[0:3], a=out[3:6], q=out[6:10], w_body=out[10:13])


def perm_to_oracle(pd, lay, op) -> np.ndarray:
    """perm[library tangent position] = oracle tangent offset (-1 at the unused padding positions of the two-sided ordering)"""
    perm = np.full(lay["nt"], -1, np.int64)
    for i in range(pd.n_knots):
        if lay["pos_r3"][i] >= 0:
            o = op.offset_knot(i, False)
            perm[lay["pos_r3"][i]:lay["pos_r3"][i] + 3] = np.arange(o, o + 3)
        if lay["pos_so3"][i] >= 0:
            o = op.offset_knot(i, True)
            perm[lay["pos_so3"][i]:lay["pos_so3"][i] + 3] = np.arange(o, o + 3)
    for which, dm in enumerate([3, 3, 3, 3, 2, 3, 3]):
        p, o = lay["pos_sens"][which], op.offset_block(which)
        assert (p < 0) == (o < 0), (which, p, o)
        if p >= 0:
            perm[p:p + dm] = np.arange(o, o + dm)
    for l, p in enumerate(lay["pos_rho"]):
        o = op.offset_block(7 + l)
        assert (p < 0) == (o < 0)
        if p >= 0:
            perm[p] = o
    pad = np.arange(lay["chain1_start"] - lay["n_pad"], lay["chain1_start"])
    real = np.setdiff1d(np.arange(lay["nt"]), pad)
    assert (perm[pad] < 0).all() and (perm[real] >= 0).all() and len(set(perm[real].tolist())) == len(real)
    return perm
