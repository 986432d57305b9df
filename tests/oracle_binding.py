"""ctypes binding of the CPU oracle (oracle/liblvi_oracle.so).  TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from lvi_exc_b200._capi import (ProblemDesc, SolveOptions, SolveSummary, SURFEL_POINT_DTYPE, c_double_p, c_int32_p,
                                c_int64_p, ptr)

_PATH = Path(__file__).resolve().parent.parent / "oracle" / "liblvi_oracle.so"
_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not _PATH.exists():
            raise RuntimeError(f"{_PATH} missing: run __graft_entry__.build() (or make -C oracle)")
        L = C.CDLL(str(_PATH))
        vp = C.c_void_p
        L.orc_voxel_build.restype = vp
        L.orc_voxel_build.argtypes = [vp, C.c_int64, C.c_int64, C.c_float, C.c_int, C.c_double]
        L.orc_voxel_status.argtypes = [vp]
        L.orc_voxel_free.argtypes = [vp]
        L.orc_voxel_num_leaves.argtypes = [vp]
        L.orc_voxel_num_leaves.restype = C.c_int64
        L.orc_voxel_num_points.argtypes = [vp]
        L.orc_voxel_num_points.restype = C.c_int64
        L.orc_voxel_grid.argtypes = [vp, c_int32_p, c_int32_p]
        L.orc_voxel_export.argtypes = [vp, c_int64_p, c_int32_p] + [c_double_p] * 5 + [c_int64_p, c_int32_p]
        L.orc_surfel_extract.restype = vp
        L.orc_surfel_extract.argtypes = [vp, C.c_double, C.c_int, C.c_float, C.c_int]
        L.orc_surfel_free.argtypes = [vp]
        L.orc_surfel_count.argtypes = [vp]
        L.orc_surfel_count.restype = C.c_int64
        L.orc_surfel_export.argtypes = [vp] + [c_double_p] * 4 + [c_int64_p, c_int32_p]
        L.orc_associate.restype = C.c_int64
        L.orc_associate.argtypes = [vp, vp, vp, C.c_int64, vp, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_int32,
                                    C.c_int32, vp, C.c_int64, c_int64_p]
        L.orc_problem_create.restype = vp
        L.orc_problem_create.argtypes = [C.POINTER(ProblemDesc), c_int32_p]
        L.orc_problem_free.argtypes = [vp]
        L.orc_problem_num_residuals.argtypes = [vp]
        L.orc_problem_num_tangent.argtypes = [vp]
        L.orc_problem_tangent_offset_knot.argtypes = [vp, C.c_int, C.c_int]
        L.orc_problem_tangent_offset_block.argtypes = [vp, C.c_int]
        L.orc_problem_evaluate.argtypes = [vp] + [c_double_p] * 5
        L.orc_problem_solve.argtypes = [vp, C.POINTER(SolveOptions), C.POINTER(SolveSummary)]
        L.orc_traj_eval.argtypes = [C.POINTER(ProblemDesc), C.c_double, c_double_p]
        L.orc_last_error.restype = C.c_char_p
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_set_num_threads.restype = None
        L.orc_undistort.argtypes = [C.POINTER(ProblemDesc), vp, C.c_int32, C.c_int64, c_double_p, C.c_int, vp]
        L.orc_associate_landmarks.argtypes = [vp, c_double_p, C.c_int64, C.c_double, c_int32_p]
        L.orc_associate_landmarks.restype = None
        _lib = L
    return _lib


def set_num_threads(n: int) -> int:
    """OpenMP threads of the oracle (torchrun exports OMP_NUM_THREADS=1; timed runs set the count explicitly)"""
    lib().orc_set_num_threads(int(n))
    return lib().orc_num_threads()


class OracleVoxelMap:
    """pclomp::VoxelGridCovariance restated on the CPU"""

    def __init__(self, cloud: np.ndarray, leaf: float = 0.5, min_points: int = 6, eig_mult: float = 0.01):
        cloud = np.ascontiguousarray(cloud, dtype=np.float32)
        assert cloud.ndim == 2 and cloud.shape[1] >= 3
        self.cloud = cloud  # keep alive: the oracle reads points in place
        self.h = lib().orc_voxel_build(cloud.ctypes.data, cloud.shape[1], cloud.shape[0], leaf, min_points, eig_mult)
        self.status = lib().orc_voxel_status(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_voxel_free(self.h)
            self.h = None

    @property
    def num_leaves(self) -> int:
        return lib().orc_voxel_num_leaves(self.h)

    def grid(self):
        mn, dv = np.zeros(3, np.int32), np.zeros(3, np.int32)
        lib().orc_voxel_grid(self.h, ptr(mn), ptr(dv))
        return mn, dv

    def export(self) -> dict:
        L = self.num_leaves
        npts = lib().orc_voxel_num_points(self.h)
        out = dict(keys=np.zeros(L, np.int64), nr_points=np.zeros(L, np.int32), mean=np.zeros((L, 3)), cov=np.zeros((L, 9)),
                   evals=np.zeros((L, 3)), evecs=np.zeros((L, 9)), icov=np.zeros((L, 9)), leaf_start=np.zeros(L + 1, np.int64),
                   point_index=np.zeros(npts, np.int32))
        lib().orc_voxel_export(self.h, ptr(out["keys"]), ptr(out["nr_points"]), ptr(out["mean"]), ptr(out["cov"]), ptr(out["evals"]),
                               ptr(out["evecs"]), ptr(out["icov"]), ptr(out["leaf_start"]), ptr(out["point_index"]))
        return out


class OracleSurfels:
    def __init__(self, vmap: OracleVoxelMap, lam: float = 0.6, min_leaf_points: int = 10, thr: float = 0.05, min_inliers: int = 20):
        self.vmap = vmap
        self.h = lib().orc_surfel_extract(vmap.h, lam, min_leaf_points, thr, min_inliers)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_surfel_free(self.h)
            self.h = None

    @property
    def count(self) -> int:
        return lib().orc_surfel_count(self.h)

    def export(self) -> dict:
        P = self.count
        out = dict(p4=np.zeros((P, 4)), Pi=np.zeros((P, 3)), box_min=np.zeros((P, 3)), box_max=np.zeros((P, 3)),
                   leaf_key=np.zeros(P, np.int64), n_inliers=np.zeros(P, np.int32))
        lib().orc_surfel_export(self.h, ptr(out["p4"]), ptr(out["Pi"]), ptr(out["box_min"]), ptr(out["box_max"]), ptr(out["leaf_key"]),
                                ptr(out["n_inliers"]))
        return out

    def associate(self, scans_map: np.ndarray, scans_raw: np.ndarray, radius=0.05, k_per_ring=2, time_step=10, mode=0):
        """scans_map [S,H,W,C>=3] float32, scans_raw [S,H,W] RAW_POINT_DTYPE -> (downsampled SurfelPoint array, n_all)"""
        S, H, W = scans_raw.shape
        scans_map = np.ascontiguousarray(scans_map, dtype=np.float32)
        scans_raw = np.ascontiguousarray(scans_raw)
        n_all = np.zeros(1, np.int64)
        args = (self.vmap.h, self.h, scans_map.ctypes.data, scans_map.shape[-1], scans_raw.ctypes.data, S, W, H, radius, k_per_ring,
                time_step, mode)
        n = lib().orc_associate(*args, None, 0, ptr(n_all))
        out = np.zeros(n, dtype=SURFEL_POINT_DTYPE)
        lib().orc_associate(*args, out.ctypes.data, n, ptr(n_all))
        return out, int(n_all[0])


class OracleProblem:
    def __init__(self, data):
        self.data = data
        self.desc = data.desc()
        st = np.zeros(1, np.int32)
        self.h = lib().orc_problem_create(C.byref(self.desc), ptr(st))
        if not self.h:
            msg = lib().orc_last_error().decode()
            raise (IndexError(msg) if st[0] == -4 else RuntimeError(msg))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_problem_free(self.h)
            self.h = None

    @property
    def num_residuals(self):
        return lib().orc_problem_num_residuals(self.h)

    @property
    def num_tangent(self):
        return lib().orc_problem_num_tangent(self.h)

    def offset_knot(self, i, so3: bool):
        return lib().orc_problem_tangent_offset_knot(self.h, i, int(so3))

    def offset_block(self, which):
        return lib().orc_problem_tangent_offset_block(self.h, which)

    def evaluate(self, jacobian=False, gradient=True):
        cost, fixed = np.zeros(1), np.zeros(1)
        res = np.zeros(self.num_residuals)
        g = np.zeros(self.num_tangent) if gradient else None
        J = np.zeros((self.num_residuals, self.num_tangent)) if jacobian else None
        rc = lib().orc_problem_evaluate(self.h, ptr(cost), ptr(fixed), ptr(res), ptr(g), ptr(J))
        if rc:
            raise (IndexError if rc == -4 else RuntimeError)(lib().orc_last_error().decode())
        return dict(cost=float(cost[0]), fixed_cost=float(fixed[0]), residuals=res, gradient=g, J=J)

    def solve(self, max_iterations=30, verbose=False, **kw) -> SolveSummary:
        opt = SolveOptions.default(max_iterations, verbose)
        for k, v in kw.items():
            setattr(opt, k, v)
        s = SolveSummary()
        rc = lib().orc_problem_solve(self.h, C.byref(opt), C.byref(s))
        if rc:
            raise (IndexError if rc == -4 else RuntimeError)(lib().orc_last_error().decode())
        return s


def traj_eval(data, t: float) -> dict:
    out = np.zeros(16)
    d = data.desc()
    rc = lib().orc_traj_eval(C.byref(d), t, ptr(out))
    if rc:
        raise IndexError(lib().orc_last_error().decode())
    return dict(p=out[0:3], v=out[3:6], a=out[6:9], q=out[9:13], w=out[13:16])
