"""`backend` protocol of lvi_exc_b200.pipeline.run_calibration implemented with the CPU oracle.
TEST INFRASTRUCTURE ONLY — the product backend is lvi_exc_b200.backend.CudaBackend."""
from __future__ import annotations

import ctypes as C

import numpy as np

from lvi_exc_b200 import synth
from lvi_exc_b200._capi import ptr
from tests import oracle_binding as ob


class OracleSurfelMap:
    def __init__(self, cloud, leaf, lam):
        self.vmap = ob.OracleVoxelMap(cloud, leaf)
        self.surfels = ob.OracleSurfels(self.vmap, lam)
        self.planes = self.surfels.export()
        self.planes_Pi = self.planes["Pi"]


class OracleBackend:
    name = "oracle"

    def __init__(self, assoc_mode: int = 1, verbose=False):
        self.assoc_mode = assoc_mode  # 0 = reference-faithful O(P*W*H) sweep, 1 = voxel lookup
        self.verbose = verbose

    def solve(self, pd, max_iterations):
        return ob.OracleProblem(pd).solve(max_iterations, verbose=self.verbose)

    def build_surfel_map(self, cloud, leaf, lam):
        return OracleSurfelMap(cloud, leaf, lam)

    def map_cloud(self, scans_in_map, keys=None):
        a = scans_in_map if keys is None else scans_in_map[np.nonzero(keys)[0]]
        return a.reshape(-1, a.shape[-1])

    def associate(self, smap, scans_in_map, scans_raw, radius, k, step):
        pts, _ = smap.surfels.associate(scans_in_map, scans_raw, radius, k, step, self.assoc_mode)
        return pts

    def transform(self, scans_xyzi, poses):
        S = scans_xyzi.shape[0]
        out = np.zeros_like(scans_xyzi)
        for s in range(S):
            T = poses[s].astype(np.float32)
            x, y, z = scans_xyzi[s, ..., 0], scans_xyzi[s, ..., 1], scans_xyzi[s, ..., 2]
            for r in range(3):
                out[s, ..., r] = T[r, 0] * x + T[r, 1] * y + T[r, 2] * z + T[r, 3]
            out[s, ..., 3] = 1.0
            out[s, ..., 4] = scans_xyzi[s, ..., 4]
        return out

    def undistort(self, pd, scans_raw, target_time, correct_position):
        S, H, W = scans_raw.shape
        if target_time is None:   # undistortScan(): each scan expressed at its own stamp = first point's firing time
            tt = np.ascontiguousarray(scans_raw[:, 0, 0]["timestamp"], dtype=np.float64)
        else:
            tt = np.full(S, float(target_time))
        out = np.zeros((S, H, W, 8), dtype=np.float32)
        d = pd.desc()
        raw = np.ascontiguousarray(scans_raw)
        ob.lib().orc_undistort(C.byref(d), raw.ctypes.data, S, H * W, ptr(tt), int(correct_position), out.ctypes.data)
        return out

    def traj_eval(self, pd, t):
        return ob.traj_eval(pd, t)

    def traj_eval_many(self, pd, times):
        n = len(times)
        pos, quat, valid = np.zeros((n, 3)), np.zeros((n, 4)), np.zeros(n, bool)
        quat[:, 3] = 1.0
        for i, t in enumerate(times):
            try:
                e = ob.traj_eval(pd, float(t))
            except IndexError:
                continue
            pos[i], quat[i], valid[i] = e["p"], e["q"], True
        return pos, quat, valid

    def associate_landmarks(self, smap, pts, radius):
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        out = np.zeros(len(pts), np.int32)
        ob.lib().orc_associate_landmarks(smap.surfels.h, ptr(pts), len(pts), radius, ptr(out))
        return out
