"""CUDA backend of the calibration pipeline: every heavy step goes through the C-ABI of liblvi_exc_b200.so
(include/lvi_exc_b200.h).  There is no CPU fallback: constructing the backend without the library or without a CUDA
device raises.

Mirrors, on the host side, the objects the reference driver holds (L/test/lvi_initialize_surfel_orb.cpp:227-241):
  CudaSurfelMap  = LiDAROdometry's pclomp::NormalDistributionsTransform target cells + SurfelAssociation::surfel_planes_
  CudaBackend    = the calls LIinitializer::DataAssociation / Mapping / trajInitFrom* make (T:1169-1300)

Point clouds stay resident in HBM between steps (torch tensors are used only as device-memory holders; all compute
is in the library's own kernels on the library's own stream).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import SURFEL_POINT_DTYPE, SolveOptions, SolveSummary, check, ptr


def _torch():
    import torch
    return torch


class CudaScanBatch:
    """lvi_scan_batch: organised scans in one frame, packed in HBM (ScanUndistortion::scan_data_in_map_ on the device)."""

    def __init__(self, backend: "CudaBackend", handle, S: int, H: int, W: int):
        self.b, self.h, self.shape = backend, handle, (S, H, W)

    def numpy(self) -> np.ndarray:
        """the scans as pcl::PointXYZI records: [S, H, W, 8] float32 (x, y, z, 1, intensity, 0, 0, 0)"""
        out = np.zeros(self.shape + (8,), dtype=np.float32)
        check(self.b.lib.lvi_scan_batch_export_xyzi(self.b.ctx, self.h, C.c_void_p(out.ctypes.data), 0))
        return out

    def cpu(self):
        return self

    def close(self):
        if self.h:
            self.b.lib.lvi_scan_batch_destroy(self.h); self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CudaMapCloud:
    """the cloud a map is built from: the key scans of a batch (LiDAROdometry::updateKeyScan, L/src/core/lidar_odometry.cpp:89-104)"""

    def __init__(self, batch: CudaScanBatch, keep: np.ndarray | None):
        self.batch = batch
        self.keep = None if keep is None else np.ascontiguousarray(keep, dtype=np.uint8)


class CudaSurfelMap:
    def __init__(self, backend: "CudaBackend", cloud, leaf: float, lam: float, min_points=6, eig_mult=0.01, min_leaf_points=10,
                 ransac_thr=0.05, min_inliers=20, sharded=False):
        self.b = backend
        lib = backend.lib
        self.vmap = C.c_void_p()
        self.surfels = C.c_void_p()
        self.shard_stats = None
        self._planes = None
        self._Pi = None
        if isinstance(cloud, CudaScanBatch):
            cloud = CudaMapCloud(cloud, None)
        if sharded:   # every rank passes the scans of its own time chunk: grid, leaves and planes are those of the whole cloud
            assert isinstance(cloud, CudaMapCloud)
            st = np.zeros(4, np.int64)
            check(lib.lvi_map_build_sharded(backend.ctx, cloud.batch.h, ptr(cloud.keep), leaf, min_points, eig_mult, lam, min_leaf_points, ransac_thr,
                                            min_inliers, C.byref(self.vmap), C.byref(self.surfels), ptr(st)))
            self.shard_stats = dict(points_sent=int(st[0]), points_received=int(st[1]), leaves_built=int(st[2]), planes_built=int(st[3]))
            self.num_leaves = lib.lvi_voxel_num_leaves(self.vmap)
            self.num_planes = lib.lvi_surfel_count(self.surfels)
            return
        if isinstance(cloud, CudaMapCloud):
            check(lib.lvi_voxel_build_batch(backend.ctx, cloud.batch.h, ptr(cloud.keep), leaf, min_points, eig_mult, C.byref(self.vmap)))
        else:   # a PCL-shaped cloud (numpy / device tensor): the ABI's 32 B layout
            dev, n, stride, keep = backend._as_device_cloud(cloud)
            self._keep = keep
            check(lib.lvi_voxel_build_d(backend.ctx, dev, stride, n, leaf, min_points, eig_mult, C.byref(self.vmap)))
        check(lib.lvi_surfel_extract(backend.ctx, self.vmap, lam, min_leaf_points, ransac_thr, min_inliers, C.byref(self.surfels)))
        self._keep = None  # the map keeps its own sorted copy of the points
        self.num_leaves = lib.lvi_voxel_num_leaves(self.vmap)
        self.num_planes = lib.lvi_surfel_count(self.surfels)

    # The plane tables stay on the device (association reads them there); the host copies are made when somebody asks: `planes_Pi` (the
    # closest points the residual tables carry) downloads 24 B per plane, `planes` everything (116 B per plane).
    @property
    def planes_Pi(self) -> np.ndarray:
        if self._Pi is None:
            if self._planes is not None:
                self._Pi = self._planes["Pi"]
            else:
                self._Pi = np.zeros((self.num_planes, 3))
                check(self.b.lib.lvi_surfel_export(self.b.ctx, self.surfels, None, ptr(self._Pi), None, None, None, None))
        return self._Pi

    @property
    def planes(self) -> dict:
        if self._planes is None:
            self._planes = self.export_planes()
        return self._planes

    def export_planes(self) -> dict:
        P = self.num_planes
        out = dict(p4=np.zeros((P, 4)), Pi=np.zeros((P, 3)), box_min=np.zeros((P, 3)), box_max=np.zeros((P, 3)),
                   leaf_key=np.zeros(P, np.int64), n_inliers=np.zeros(P, np.int32))
        check(self.b.lib.lvi_surfel_export(self.b.ctx, self.surfels, ptr(out["p4"]), ptr(out["Pi"]), ptr(out["box_min"]),
                                           ptr(out["box_max"]), ptr(out["leaf_key"]), ptr(out["n_inliers"])))
        return out

    def export_leaves(self) -> dict:
        L = self.num_leaves
        npts = self.b.lib.lvi_voxel_num_points(self.vmap)
        out = dict(keys=np.zeros(L, np.int64), nr_points=np.zeros(L, np.int32), mean=np.zeros((L, 3)), cov=np.zeros((L, 9)),
                   evals=np.zeros((L, 3)), evecs=np.zeros((L, 9)), icov=np.zeros((L, 9)), leaf_start=np.zeros(L + 1, np.int64),
                   point_index=np.zeros(npts, np.int32))
        check(self.b.lib.lvi_voxel_export(self.b.ctx, self.vmap, ptr(out["keys"]), ptr(out["nr_points"]), ptr(out["mean"]), ptr(out["cov"]),
                                          ptr(out["evals"]), ptr(out["evecs"]), ptr(out["icov"]), ptr(out["leaf_start"]), ptr(out["point_index"])))
        return out

    def grid(self):
        mn, dv = np.zeros(3, np.int32), np.zeros(3, np.int32)
        check(self.b.lib.lvi_voxel_grid(self.vmap, ptr(mn), ptr(dv)))
        return mn, dv

    def close(self):
        if self.surfels:
            self.b.lib.lvi_surfel_destroy(self.surfels); self.surfels = C.c_void_p()
        if self.vmap:
            self.b.lib.lvi_voxel_destroy(self.vmap); self.vmap = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CudaProblem:
    """kontiki::TrajectoryEstimator + ceres::Problem on the device (lvi_problem)."""

    def __init__(self, backend: "CudaBackend", pd):
        self.b, self.pd = backend, pd
        self.desc = pd.desc()
        self.h = C.c_void_p()
        check(backend.lib.lvi_problem_create(backend.ctx, C.byref(self.desc), C.byref(self.h)))

    @property
    def num_residuals(self):
        return self.b.lib.lvi_problem_num_residuals(self.h)

    @property
    def num_tangent(self):
        return self.b.lib.lvi_problem_num_tangent(self.h)

    def evaluate(self, jacobian=False, gradient=True, residuals=True):
        cost = np.zeros(1)
        res = np.zeros(self.num_residuals) if residuals else None
        g = np.zeros(self.num_tangent) if gradient else None
        check(self.b.lib.lvi_problem_evaluate(self.h, ptr(cost), ptr(res), ptr(g)))
        J = None
        if jacobian:
            J = np.zeros((self.num_residuals, self.num_tangent))
            check(self.b.lib.lvi_problem_jacobian_dense(self.h, ptr(J)))
        return dict(cost=float(cost[0]), residuals=res, gradient=g, J=J)

    def solve(self, max_iterations=30, verbose=False, **kw) -> SolveSummary:
        opt = SolveOptions.default(max_iterations, verbose)
        for k, v in kw.items():
            setattr(opt, k, v)
        s = SolveSummary()
        check(self.b.lib.lvi_problem_solve(self.h, C.byref(opt), C.byref(s)))
        return s

    def bench_iterations(self, iters: int) -> np.ndarray:
        ms = (C.c_float * 6)()
        check(self.b.lib.lvi_problem_bench_iterations(self.h, iters, ms))
        return np.array(list(ms), dtype=np.float64)

    def close(self):
        if self.h:
            self.b.lib.lvi_problem_destroy(self.h); self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CudaBackend:
    name = "cuda"

    def __init__(self, device: int = 0, verbose: bool = False, nccl_id: bytes | None = None, rank: int = 0, world: int = 1):
        self.lib = _capi.load()          # raises LibraryMissing when the CUDA library has not been built
        self.verbose = verbose
        self.device = device
        self.ctx = C.c_void_p()
        if world > 1:
            assert nccl_id is not None and len(nccl_id) == 128
            buf = C.create_string_buffer(nccl_id, 128)
            check(self.lib.lvi_ctx_create_nccl(device, buf, rank, world, C.byref(self.ctx)))
        else:
            check(self.lib.lvi_ctx_create(device, None, 0, 1, C.byref(self.ctx)))  # LVI_ERR_NO_DEVICE without a GPU
        self.rank, self.world = rank, world
        self._dev_cache = {}

    def close(self):
        if self.ctx:
            self.lib.lvi_ctx_destroy(self.ctx); self.ctx = C.c_void_p()

    @property
    def launches(self) -> int:
        return int(self.lib.lvi_ctx_launch_count(self.ctx))

    def synchronize(self):
        check(self.lib.lvi_ctx_synchronize(self.ctx))

    def kernel_timing(self, enable: bool):
        check(self.lib.lvi_ctx_kernel_timing(self.ctx, int(enable)))

    def kernel_times(self) -> dict:
        """{kernel name: (launches, total ms)} of the launches since timing was enabled / last drained (CUDA events on the launch stream)"""
        buf = C.create_string_buffer(1 << 16)
        self.lib.lvi_ctx_kernel_times(self.ctx, buf, len(buf))
        out = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.rsplit(" ", 2)
            out[name] = (int(cnt), float(ms))
        return out

    def _download_records(self, dev_bytes, n: int, dtype) -> np.ndarray:
        """n records of `dtype` from a device byte tensor through a pinned staging buffer that is kept (and grown) across calls; the
        returned array is a view of that buffer, valid until the next download"""
        torch = _torch()
        nbytes = n * dtype.itemsize
        if nbytes == 0:
            return np.zeros(0, dtype=dtype)
        pin = getattr(self, "_pinned", None)
        if pin is None or pin.numel() < nbytes:
            pin = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, pin_memory=True)
            self._pinned = pin
        pin[:nbytes].copy_(dev_bytes[:nbytes], non_blocking=True)
        torch.cuda.synchronize(self.device)
        return pin[:nbytes].numpy().view(dtype)

    # ---- device-memory helpers (torch = allocator only) ------------------------------------------------------
    def to_device(self, a, cache: bool = False):
        """numpy -> device tensor; `cache=True` keeps the copy resident across calls (the raw scans are uploaded once)"""
        torch = _torch()
        if isinstance(a, torch.Tensor):
            return a
        if cache:
            hit = self._dev_cache.get(id(a))
            if hit is not None and hit[0] is a:
                return hit[1]
            t = self.to_device(a)
            self._dev_cache = {id(a): (a, t)}   # one resident batch at a time
            return t
        a = np.ascontiguousarray(a)
        if a.dtype.fields is not None:   # PointXYZIT records -> [..., 8] float32 view (same bytes)
            a = a.view(np.float32).reshape(a.shape + (8,))
        t = torch.from_numpy(a).to(f"cuda:{self.device}")
        torch.cuda.synchronize(self.device)
        return t

    def _as_device_cloud(self, cloud):
        """-> (device pointer, n_points, stride_bytes, keep-alive)"""
        t = self.to_device(cloud)
        assert t.dtype == _torch().float32 and t.shape[-1] >= 3
        t = t.reshape(-1, t.shape[-1]).contiguous()
        _torch().cuda.synchronize(self.device)
        return C.c_void_p(t.data_ptr()), t.shape[0], t.shape[1] * 4, t

    # ---- backend protocol of pipeline.run_calibration -----------------------------------------------------------
    def solve(self, pd, max_iterations, **kw):
        prob = CudaProblem(self, pd)
        try:
            return prob.solve(max_iterations, verbose=self.verbose, **kw)
        finally:
            prob.close()

    def build_surfel_map(self, cloud, leaf, lam):
        return CudaSurfelMap(self, cloud, leaf, lam)

    def batch_from_xyzi(self, scans_xyzi) -> CudaScanBatch:
        """import PCL-shaped scans [S, H, W, >=3] float32 (numpy or device tensor) as a packed scan batch"""
        t = self.to_device(scans_xyzi).contiguous()
        S, H, W = t.shape[0], t.shape[1], t.shape[2]
        h = C.c_void_p()
        _torch().cuda.synchronize(self.device)
        check(self.lib.lvi_scan_batch_from_xyzi_d(self.ctx, C.c_void_p(t.data_ptr()), t.shape[-1] * 4, S, H * W, C.byref(h)))
        return CudaScanBatch(self, h, S, H, W)

    def map_cloud(self, scans_in_map, keys=None):
        """the cloud the map is built from: all scans, or the key scans only (first association)"""
        if isinstance(scans_in_map, CudaScanBatch):
            return CudaMapCloud(scans_in_map, keys)
        a = scans_in_map if keys is None else scans_in_map[np.nonzero(keys)[0]]
        return a.reshape(-1, a.shape[-1])

    def associate(self, smap: CudaSurfelMap, scans_in_map, scans_raw, radius, k, step):
        torch = _torch()
        r = self.to_device(scans_raw, cache=True).contiguous()
        S, H, W = r.shape[0], r.shape[1], r.shape[2]
        n_out, n_all = C.c_int64(0), C.c_int64(0)
        cap = (S * H * W) // step + 1      # every emitted point is a scan point: one pass with a worst-case buffer instead of a sizing pass
        od = torch.empty(cap * 64, dtype=torch.uint8, device=r.device)
        torch.cuda.synchronize(self.device)
        if isinstance(scans_in_map, CudaScanBatch):
            assert scans_in_map.shape == (S, H, W)
            check(self.lib.lvi_associate_batch(self.ctx, smap.vmap, smap.surfels, scans_in_map.h, C.c_void_p(r.data_ptr()), W, H, radius, k, step,
                                               C.c_void_p(od.data_ptr()), cap, C.byref(n_out), C.byref(n_all)))
        else:
            m = self.to_device(scans_in_map).contiguous()
            assert m.shape[:3] == (S, H, W)
            torch.cuda.synchronize(self.device)
            check(self.lib.lvi_associate_d(self.ctx, smap.vmap, smap.surfels, C.c_void_p(m.data_ptr()), m.shape[-1] * 4, C.c_void_p(r.data_ptr()), S, W, H,
                                           radius, k, step, C.c_void_p(od.data_ptr()), cap, C.byref(n_out), C.byref(n_all)))
        self.last_n_all = n_all.value
        return self._download_records(od, n_out.value, SURFEL_POINT_DTYPE).copy()   # callers keep the points across associations

    def build_surfel_map_sharded(self, local_cloud, leaf, lam):
        return CudaSurfelMap(self, local_cloud, leaf, lam, sharded=True)

    def associate_sharded(self, smap: CudaSurfelMap, local_scans_in_map: CudaScanBatch, local_scans_raw, radius, k, step, total_points=None, download=True):
        """association of every rank's own scans against the gathered planes; returns the decimated points of ALL ranks (time order) on every
        rank.  total_points (points of all ranks) bounds the output so that one call suffices; without it a sizing call runs first."""
        torch = _torch()
        r = self.to_device(local_scans_raw, cache=True).contiguous()
        S, H, W = r.shape[0], r.shape[1], r.shape[2]
        assert local_scans_in_map.shape == (S, H, W)
        n_out, n_all = C.c_int64(0), C.c_int64(0)
        torch.cuda.synchronize(self.device)
        args = (self.ctx, smap.vmap, smap.surfels, local_scans_in_map.h, C.c_void_p(r.data_ptr()), W, H, radius, k, step)
        if total_points is None:
            check(self.lib.lvi_associate_sharded(*args, None, 0, C.byref(n_out), C.byref(n_all)))
            cap = max(n_out.value, 1)
        else:
            cap = total_points // step + self.world + 1
        od = torch.empty(cap * 64, dtype=torch.uint8, device=r.device)
        torch.cuda.synchronize(self.device)
        check(self.lib.lvi_associate_sharded(*args, C.c_void_p(od.data_ptr()), cap, C.byref(n_out), C.byref(n_all)))
        self.last_n_all = n_all.value
        if not download:   # the points stay in HBM (device tensor of 64-byte records, all ranks' points on every rank)
            return od[:n_out.value * 64]
        return self._download_records(od, n_out.value, SURFEL_POINT_DTYPE)

    def transform(self, scans_xyzi, poses):
        torch = _torch()
        if isinstance(scans_xyzi, CudaScanBatch):
            S = scans_xyzi.shape[0]
            poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(S, 16)
            h = C.c_void_p()
            check(self.lib.lvi_scan_batch_transform(self.ctx, scans_xyzi.h, ptr(poses), C.byref(h)))
            return CudaScanBatch(self, h, *scans_xyzi.shape)
        t = self.to_device(scans_xyzi).contiguous()
        S = t.shape[0]
        pts = int(np.prod(t.shape[1:-1]))
        assert t.shape[-1] == 8
        out = torch.empty_like(t)
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(S, 16)
        torch.cuda.synchronize(self.device)
        check(self.lib.lvi_transform_scans_d(self.ctx, C.c_void_p(t.data_ptr()), S, pts, ptr(poses), C.c_void_p(out.data_ptr())))
        return out

    def undistort(self, pd, scans_raw, target_time, correct_position):
        torch = _torch()
        r = self.to_device(scans_raw, cache=True).contiguous()
        S, H, W = r.shape[0], r.shape[1], r.shape[2]
        if target_time is None:   # undistortScan(): each scan expressed at its own stamp (first point's firing time)
            tt = r[:, 0, 0, 6:8].contiguous().cpu().numpy().view(np.float64).reshape(S).copy()
        else:
            tt = np.full(S, float(target_time))
        d = pd.desc()
        bad = C.c_int32(0)
        torch.cuda.synchronize(self.device)
        h = C.c_void_p()
        check(self.lib.lvi_scan_batch_undistort_d(self.ctx, C.byref(d), C.c_void_p(r.data_ptr()), S, H * W, ptr(tt), int(correct_position),
                                                  C.byref(h), C.byref(bad)))
        out = CudaScanBatch(self, h, S, H, W)
        if bad.value:
            raise IndexError(f"{bad.value} scan target time(s) outside the trajectory")
        return out

    def undistort_xyzi(self, pd, scans_raw, target_time, correct_position):
        """same through the PCL-layout entry point lvi_undistort_d: -> device tensor [S, H, W, 8] float32"""
        torch = _torch()
        r = self.to_device(scans_raw, cache=True).contiguous()
        S, H, W = r.shape[0], r.shape[1], r.shape[2]
        tt = (r[:, 0, 0, 6:8].contiguous().cpu().numpy().view(np.float64).reshape(S).copy() if target_time is None else np.full(S, float(target_time)))
        out = torch.empty((S, H, W, 8), dtype=torch.float32, device=r.device)
        d = pd.desc()
        bad = C.c_int32(0)
        torch.cuda.synchronize(self.device)
        check(self.lib.lvi_undistort_d(self.ctx, C.byref(d), C.c_void_p(r.data_ptr()), S, H * W, ptr(tt), int(correct_position),
                                       C.c_void_p(out.data_ptr()), C.byref(bad)))
        if bad.value:
            raise IndexError(f"{bad.value} scan target time(s) outside the trajectory")
        return out

    def traj_eval_many(self, pd, times):
        times = np.ascontiguousarray(times, dtype=np.float64)
        n = len(times)
        pos, quat, valid = np.zeros((n, 3)), np.zeros((n, 4)), np.zeros(n, np.uint8)
        d = pd.desc()
        check(self.lib.lvi_trajectory_evaluate(self.ctx, C.byref(d), ptr(times), n, ptr(pos), ptr(quat), ptr(valid)))
        return pos, quat, valid.astype(bool)

    def traj_eval(self, pd, t):
        pos, quat, valid = self.traj_eval_many(pd, [t])
        if not valid[0]:
            raise IndexError("time out of range for trajectory")
        return dict(p=pos[0], q=quat[0])

    def associate_landmarks(self, smap: CudaSurfelMap, pts, radius):
        pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
        out = np.zeros(len(pts), np.int32)
        check(self.lib.lvi_associate_landmarks(self.ctx, smap.surfels, ptr(pts), len(pts), radius, ptr(out)))
        return out

    def band_solve_dense(self, A, rhs, nb, nbo, bw, chain1_start=None, n_mid=0):
        A = np.ascontiguousarray(A, dtype=np.float64); rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        x = np.zeros(nb + nbo)
        check(self.lib.lvi_band_solve_dense(self.ctx, nb, nbo, bw, nb if chain1_start is None else chain1_start, n_mid, ptr(A), ptr(rhs), ptr(x)))
        return x
