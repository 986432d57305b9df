"""ctypes mirror of include/lvi_exc_b200.h (the C-ABI of the CUDA library).

This is plumbing only: structure layouts, the library loader and numpy<->pointer helpers.  There is no CPU
fallback — `load()` raises if the CUDA library has not been built, and every compute entry point of the
library itself returns LVI_ERR_NO_DEVICE on a host without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
LIB_PATH = Path(__file__).resolve().parent / "lib" / "liblvi_exc_b200.so"

LVI_OK = 0
LVI_ERR_INVALID, LVI_ERR_NO_DEVICE, LVI_ERR_CUDA, LVI_ERR_RANGE, LVI_ERR_DOMAIN = -1, -2, -3, -4, -5
LVI_ERR_OVERFLOW, LVI_ERR_NCCL, LVI_ERR_NUMERIC = -6, -7, -8
LVI_CONVERGENCE, LVI_NO_CONVERGENCE, LVI_FAILURE = 0, 1, 2
LVI_MAX_ITER_LOG = 256

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
c_uint8_p = C.POINTER(C.c_uint8)


class ProblemDesc(C.Structure):
    """lvi_problem_desc"""
    _fields_ = [
        ("t0", C.c_double), ("dt", C.c_double), ("n_knots", C.c_int32), ("_pad0", C.c_int32),
        ("r3_knots", c_double_p), ("so3_knots", c_double_p),
        ("lidar_q", c_double_p), ("lidar_p", c_double_p), ("cam_q", c_double_p), ("cam_p", c_double_p),
        ("gravity", c_double_p), ("acc_bias", c_double_p), ("gyr_bias", c_double_p),
        ("lidar_toff", C.c_double), ("cam_toff", C.c_double), ("imu_toff", C.c_double),
        ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double), ("readout", C.c_double), ("distortion", C.c_double * 5),
        ("cam_rows", C.c_int32), ("cam_cols", C.c_int32),
        ("n_landmarks", C.c_int32), ("n_planes", C.c_int32),
        ("rho", c_double_p), ("rho_locked", c_uint8_p), ("planes", c_double_p),
        ("lock_r3", C.c_int32), ("lock_so3", C.c_int32), ("lock_lidar_q", C.c_int32), ("lock_lidar_p", C.c_int32),
        ("lock_cam_q", C.c_int32), ("lock_cam_p", C.c_int32), ("lock_acc_bias", C.c_int32), ("lock_gyr_bias", C.c_int32),
        ("n_gyro", C.c_int32), ("_pad1", C.c_int32),
        ("gyro_t", c_double_p), ("gyro_w", c_double_p), ("gyro_weight", c_double_p),
        ("n_accel", C.c_int32), ("_pad2", C.c_int32),
        ("accel_t", c_double_p), ("accel_a", c_double_p), ("accel_weight", c_double_p),
        ("n_surfel", C.c_int32), ("_pad3", C.c_int32),
        ("surfel_t", c_double_p), ("surfel_tmap", c_double_p), ("surfel_point", c_double_p),
        ("surfel_plane", c_int32_p), ("surfel_weight", c_double_p), ("surfel_huber", c_double_p),
        ("n_cam", C.c_int32), ("_pad4", C.c_int32),
        ("cam_t0_ref", c_double_p), ("cam_t0_obs", c_double_p), ("cam_uv_ref", c_double_p), ("cam_uv_obs", c_double_p),
        ("cam_landmark", c_int32_p), ("cam_weight", c_double_p), ("cam_huber", c_double_p),
        ("n_camsurf", C.c_int32), ("_pad5", C.c_int32),
        ("cs_t", c_double_p), ("cs_tmap", c_double_p), ("cs_uv", c_double_p),
        ("cs_landmark", c_int32_p), ("cs_plane", c_int32_p), ("cs_weight", c_double_p), ("cs_huber", c_double_p),
        ("n_orient", C.c_int32), ("_pad6", C.c_int32),
        ("orient_t", c_double_p), ("orient_q", c_double_p), ("orient_weight", c_double_p),
    ]


class SolveOptions(C.Structure):
    """lvi_solve_options"""
    _fields_ = [
        ("max_num_iterations", C.c_int32), ("verbose", C.c_int32),
        ("initial_trust_region_radius", C.c_double), ("max_trust_region_radius", C.c_double),
        ("min_trust_region_radius", C.c_double), ("min_relative_decrease", C.c_double),
        ("min_lm_diagonal", C.c_double), ("max_lm_diagonal", C.c_double),
        ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double),
        ("max_num_consecutive_invalid_steps", C.c_int32), ("jacobi_scaling", C.c_int32),
    ]

    @staticmethod
    def default(max_iterations: int = 30, verbose: bool = False) -> "SolveOptions":
        # Ceres defaults in force at K/trajectory_estimator.h:38-68
        return SolveOptions(max_iterations, int(verbose), 1e4, 1e16, 1e-32, 1e-3, 1e-6, 1e32, 1e-6, 1e-10, 1e-8, 5, 1)


class SolveSummary(C.Structure):
    """lvi_solve_summary"""
    _fields_ = [
        ("termination_type", C.c_int32), ("num_iterations", C.c_int32),
        ("num_successful_steps", C.c_int32), ("num_unsuccessful_steps", C.c_int32),
        ("initial_cost", C.c_double), ("final_cost", C.c_double), ("fixed_cost", C.c_double),
        ("num_residual_blocks", C.c_int32), ("num_residuals", C.c_int32), ("num_effective_parameters", C.c_int32),
        ("band_width", C.c_int32), ("border_width", C.c_int32), ("_pad", C.c_int32),
        ("time_total_ms", C.c_double), ("time_jacobian_ms", C.c_double), ("time_linear_solve_ms", C.c_double),
        ("n_log", C.c_int32), ("_pad2", C.c_int32),
        ("log_cost", C.c_double * LVI_MAX_ITER_LOG), ("log_cost_change", C.c_double * LVI_MAX_ITER_LOG),
        ("log_gradient_max_norm", C.c_double * LVI_MAX_ITER_LOG), ("log_step_norm", C.c_double * LVI_MAX_ITER_LOG),
        ("log_radius", C.c_double * LVI_MAX_ITER_LOG), ("log_successful", C.c_uint8 * LVI_MAX_ITER_LOG),
    ]

    def brief(self) -> str:
        term = {0: "CONVERGENCE", 1: "NO_CONVERGENCE", 2: "FAILURE"}[self.termination_type]
        return (f"iterations: {self.num_iterations}, initial cost: {self.initial_cost:.6e}, "
                f"final cost: {self.final_cost:.6e}, termination: {term}")


SURFEL_POINT_DTYPE = np.dtype([("timestamp", "<f8"), ("point", "<f8", 3), ("point_in_map", "<f8", 3), ("plane_id", "<i8")])
RAW_POINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("_pad", "<f4"), ("intensity", "<f4"), ("_pad2", "<f4"),
                            ("timestamp", "<f8")])
assert SURFEL_POINT_DTYPE.itemsize == 64 and RAW_POINT_DTYPE.itemsize == 32


def ptr(a, ty=None):
    """pointer to a C-contiguous numpy array (None -> NULL)"""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    if ty is None:
        ty = {np.dtype("f8"): c_double_p, np.dtype("i4"): c_int32_p, np.dtype("i8"): c_int64_p,
              np.dtype("u1"): c_uint8_p}.get(a.dtype)
    if ty is None:
        return C.c_void_p(a.ctypes.data)
    return a.ctypes.data_as(ty)


class LibraryMissing(RuntimeError):
    pass


_lib = None

# every symbol include/lvi_exc_b200.h declares (tests/test_abi.py checks the built library exports all of them)
ABI_SYMBOLS = [
    "lvi_last_error", "lvi_abi_version", "lvi_abi_sizeof", "lvi_device_count", "lvi_ctx_create", "lvi_ctx_destroy", "lvi_ctx_synchronize",
    "lvi_ctx_stream", "lvi_ctx_launch_count", "lvi_ctx_kernel_timing", "lvi_ctx_kernel_times", "lvi_nccl_unique_id", "lvi_ctx_create_nccl",
    "lvi_voxel_build", "lvi_voxel_build_d", "lvi_voxel_destroy", "lvi_voxel_num_leaves", "lvi_voxel_num_points",
    "lvi_voxel_grid", "lvi_voxel_export", "lvi_surfel_extract", "lvi_surfel_destroy", "lvi_surfel_count",
    "lvi_surfel_export", "lvi_associate", "lvi_associate_d", "lvi_solve_options_default", "lvi_problem_create",
    "lvi_problem_destroy", "lvi_problem_solve", "lvi_problem_solve_cb", "lvi_problem_evaluate", "lvi_problem_num_residuals",
    "lvi_problem_num_tangent", "lvi_problem_tangent_offset_knot", "lvi_problem_tangent_offset_block",
    "lvi_problem_jacobian_dense", "lvi_problem_bench_iterations", "lvi_problem_layout", "lvi_associate_landmarks", "lvi_undistort",
    "lvi_undistort_d", "lvi_transform_scans", "lvi_transform_scans_d", "lvi_trajectory_evaluate", "lvi_trajectory_evaluate_full", "lvi_band_solve_dense",
    "lvi_scan_batch_undistort_d", "lvi_scan_batch_transform", "lvi_scan_batch_from_xyzi_d", "lvi_scan_batch_export_xyzi",
    "lvi_scan_batch_destroy", "lvi_scan_batch_num_points", "lvi_scan_batch_num_scans", "lvi_scan_batch_points_d",
    "lvi_voxel_build_batch", "lvi_associate_batch", "lvi_map_build_sharded", "lvi_associate_sharded",
]


def load() -> C.CDLL:
    """Load the CUDA library built by `__graft_entry__.build()`; raise loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("LVI_EXC_B200_LIB", LIB_PATH))
    if not path.exists():
        raise LibraryMissing(f"{path} not found: run `python -c 'import __graft_entry__ as g; g.build()'` first "
                             "(there is no CPU fallback)")
    lib = C.CDLL(str(path))
    vp = C.c_void_p
    lib.lvi_last_error.restype = C.c_char_p
    lib.lvi_ctx_create.argtypes = [C.c_int, vp, C.c_int, C.c_int, C.POINTER(vp)]
    lib.lvi_ctx_create_nccl.argtypes = [C.c_int, vp, C.c_int, C.c_int, C.POINTER(vp)]
    lib.lvi_nccl_unique_id.argtypes = [vp]
    lib.lvi_ctx_destroy.argtypes = [vp]
    lib.lvi_ctx_synchronize.argtypes = [vp]
    lib.lvi_ctx_stream.argtypes = [vp]
    lib.lvi_ctx_stream.restype = vp
    lib.lvi_ctx_launch_count.argtypes = [vp]
    lib.lvi_ctx_launch_count.restype = C.c_int64
    lib.lvi_ctx_kernel_timing.argtypes = [vp, C.c_int]
    lib.lvi_ctx_kernel_times.argtypes = [vp, C.c_char_p, C.c_int64]
    lib.lvi_ctx_kernel_times.restype = C.c_int64
    for name in ("lvi_voxel_build", "lvi_voxel_build_d"):
        getattr(lib, name).argtypes = [vp, vp, C.c_size_t, C.c_int64, C.c_float, C.c_int, C.c_double, C.POINTER(vp)]
    lib.lvi_voxel_destroy.argtypes = [vp]
    lib.lvi_voxel_num_leaves.argtypes = [vp]
    lib.lvi_voxel_num_leaves.restype = C.c_int64
    lib.lvi_voxel_num_points.argtypes = [vp]
    lib.lvi_voxel_num_points.restype = C.c_int64
    lib.lvi_voxel_grid.argtypes = [vp, c_int32_p, c_int32_p]
    lib.lvi_voxel_export.argtypes = [vp, vp, c_int64_p, c_int32_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                     c_int64_p, c_int32_p]
    lib.lvi_surfel_extract.argtypes = [vp, vp, C.c_double, C.c_int, C.c_float, C.c_int, C.POINTER(vp)]
    lib.lvi_surfel_destroy.argtypes = [vp]
    lib.lvi_surfel_count.argtypes = [vp]
    lib.lvi_surfel_count.restype = C.c_int64
    lib.lvi_surfel_export.argtypes = [vp, vp, c_double_p, c_double_p, c_double_p, c_double_p, c_int64_p, c_int32_p]
    for name in ("lvi_associate", "lvi_associate_d"):
        getattr(lib, name).argtypes = [vp, vp, vp, vp, C.c_size_t, vp, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int32,
                                       C.c_int32, vp, C.c_int64, c_int64_p, c_int64_p]
    lib.lvi_solve_options_default.argtypes = [C.POINTER(SolveOptions)]
    lib.lvi_solve_options_default.restype = None
    lib.lvi_problem_create.argtypes = [vp, C.POINTER(ProblemDesc), C.POINTER(vp)]
    lib.lvi_problem_destroy.argtypes = [vp]
    lib.lvi_problem_solve.argtypes = [vp, C.POINTER(SolveOptions), C.POINTER(SolveSummary)]
    lib.lvi_problem_evaluate.argtypes = [vp, c_double_p, c_double_p, c_double_p]
    lib.lvi_problem_num_residuals.argtypes = [vp]
    lib.lvi_problem_num_tangent.argtypes = [vp]
    lib.lvi_problem_tangent_offset_knot.argtypes = [vp, C.c_int]
    lib.lvi_problem_tangent_offset_block.argtypes = [vp, C.c_int]
    lib.lvi_problem_jacobian_dense.argtypes = [vp, c_double_p]
    lib.lvi_problem_bench_iterations.argtypes = [vp, C.c_int, C.POINTER(C.c_float)]
    lib.lvi_problem_layout.argtypes = [vp, c_int32_p]
    lib.lvi_associate_landmarks.argtypes = [vp, vp, c_double_p, C.c_int64, C.c_double, c_int32_p]
    for name in ("lvi_undistort", "lvi_undistort_d"):
        getattr(lib, name).argtypes = [vp, C.POINTER(ProblemDesc), vp, C.c_int32, C.c_int64, c_double_p, C.c_int, vp, c_int32_p]
    for name in ("lvi_transform_scans", "lvi_transform_scans_d"):
        getattr(lib, name).argtypes = [vp, vp, C.c_int32, C.c_int64, c_double_p, vp]
    lib.lvi_trajectory_evaluate.argtypes = [vp, C.POINTER(ProblemDesc), c_double_p, C.c_int64, c_double_p, c_double_p, c_uint8_p]
    lib.lvi_trajectory_evaluate_full.argtypes = [vp, C.POINTER(ProblemDesc), c_double_p, C.c_int64, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_uint8_p]
    lib.lvi_scan_batch_undistort_d.argtypes = [vp, C.POINTER(ProblemDesc), vp, C.c_int32, C.c_int64, c_double_p, C.c_int, C.POINTER(vp), c_int32_p]
    lib.lvi_scan_batch_transform.argtypes = [vp, vp, c_double_p, C.POINTER(vp)]
    lib.lvi_scan_batch_from_xyzi_d.argtypes = [vp, vp, C.c_size_t, C.c_int32, C.c_int64, C.POINTER(vp)]
    lib.lvi_scan_batch_export_xyzi.argtypes = [vp, vp, vp, C.c_int]
    lib.lvi_scan_batch_destroy.argtypes = [vp]
    lib.lvi_scan_batch_num_points.argtypes = [vp]
    lib.lvi_scan_batch_num_points.restype = C.c_int64
    lib.lvi_scan_batch_num_scans.argtypes = [vp]
    lib.lvi_scan_batch_num_scans.restype = C.c_int32
    lib.lvi_scan_batch_points_d.argtypes = [vp]
    lib.lvi_scan_batch_points_d.restype = vp
    lib.lvi_voxel_build_batch.argtypes = [vp, vp, c_uint8_p, C.c_float, C.c_int, C.c_double, C.POINTER(vp)]
    lib.lvi_associate_batch.argtypes = [vp, vp, vp, vp, vp, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_int32, vp, C.c_int64, c_int64_p, c_int64_p]
    lib.lvi_associate_sharded.argtypes = [vp, vp, vp, vp, vp, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_int32, vp, C.c_int64, c_int64_p, c_int64_p]
    lib.lvi_map_build_sharded.argtypes = [vp, vp, c_uint8_p, C.c_float, C.c_int, C.c_double, C.c_double, C.c_int, C.c_float, C.c_int, C.POINTER(vp), C.POINTER(vp), c_int64_p]
    lib.lvi_band_solve_dense.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p]
    _lib = lib
    return lib


class LviError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"lvi error {code}: {msg}")
        self.code = code


def check(code: int) -> None:
    """Map C-ABI status codes onto the exception types the reference throws."""
    if code == LVI_OK:
        return
    msg = load().lvi_last_error().decode()
    if code == LVI_ERR_RANGE:
        raise IndexError(msg)       # std::range_error
    if code == LVI_ERR_DOMAIN:
        raise ValueError(msg)       # std::domain_error
    if code == LVI_ERR_OVERFLOW:
        raise OverflowError(msg)
    raise LviError(code, msg)
