/* lvi_exc_b200.h — C-ABI of the B200-native LVI-ExC calibration hot path.
 *
 * The reference (peterWon/LVI-ExC) has no FFI: its boundary is the C++ template API that
 * `lvi_init_orb_surfel` calls.  This header is the flat, torch-free surface the C++
 * facade (`include/lvi_exc_b200/…`, Kontiki/LI-Calib/pclomp/Ceres-shaped names) lowers onto.
 * Every entry point cites the reference interface it replaces.  Paths are relative to
 * /root/reference/src ; `L/` = lvi_exc/, `K/` = lvi_exc/thirdparty/Kontiki/include/kontiki/,
 * `N/` = ndt_omp/include/pclomp/.
 *
 * Conventions
 *   - all pointers are HOST pointers unless the name ends in `_d` (device pointer on the ctx device);
 *   - caller owns every input/output buffer, the library owns handles;
 *   - every function returns LVI_OK (0) or a negative lvi_status; lvi_last_error() has the text;
 *   - quaternions are stored x,y,z,w (Eigen coeffs() order, K/trajectories/uniform_so3_spline_trajectory.h);
 *   - no CPU fallback exists: with no CUDA device every compute call returns LVI_ERR_NO_DEVICE.
 */
#ifndef LVI_EXC_B200_H
#define LVI_EXC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LVI_ABI_VERSION 1

typedef enum lvi_status {
  LVI_OK = 0,
  LVI_ERR_INVALID = -1,      /* bad argument */
  LVI_ERR_NO_DEVICE = -2,    /* no CUDA device / wrong architecture */
  LVI_ERR_CUDA = -3,         /* CUDA runtime failure */
  LVI_ERR_RANGE = -4,        /* std::range_error in the reference (time outside the spline, K/trajectory_estimator.h:111-122) */
  LVI_ERR_DOMAIN = -5,       /* std::domain_error (non-unit control point, K/trajectories/uniform_so3_spline_trajectory.h:23-27) */
  LVI_ERR_OVERFLOW = -6,     /* voxel index overflow guard, N/voxel_grid_covariance_omp_impl.hpp:75-84 */
  LVI_ERR_NCCL = -7,
  LVI_ERR_NUMERIC = -8       /* Cholesky breakdown etc. */
} lvi_status;

typedef struct lvi_ctx lvi_ctx;         /* one CUDA device + stream (+ optional NCCL communicator) */
typedef struct lvi_voxel_map lvi_voxel_map; /* replaces pclomp::VoxelGridCovariance (N/voxel_grid_covariance_omp.h:92-300) */
typedef struct lvi_surfel_set lvi_surfel_set; /* replaces SurfelAssociation::surfel_planes_ (L/include/core/surfel_association.h:48-55) */
typedef struct lvi_scan_batch lvi_scan_batch; /* replaces ScanUndistortion::scan_data_in_map_ / map_cloud_ (L/include/core/scan_undistortion.h:182-188) */
typedef struct lvi_problem lvi_problem; /* replaces kontiki::TrajectoryEstimator + ceres::Problem (K/trajectory_estimator.h:19-135) */

const char* lvi_last_error(void);
int lvi_abi_version(void);
/* sizeof of the ABI structures, for binding generators: 0 lvi_problem_desc, 1 lvi_solve_options, 2 lvi_solve_summary,
 * 3 lvi_point_xyzit, 4 lvi_surfel_point (defined below); -1 otherwise */
int64_t lvi_abi_sizeof(int which);
/* number of CUDA devices visible (0 on a CPU-only host; never an error) */
int lvi_device_count(void);

/* ---- context -------------------------------------------------------------------------------------------- */
/* `nccl_comm` is an ncclComm_t owned by the caller (may be NULL: single-GPU). rank/world describe the
 * data-parallel sharding of SURVEY §8(e); with world==1 no collective is issued. */
int lvi_ctx_create(int device, void* nccl_comm, int rank, int world, lvi_ctx** out);
int lvi_ctx_destroy(lvi_ctx* ctx);
int lvi_ctx_synchronize(lvi_ctx* ctx);
/* the cudaStream_t all work of this context is enqueued on (for external CUDA-event timing) */
void* lvi_ctx_stream(lvi_ctx* ctx);
/* number of kernels this context has launched so far (bench.py `gpu_launches`) */
int64_t lvi_ctx_launch_count(lvi_ctx* ctx);
/* measurement: with timing enabled every kernel this context launches is bracketed by CUDA events on the stream it is launched on;
 * lvi_ctx_kernel_times waits for the streams and drains what was recorded as text lines "kernel_name launches total_ms" (returns the
 * length the full text needs; bench.py's per-kernel roofline figures come from here) */
int lvi_ctx_kernel_timing(lvi_ctx* ctx, int enable);
int64_t lvi_ctx_kernel_times(lvi_ctx* ctx, char* out, int64_t cap);
/* host-side NCCL bootstrap helpers so a Python/C host can create the communicator without linking NCCL */
int lvi_nccl_unique_id(void* id128 /* 128 bytes out */);
int lvi_ctx_create_nccl(int device, const void* id128, int rank, int world, lvi_ctx** out);

/* ---- (a-1) NDT voxel covariance build ------------------------------------------------------------------ */
/* Replaces VoxelGridCovariance::applyFilter (N/voxel_grid_covariance_omp_impl.hpp:49-374) as reached from
 * NDT::setInputTarget (N/ndt_omp.h:117-122,275-282; callers L/src/core/lidar_odometry.cpp:102 and
 * L/test/lvi_initialize_surfel_orb.cpp:1185).
 *   xyz          : points, `stride_bytes` apart (pcl::PointXYZI = 32), x,y,z float at offset 0; non-finite skipped
 *   leaf_size    : NDT resolution (cfg ndtResolution 0.5)
 *   min_points   : min_points_per_voxel_ (6, N/voxel_grid_covariance_omp.h:205)
 *   eig_mult     : min_covar_eigvalue_mult_ (0.01, :206)
 * The map keeps the (sorted) per-leaf point lists (Leaf::pointList_) on the device. */
int lvi_voxel_build(lvi_ctx* ctx, const void* xyz, size_t stride_bytes, int64_t n_points,
                    float leaf_size, int min_points, double eig_mult, lvi_voxel_map** out);
/* same, input already resident in HBM (x,y,z float, stride_bytes apart) */
int lvi_voxel_build_d(lvi_ctx* ctx, const void* xyz_d, size_t stride_bytes, int64_t n_points,
                      float leaf_size, int min_points, double eig_mult, lvi_voxel_map** out);
int lvi_voxel_destroy(lvi_voxel_map* m);
int64_t lvi_voxel_num_leaves(const lvi_voxel_map* m);
int64_t lvi_voxel_num_points(const lvi_voxel_map* m); /* finite points that were binned */
/* grid geometry: min_b_[3], div_b_[3] (N/voxel_grid_covariance_omp_impl.hpp:87-103) */
int lvi_voxel_grid(const lvi_voxel_map* m, int32_t min_b[3], int32_t div_b[3]);
/* getLeaves(): leaves in ascending linear-index order (std::map order). Any pointer may be NULL.
 *   keys[L] int64 ; nr_points[L] int32 (-1 = rejected leaf, :343-346,368-372) ; mean[L*3] ; cov[L*9] row-major ;
 *   evals[L*3] ascending ; evecs[L*9] row-major, columns are eigenvectors ; icov[L*9] ;
 *   leaf_start[L+1] CSR offsets into point_index ; point_index[num_points] original cloud indices, cloud order per leaf */
int lvi_voxel_export(lvi_ctx* ctx, const lvi_voxel_map* m, int64_t* keys, int32_t* nr_points, double* mean,
                     double* cov, double* evals, double* evecs, double* icov, int64_t* leaf_start,
                     int32_t* point_index);

/* ---- (a-2) surfel extraction ---------------------------------------------------------------------------- */
/* Replaces SurfelAssociation::setSurfelMap / checkPlaneType / fitPlane (L/src/core/surfel_association.cpp:50-108,
 * 246-294).  lambda = plane_lambda (0.6 first pass, 0.7 refine; T:127,1182), min_leaf_points 10 (:63),
 * ransac_threshold 0.05 (:279), min_inliers 20 (:284). plane_id = position in ascending-leaf order. */
int lvi_surfel_extract(lvi_ctx* ctx, const lvi_voxel_map* m, double lambda, int min_leaf_points,
                       float ransac_threshold, int min_inliers, lvi_surfel_set** out);
int lvi_surfel_destroy(lvi_surfel_set* s);
int64_t lvi_surfel_count(const lvi_surfel_set* s);
/* p4[P*4] (n,d) ; Pi[P*3] = -d*n ; box_min/max[P*3] ; leaf_key[P] ; n_inliers[P]. Any pointer may be NULL. */
int lvi_surfel_export(lvi_ctx* ctx, const lvi_surfel_set* s, double* p4, double* Pi, double* box_min,
                      double* box_max, int64_t* leaf_key, int32_t* n_inliers);

/* ---- (a-3) scan -> surfel association ------------------------------------------------------------------- */
/* One raw scan point as laid out by the reference's PointXYZIT (L/include/utils/pcl_utils.h:39-58): 32 bytes. */
typedef struct lvi_point_xyzit { float x, y, z, _pad; float intensity; float _pad2; double timestamp; } lvi_point_xyzit;
/* One associated point = SurfelAssociation::SurfelPoint (L/include/core/surfel_association.h:41-46) */
typedef struct lvi_surfel_point { double timestamp; double point[3]; double point_in_map[3]; int64_t plane_id; } lvi_surfel_point;

/* Replaces the per-scan loop `getAssociation(scan_inM, scan_raw, k)` + `averageTimeDownSmaple(step)`
 * (L/src/core/surfel_association.cpp:111-159,240-244,305-331; driver T:1191-1199) over a batch of `n_scans`
 * organised scans (each W x H, index h*W+w).
 *   scans_in_map : n_scans*W*H points, x,y,z float at `map_stride_bytes` spacing (pcl::PointXYZI = 32); NaN = no return
 *   scans_raw    : n_scans*W*H lvi_point_xyzit
 *   radius       : associated_radius (0.05) ; k_per_ring : selected_num_per_ring (2) ; time_step : 10
 * Output: `*n_out` downsampled points written to `out` (capacity `cap`; call with out==NULL to size), in the
 * reference's emission order; `*n_all` = spoints_all_.size() before the every-`time_step` decimation. */
int lvi_associate(lvi_ctx* ctx, const lvi_voxel_map* m, const lvi_surfel_set* s, const void* scans_in_map,
                  size_t map_stride_bytes, const lvi_point_xyzit* scans_raw, int32_t n_scans, int32_t W, int32_t H,
                  double radius, int32_t k_per_ring, int32_t time_step, lvi_surfel_point* out, int64_t cap,
                  int64_t* n_out, int64_t* n_all);
/* device-resident variant: inputs in HBM; output stays in HBM at out_d (capacity cap) */
int lvi_associate_d(lvi_ctx* ctx, const lvi_voxel_map* m, const lvi_surfel_set* s, const void* scans_in_map_d,
                    size_t map_stride_bytes, const lvi_point_xyzit* scans_raw_d, int32_t n_scans, int32_t W,
                    int32_t H, double radius, int32_t k_per_ring, int32_t time_step, lvi_surfel_point* out_d,
                    int64_t cap, int64_t* n_out, int64_t* n_all);

/* ---- (a-4..a-13) continuous-time B-spline least squares ------------------------------------------------- */
/* Flat description of one TrajectoryEstimator problem.  It is what the C++ facade
 * records from AddMeasurement<M>() calls (K/trajectory_estimator.h:71-74) and lock flags
 * (L/src/core/trajectory_manager_lvi.cpp:202-227,314-327).  In/out arrays are updated in place by
 * lvi_problem_solve, like Ceres updates DynamicParameterStore memory (K/../entity/paramstore/dynamic_pstore.h:25-30). */
typedef struct lvi_problem_desc {
  /* SplitTrajectory(r3_dt, so3_dt, r3_t0, so3_t0) with equal dt/t0 (L/include/core/trajectory_manager_lvi.h:120-124) */
  double t0, dt;
  int32_t n_knots;
  int32_t _pad0;
  double* r3_knots;   /* [n_knots*3]  in/out */
  double* so3_knots;  /* [n_knots*4]  in/out, unit quaternions x,y,z,w */
  /* sensors: relative pose blocks [q4, p3, t_off1] (K/sensors/sensors.h:93-167); t_off is locked
   * (cfg optimize_time_offset false) and enters as a constant */
  double* lidar_q;  double* lidar_p;   /* in/out */
  double* cam_q;    double* cam_p;     /* in/out */
  double* gravity;  /* [2] roll,pitch in/out (K/sensors/imu.h:41-70), never locked (Q5) */
  double* acc_bias; double* gyr_bias;  /* [3] in/out (K/sensors/constant_bias_imu.h:51-61) */
  double lidar_toff, cam_toff, imu_toff;
  /* PinholeCamera(rows, cols, readout, k1, k2, p1, p2, k3, fx, fy, cx, cy) (K/sensors/pinhole_camera.h:253-281).  distortion = k1 k2 p1 p2 k3
   * of the radial-tangential model (:199-215); it is applied when |k1|, |k2| or |p1| exceeds 1e-5 -- the reference's own test (:78, which
   * never looks at p2 or k3): Project distorts the normalised point (:217-240), Unproject inverts by 8 fixed-point steps (:168-187).
   * lvi.yaml ships all zeros (Q13). */
  double fx, fy, cx, cy, readout;
  double distortion[5];
  int32_t cam_rows, cam_cols;
  /* landmarks: inverse depth blocks, lower bound 0 (K/measurements/static_rscamera_measurement.h:183-189) */
  int32_t n_landmarks;
  int32_t n_planes;
  double* rho;                 /* [n_landmarks] in/out */
  const uint8_t* rho_locked;   /* [n_landmarks] or NULL (Landmark::Lock) */
  const double* planes;        /* [n_planes*3] closest-point vectors Pi, constant (K/measurements/lidar_surfel_point.h:185-190) */
  /* lock flags (1 = SetParameterBlockConstant) */
  int32_t lock_r3, lock_so3, lock_lidar_q, lock_lidar_p, lock_cam_q, lock_cam_p, lock_acc_bias, lock_gyr_bias;
  /* GyroscopeMeasurement(imu, t, w, weight)  K/measurements/gyroscope_measurement.h:19-48 */
  int32_t n_gyro; int32_t _pad1;
  const double* gyro_t; const double* gyro_w; const double* gyro_weight;      /* [n],[n*3],[n] */
  /* AccelerometerMeasurement(imu, t, a, weight)  K/measurements/accelerometer_measurement.h:20-49 */
  int32_t n_accel; int32_t _pad2;
  const double* accel_t; const double* accel_a; const double* accel_weight;
  /* LiDARSurfelPoint(lidar, point, plane, t, map_time, huber, weight)  K/measurements/lidar_surfel_point.h:18-27 */
  int32_t n_surfel; int32_t _pad3;
  const double* surfel_t; const double* surfel_tmap; const double* surfel_point; /* [n],[n],[n*3] */
  const int32_t* surfel_plane; const double* surfel_weight; const double* surfel_huber;
  /* StaticRsCameraMeasurement(camera, obs, huber, weight)  K/measurements/static_rscamera_measurement.h:67-109;
   * ref = obs->landmark()->reference().  t0 = View::t0(), uv = Observation::uv() */
  int32_t n_cam; int32_t _pad4;
  const double* cam_t0_ref; const double* cam_t0_obs; const double* cam_uv_ref; const double* cam_uv_obs; /* [n],[n],[n*2],[n*2] */
  const int32_t* cam_landmark; const double* cam_weight; const double* cam_huber;
  /* CameraSurfelLandmark(cam, lidar, landmark, plane, t, map_time, huber, weight)  K/measurements/camera_surfel_landmark.h:19-26 */
  int32_t n_camsurf; int32_t _pad5;
  const double* cs_t; const double* cs_tmap; const double* cs_uv;  /* [n],[n],[n*2] reference uv */
  const int32_t* cs_landmark; const int32_t* cs_plane; const double* cs_weight; const double* cs_huber;
  /* OrientationMeasurement(t, q, weight)  K/measurements/orientation_measurement.h:20-22 */
  int32_t n_orient; int32_t _pad6;
  const double* orient_t; const double* orient_q; const double* orient_weight; /* [n],[n*4],[n] */
} lvi_problem_desc;

/* Solver::Options fields that K/trajectory_estimator.h:38-68 sets, plus the Ceres (<=2.1) defaults in force
 * (SURVEY Appendix C).  Zero-initialise then call lvi_solve_options_default(). */
typedef struct lvi_solve_options {
  int32_t max_num_iterations;          /* Solve(max_iterations=30) */
  int32_t verbose;                     /* minimizer_progress_to_stdout */
  double initial_trust_region_radius;  /* 1e4 */
  double max_trust_region_radius;      /* 1e16 */
  double min_trust_region_radius;      /* 1e-32 */
  double min_relative_decrease;        /* 1e-3 */
  double min_lm_diagonal;              /* 1e-6 */
  double max_lm_diagonal;              /* 1e32 */
  double function_tolerance;           /* 1e-6 */
  double gradient_tolerance;           /* 1e-10 */
  double parameter_tolerance;          /* 1e-8 */
  int32_t max_num_consecutive_invalid_steps; /* 5 */
  int32_t jacobi_scaling;              /* 1 */
} lvi_solve_options;
void lvi_solve_options_default(lvi_solve_options* o);

enum { LVI_CONVERGENCE = 0, LVI_NO_CONVERGENCE = 1, LVI_FAILURE = 2, LVI_USER_SUCCESS = 3, LVI_USER_FAILURE = 4 };
#define LVI_MAX_ITER_LOG 256
/* ceres::Solver::Summary subset (BriefReport fields) + per-iteration log (IterationSummary) */
typedef struct lvi_solve_summary {
  int32_t termination_type;     /* LVI_CONVERGENCE / LVI_NO_CONVERGENCE / LVI_FAILURE */
  int32_t num_iterations;       /* iterations.size()-1 : LM steps attempted */
  int32_t num_successful_steps;
  int32_t num_unsuccessful_steps;
  double initial_cost, final_cost, fixed_cost;
  int32_t num_residual_blocks, num_residuals, num_effective_parameters, band_width, border_width, _pad;
  double time_total_ms, time_jacobian_ms, time_linear_solve_ms;
  int32_t n_log;
  int32_t _pad2;
  double log_cost[LVI_MAX_ITER_LOG];
  double log_cost_change[LVI_MAX_ITER_LOG];
  double log_gradient_max_norm[LVI_MAX_ITER_LOG];
  double log_step_norm[LVI_MAX_ITER_LOG];
  double log_radius[LVI_MAX_ITER_LOG];
  uint8_t log_successful[LVI_MAX_ITER_LOG];
} lvi_solve_summary;

/* Records + lowers the problem to device SoA tables; parameters are copied to HBM and stay resident. */
int lvi_problem_create(lvi_ctx* ctx, const lvi_problem_desc* desc, lvi_problem** out);
int lvi_problem_destroy(lvi_problem* p);
/* ceres::Solve with TRUST_REGION / LEVENBERG_MARQUARDT / exact Schur-equivalent linear solve
 * (K/trajectory_estimator.h:38-68); writes the optimum back into the desc's in/out arrays. Under world>1
 * every rank passes the SAME full problem; the library evaluates a contiguous time chunk of every residual table
 * per rank and all-reduces the normal equations (NCCL), so every rank returns identical parameters. */
int lvi_problem_solve(lvi_problem* p, const lvi_solve_options* opt, lvi_solve_summary* summary);
/* ceres::IterationCallback (K/trajectory_estimator.h:88-94 AddCallback; L/include/utils/ceres_callbacks.h:31-66): called after the
 * evaluation of iteration 0 and after every LM iteration with ceres::IterationSummary's fields.  Return 0 = SOLVER_CONTINUE,
 * 1 = SOLVER_ABORT, 2 = SOLVER_TERMINATE_SUCCESSFULLY.  With update_state != 0 (Solver::Options::update_state_every_iteration, the
 * needs_state flag of AddCallback) the caller's in/out parameter arrays hold the current iterate when the callback runs. */
typedef struct lvi_iteration_summary {
  int32_t iteration; int32_t step_is_successful;
  double cost, cost_change, gradient_max_norm, step_norm, trust_region_radius;
} lvi_iteration_summary;
typedef int (*lvi_iteration_callback)(const lvi_iteration_summary* it, void* user);
int lvi_problem_solve_cb(lvi_problem* p, const lvi_solve_options* opt, lvi_solve_summary* summary, lvi_iteration_callback cb, void* user,
                         int update_state);
/* One evaluation at the current parameters (ceres::Problem::Evaluate analogue, used for parity tests):
 *   cost (with loss, excluding fixed cost) ; residuals[num_residuals] after loss correction, in table order
 *   gyro,accel,surfel,cam,camsurf,orient ; gradient[num_effective_parameters] in the library's tangent order
 * Any pointer may be NULL. */
int lvi_problem_evaluate(lvi_problem* p, double* cost, double* residuals, double* gradient);
/* tangent layout helpers for tests: offsets of knot i (6 dims: r3 then so3) and of the sensor blocks in `gradient` */
int lvi_problem_num_residuals(const lvi_problem* p);
int lvi_problem_num_tangent(const lvi_problem* p);
int lvi_problem_tangent_offset_knot(const lvi_problem* p, int knot);   /* -1 if the block is constant */
/* which: 0 lidar_q 1 lidar_p 2 cam_q 3 cam_p 4 gravity 5 acc_bias 6 gyr_bias ; 7+l : rho_l */
int lvi_problem_tangent_offset_block(const lvi_problem* p, int which);
/* dense Jacobian (after loss correction and local parameterisation, before Jacobi scaling), row-major
 * [num_residuals x num_tangent]; only for small test problems. */
int lvi_problem_jacobian_dense(lvi_problem* p, double* J);
/* run `iters` LM iterations' worth of work without convergence tests (bench.py `value`): each = residuals +
 * Jacobians + normal equations + damped solve + trial-step cost. Parameters are restored afterwards. */
int lvi_problem_bench_iterations(lvi_problem* p, int iters,
                                 float* ms_per_phase /* [6] linearize, build_system, band_factor, corner+backsolve, trial cost, whole loop / iters */);
/* linear-system layout: out[8] = band dims, border dims, half bandwidth, block columns, sub-diagonal tile rows, border tile rows,
 * first position of the second chain of the two-sided ordering (== band dims: one chain), unused padding positions before it */
int lvi_problem_layout(lvi_problem* p, int32_t* out);

/* ---- (a-3') visual landmark -> surfel association ------------------------------------------------------- */
/* Inner test of SurfelAssociation::associateVisualPointsWithPlanes (L/src/core/surfel_association.cpp:196-210) for
 * landmark positions already expressed in the map (L0) frame: strict bbox test + point2PlaneDistance <= 2*radius;
 * the LAST matching plane wins (:206).  pts[n*3] double; plane_out[n] = plane id or -1. */
int lvi_associate_landmarks(lvi_ctx* ctx, const lvi_surfel_set* s, const double* pts, int64_t n, double radius,
                            int32_t* plane_out);

/* ---- (f-1) scan undistortion / map assembly -------------------------------------------------------------- */
/* Replaces ScanUndistortion::undistort (L/include/core/scan_undistortion.h:132-180) for a batch of n_scans raw scans of
 * pts_per_scan points: out = q_G_to_target * (q_Lk_to_G * p + [correct_position] (p_Lk_in_G - p_target_in_G)), with
 * the LiDAR pose evaluated from the spline at each point's own timestamp (TrajectoryManagerLVI::evaluateLidarPose,
 * L/src/core/trajectory_manager_lvi.cpp:398-408).  target_time[s] is the time scan s is expressed at: its own stamp
 * for undistortScan() (:40-57), the map time for undistortScanInMap() (:59-74).  Output points are
 * pcl::PointXYZI-shaped (32 B: x,y,z,1,intensity,0,0,0); NaN in -> NaN out; a point whose time is outside the
 * trajectory stays a default-constructed (zero) point as in the reference.  *n_bad_targets (may be NULL) counts scans
 * whose target time is outside the trajectory (all their points are NaN).
 * `traj` uses t0, dt, n_knots, r3_knots, so3_knots, lidar_q, lidar_p, lidar_toff (host pointers). */
int lvi_undistort(lvi_ctx* ctx, const lvi_problem_desc* traj, const lvi_point_xyzit* scans_raw, int32_t n_scans,
                  int64_t pts_per_scan, const double* target_time, int correct_position, void* out_xyzi,
                  int32_t* n_bad_targets);
int lvi_undistort_d(lvi_ctx* ctx, const lvi_problem_desc* traj, const lvi_point_xyzit* scans_raw_d, int32_t n_scans,
                    int64_t pts_per_scan, const double* target_time, int correct_position, void* out_xyzi_d,
                    int32_t* n_bad_targets);
/* IMU pose of the trajectory (SplitTrajectory::Evaluate, K/trajectories/split_trajectory.h:41-58; Position|Orientation)
 * at n times: pos[n*3], quat[n*4] (x,y,z,w), valid[n] = 0 where t is outside [MinTime, MaxTime). */
int lvi_trajectory_evaluate(lvi_ctx* ctx, const lvi_problem_desc* traj, const double* t, int64_t n, double* pos,
                            double* quat, uint8_t* valid);
/* every quantity of kontiki::trajectories::TrajectoryEvaluation (K/trajectories/trajectory.h:27-37; Trajectory::Position / Velocity /
 * Acceleration / Orientation / AngularVelocity, :95-133) at n times: pos/vel/acc[n*3] of the R3 spline, quat[n*4] (x,y,z,w) and the
 * WORLD-frame angular velocity[n*3] of the SO3 spline.  Any output may be NULL; r3_knots may be NULL (SO3-only trajectory). */
int lvi_trajectory_evaluate_full(lvi_ctx* ctx, const lvi_problem_desc* traj, const double* t, int64_t n, double* pos, double* vel,
                                 double* acc, double* quat, double* angvel, uint8_t* valid);
/* pcl::transformPointCloud(scan, out, pose) with the pose cast to a float 4x4 (L/include/core/scan_undistortion.h:111,
 * L/src/core/lidar_odometry.cpp:98), one pose per scan: poses[n_scans*16] row-major double. In/out PointXYZI (32 B). */
int lvi_transform_scans(lvi_ctx* ctx, const void* scans_xyzi, int32_t n_scans, int64_t pts_per_scan,
                        const double* poses, void* out_xyzi);
int lvi_transform_scans_d(lvi_ctx* ctx, const void* scans_xyzi_d, int32_t n_scans, int64_t pts_per_scan,
                          const double* poses, void* out_xyzi_d);

/* ---- scan batches: the map-frame scans kept in HBM between the steps of one data association ------------------------------- */
/* The reference's ScanUndistortion keeps the de-skewed scans (scan_data_in_map_) and their concatenation (map_cloud_) as PCL clouds and
 * hands them to LiDAROdometry / SurfelAssociation (L/include/core/scan_undistortion.h:59-116,182-188; driver T:1169-1210).  A scan batch is
 * that object on the device: n_scans organised scans of pts_per_scan points in ONE frame, stored packed (16 B per point: x, y, z,
 * intensity) with each scan's min/max, so the voxel build and the association stream 16 B per point and the map build needs no min/max
 * pass.  The 32 B pcl::PointXYZI layout exists only at the boundary (import / export below). */
/* ScanUndistortion::undistortScan / undistortScanInMap (same arguments as lvi_undistort_d), result kept as a batch */
int lvi_scan_batch_undistort_d(lvi_ctx* ctx, const lvi_problem_desc* traj, const lvi_point_xyzit* scans_raw_d, int32_t n_scans,
                               int64_t pts_per_scan, const double* target_time, int correct_position, lvi_scan_batch** out,
                               int32_t* n_bad_targets);
/* pcl::transformPointCloud with one float 4x4 per scan (same as lvi_transform_scans_d), batch in -> new batch out */
int lvi_scan_batch_transform(lvi_ctx* ctx, const lvi_scan_batch* in, const double* poses, lvi_scan_batch** out);
/* import a device-resident PCL cloud (x,y,z float at offset 0, intensity at offset 16 when stride_bytes >= 20) as a batch */
int lvi_scan_batch_from_xyzi_d(lvi_ctx* ctx, const void* xyzi_d, size_t stride_bytes, int32_t n_scans, int64_t pts_per_scan,
                               lvi_scan_batch** out);
/* write the batch as pcl::PointXYZI records (32 B: x,y,z,1,intensity,0,0,0) to a host (out_is_device = 0) or device buffer */
int lvi_scan_batch_export_xyzi(lvi_ctx* ctx, const lvi_scan_batch* b, void* out_xyzi, int out_is_device);
int lvi_scan_batch_destroy(lvi_scan_batch* b);
int64_t lvi_scan_batch_num_points(const lvi_scan_batch* b);
int32_t lvi_scan_batch_num_scans(const lvi_scan_batch* b);
const void* lvi_scan_batch_points_d(const lvi_scan_batch* b); /* packed device buffer: float x,y,z,intensity per point */
/* lvi_voxel_build over the scans of a batch.  scan_keep (host, [n_scans], may be NULL = all) selects the scans that make up the map
 * cloud: LiDAROdometry::feedScan only adds KEY scans to it (L/src/core/lidar_odometry.cpp:89-128).  Point indices reported by
 * lvi_voxel_export count the points of the selected scans in batch order (= the concatenated map_cloud_). */
int lvi_voxel_build_batch(lvi_ctx* ctx, const lvi_scan_batch* batch, const uint8_t* scan_keep, float leaf_size, int min_points,
                          double eig_mult, lvi_voxel_map** out);
/* lvi_associate_d with the map-frame scans given as a batch (W x H must equal the batch's pts_per_scan) */
int lvi_associate_batch(lvi_ctx* ctx, const lvi_voxel_map* m, const lvi_surfel_set* s, const lvi_scan_batch* batch,
                        const lvi_point_xyzit* scans_raw_d, int32_t W, int32_t H, double radius, int32_t k_per_ring,
                        int32_t time_step, lvi_surfel_point* out_d, int64_t cap, int64_t* n_out, int64_t* n_all);

/* ---- (e) the map path sharded over the GPUs of a node --------------------------------------------------------------------- */
/* SURVEY §8e collectives 1-2.  Every rank passes the map-frame scans of ITS time chunk (ranks = consecutive time chunks).  The voxel grid
 * (min_b_ / div_b_) is that of the whole cloud (ncclAllReduce min / max), every leaf is built by the rank that owns its voxel-index range
 * from ALL its points in cloud order (all-to-all of points), and the surfel planes of all ranks are gathered on every rank in ascending
 * voxel-index order, so plane ids, planes and boxes equal those of lvi_voxel_build_batch + lvi_surfel_extract over the concatenated scans.
 *   *out_map     : look-up map over the surfel leaves (what lvi_associate_* needs); lvi_voxel_export is not available on it
 *   *out_surfels : all planes (lvi_surfel_count / lvi_surfel_export / lvi_associate_landmarks work as usual)
 *   stats[4]     : (may be NULL) points sent to other ranks, points received, leaves and planes built by this rank
 * With world == 1: lvi_voxel_build_batch + lvi_surfel_extract. */
int lvi_map_build_sharded(lvi_ctx* ctx, const lvi_scan_batch* local_scans, const uint8_t* scan_keep, float leaf_size, int min_points,
                          double eig_mult, double lambda, int min_leaf_points, float ransac_threshold, int min_inliers,
                          lvi_voxel_map** out_map, lvi_surfel_set** out_surfels, int64_t* stats);
/* lvi_associate_batch over every rank's own scans.  The every-`time_step`-th decimation runs over the associated points of ALL ranks in
 * time order (ncclAllGather of the hit counts); the selected points of all ranks are gathered into out_d on every rank (capacity cap;
 * out_d == NULL sizes: *n_out = selected points of all ranks, *n_all = associated points of all ranks). */
int lvi_associate_sharded(lvi_ctx* ctx, const lvi_voxel_map* m, const lvi_surfel_set* s, const lvi_scan_batch* local_scans,
                          const lvi_point_xyzit* scans_raw_d, int32_t W, int32_t H, double radius, int32_t k_per_ring, int32_t time_step,
                          lvi_surfel_point* out_d, int64_t cap, int64_t* n_out, int64_t* n_all);

/* ---- diagnostics ---------------------------------------------------------------------------------------------- */
/* Solves A x = rhs through the band+arrow tile Cholesky used by lvi_problem_solve (tests only). A_dense is
 * [n x n] row-major symmetric positive definite, n = nb + nbo; within the first nb rows/cols entries with
 * |i-j| > bw must be zero.  Two-sided ordering: band positions [0, chain1_start) and [chain1_start, nb) are two
 * chains that must not couple directly (chain1_start a multiple of 32, == nb for one chain); the first n_mid
 * border dims are the separator (they may only couple to the last bw positions of each chain), factored as a
 * second-level system.  LVI_ERR_NUMERIC on Cholesky breakdown. */
int lvi_band_solve_dense(lvi_ctx* ctx, int nb, int nbo, int bw, int chain1_start, int n_mid, const double* A_dense,
                         const double* rhs, double* x_out);

#ifdef __cplusplus
}
#endif
#endif /* LVI_EXC_B200_H */
