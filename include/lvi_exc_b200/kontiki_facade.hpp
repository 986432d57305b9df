// kontiki_facade.hpp — C++ host facade with the reference's own names over the C-ABI (include/lvi_exc_b200.h).
//
// The reference's boundary for the least-squares half of the hot path is the Kontiki template API that
// TrajectoryManagerLVI drives.  This header gives those names the same meaning on top of the CUDA library:
//   kontiki::trajectories::SplitTrajectory          K/trajectories/split_trajectory.h:87-131, spline_base.h:370-378
//   kontiki::sensors::{ConstantBiasImu, VLP16LiDAR, PinholeCamera}   K/sensors/sensors.h:36-167, imu.h:41-142,
//                                                   constant_bias_imu.h:33-119, vlp16_lidar.h:39-45, pinhole_camera.h:54-124
//   kontiki::sfm::{Landmark, View, Observation}      K/sfm/landmark.h:15-50, view.h:15-33, observation.h:14-36
//   kontiki::measurements::{GyroscopeMeasurement, AccelerometerMeasurement, LiDARSurfelPoint, StaticRsCameraMeasurement,
//                           CameraSurfelLandmark, OrientationMeasurement}      K/measurements/*.h (ctor argument order kept)
//   kontiki::TrajectoryEstimator<Traj>::{AddMeasurement, Solve}     K/trajectory_estimator.h:29-94
// AddMeasurement records the measurement into flat tables (instead of allocating a ceres CostFunction + parameter-pointer list);
// Solve lowers them to one lvi_problem_desc, runs lvi_problem_solve on the GPU and writes the optimum back into the trajectory /
// sensor / landmark objects in place, like Ceres updates Kontiki's DynamicParameterStore memory.  Errors map back to the exception
// types the reference throws.  No Eigen in this image: Vector3d / Quaterniond are minimal PODs with Eigen's storage order.
#pragma once
#include <array>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../lvi_exc_b200.h"

namespace Eigen_like {
struct Vector3d { double x = 0, y = 0, z = 0; Vector3d() = default; Vector3d(double a, double b, double c) : x(a), y(b), z(c) {} };
struct Vector2d { double x = 0, y = 0; Vector2d() = default; Vector2d(double a, double b) : x(a), y(b) {} };
struct Quaterniond {  // coeffs() order x, y, z, w
  double x = 0, y = 0, z = 0, w = 1;
  Quaterniond() = default;
  Quaterniond(double w_, double x_, double y_, double z_) : x(x_), y(y_), z(z_), w(w_) {}  // Eigen ctor order (w, x, y, z)
  static Quaterniond Identity() { return Quaterniond(); }
};
}  // namespace Eigen_like

namespace ceres_like {  // the slice of ceres::Solver::Summary the reference prints (BriefReport, e.g. trajectory_manager_lvi.cpp:341)
struct Summary {
  lvi_solve_summary raw{};
  int num_successful_steps = 0, num_unsuccessful_steps = 0;
  double initial_cost = 0, final_cost = 0, total_time_in_seconds = 0;
  bool IsSolutionUsable() const { return raw.termination_type != LVI_FAILURE; }
  std::string BriefReport() const {
    static const char* term[] = {"CONVERGENCE", "NO_CONVERGENCE", "FAILURE"};
    std::ostringstream s;
    s << "Ceres-shaped Solver Report (lvi_exc_b200): Iterations: " << raw.num_iterations << ", Initial cost: " << initial_cost
      << ", Final cost: " << final_cost << ", Termination: " << term[raw.termination_type];
    return s.str();
  }
};
}  // namespace ceres_like

namespace kontiki {
using Eigen_like::Quaterniond;
using Eigen_like::Vector2d;
using Eigen_like::Vector3d;

inline void throw_status(int rc) {
  if (rc == LVI_OK) return;
  const std::string msg = lvi_last_error();
  if (rc == LVI_ERR_RANGE) throw std::range_error(msg);     // K/trajectory_estimator.h:111-122, spline_base.h:221
  if (rc == LVI_ERR_DOMAIN) throw std::domain_error(msg);   // K/trajectories/uniform_so3_spline_trajectory.h:23-27
  throw std::runtime_error(msg);
}

namespace trajectories {
// SplitTrajectory(r3_dt, so3_dt, r3_t0, so3_t0); the hot path always uses equal dt / t0 (L/include/core/trajectory_manager_lvi.h:120-124)
class SplitTrajectory {
 public:
  SplitTrajectory(double r3_dt, double so3_dt, double r3_t0, double so3_t0) : dt_(r3_dt), t0_(r3_t0) {
    if (r3_dt != so3_dt || r3_t0 != so3_t0) throw std::invalid_argument("lvi_exc_b200: SplitTrajectory needs equal dt and t0 for both splines");
  }
  double dt() const { return dt_; }
  double t0() const { return t0_; }
  size_t NumKnots() const { return r3_.size() / 3; }
  double MinTime() const { return t0_; }
  double MaxTime() const { return t0_ + (static_cast<double>(NumKnots()) - 3) * dt_; }  // spline_base.h:53-56
  void AppendKnot(const Vector3d& p, const Quaterniond& q) {
    const double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    if (std::fabs(n - 1.0) > 1e-5) throw std::domain_error("SO3 control points must be unit quaternions");
    r3_.insert(r3_.end(), {p.x, p.y, p.z});
    so3_.insert(so3_.end(), {q.x, q.y, q.z, q.w});
  }
  void ExtendTo(double t, const Vector3d& fill_p, const Quaterniond& fill_q) {  // spline_base.h:374-378
    while (NumKnots() < 4 || MaxTime() < t) AppendKnot(fill_p, fill_q);
  }
  Vector3d R3ControlPoint(size_t i) const { return Vector3d(r3_[3 * i], r3_[3 * i + 1], r3_[3 * i + 2]); }
  Quaterniond SO3ControlPoint(size_t i) const { Quaterniond q; q.x = so3_[4 * i]; q.y = so3_[4 * i + 1]; q.z = so3_[4 * i + 2]; q.w = so3_[4 * i + 3]; return q; }
  void Lock(bool lock) { locked_ = lock; }
  bool IsLocked() const { return locked_; }
  std::vector<double>& r3_data() { return r3_; }
  std::vector<double>& so3_data() { return so3_; }
  // Position / Orientation at time t (Trajectory::Position / Orientation, K/trajectories/trajectory.h:95-133) on the device
  void Evaluate(lvi_ctx* ctx, double t, Vector3d* p, Quaterniond* q) {
    lvi_problem_desc d{};
    d.t0 = t0_; d.dt = dt_; d.n_knots = static_cast<int32_t>(NumKnots()); d.r3_knots = r3_.data(); d.so3_knots = so3_.data();
    double pos[3], quat[4]; uint8_t valid = 0;
    throw_status(lvi_trajectory_evaluate(ctx, &d, &t, 1, pos, quat, &valid));
    if (!valid) throw std::range_error("t is out of range for trajectory");
    if (p) *p = Vector3d(pos[0], pos[1], pos[2]);
    if (q) { q->x = quat[0]; q->y = quat[1]; q->z = quat[2]; q->w = quat[3]; }
  }
 private:
  double dt_, t0_;
  bool locked_ = false;
  std::vector<double> r3_, so3_;
};
}  // namespace trajectories

namespace sensors {
// SensorEntity: relative pose + time offset, locked by default (K/sensors/sensors.h:93-167)
class Sensor {
 public:
  Quaterniond relative_orientation() const { Quaterniond q; q.x = q_[0]; q.y = q_[1]; q.z = q_[2]; q.w = q_[3]; return q; }
  void set_relative_orientation(const Quaterniond& q) { q_ = {q.x, q.y, q.z, q.w}; }
  Vector3d relative_position() const { return Vector3d(p_[0], p_[1], p_[2]); }
  void set_relative_position(const Vector3d& p) { p_ = {p.x, p.y, p.z}; }
  double time_offset() const { return toff_; }
  void set_time_offset(double t) { toff_ = t; }
  double max_time_offset() const { return max_toff_; }
  void set_max_time_offset(double m) { max_toff_ = m; }
  void LockRelativeOrientation(bool l) { lock_q_ = l; }
  void LockRelativePosition(bool l) { lock_p_ = l; }
  void LockTimeOffset(bool l) {
    if (!l) throw std::invalid_argument("lvi_exc_b200: time-offset optimisation is not built (SURVEY §8 f-4; cfg optimize_time_offset = false)");
  }
  bool RelativeOrientationIsLocked() const { return lock_q_; }
  bool RelativePositionIsLocked() const { return lock_p_; }
  std::array<double, 4> q_{{0, 0, 0, 1}};
  std::array<double, 3> p_{{0, 0, 0}};
 protected:
  double toff_ = 0, max_toff_ = 0;
  bool lock_q_ = true, lock_p_ = true;
};
class VLP16LiDAR : public Sensor {};
class ConstantBiasImu : public Sensor {  // K/sensors/constant_bias_imu.h:33-119, imu.h:41-70,127
 public:
  Vector3d accelerometer_bias() const { return Vector3d(ba_[0], ba_[1], ba_[2]); }
  void set_accelerometer_bias(const Vector3d& b) { ba_ = {b.x, b.y, b.z}; }
  Vector3d gyroscope_bias() const { return Vector3d(bg_[0], bg_[1], bg_[2]); }
  void set_gyroscope_bias(const Vector3d& b) { bg_ = {b.x, b.y, b.z}; }
  void LockAccelerometerBias(bool l) { lock_ba_ = l; }
  void LockGyroscopeBias(bool l) { lock_bg_ = l; }
  bool AccelerometerBiasIsLocked() const { return lock_ba_; }
  bool GyroscopeBiasIsLocked() const { return lock_bg_; }
  double gravity_orientation_roll() const { return g_[0]; }
  double gravity_orientation_pitch() const { return g_[1]; }
  void set_gravity_orientation_roll(double r) { g_[0] = r; }
  void set_gravity_orientation_pitch(double p) { g_[1] = p; }
  Vector3d refined_gravity() const {  // imu.h:61-70, G = -9.79 (imu.h:25)
    const double G = -9.79, cr = std::cos(g_[0]), sr = std::sin(g_[0]), cp = std::cos(g_[1]), sp = std::sin(g_[1]);
    return Vector3d(-sp * cr * G, sr * G, -cr * cp * G);
  }
  std::array<double, 3> ba_{{0, 0, 0}}, bg_{{0, 0, 0}};
  std::array<double, 2> g_{{0.01, 0.01}};
 private:
  bool lock_ba_ = true, lock_bg_ = true;  // constant_bias_imu.h:70-71
};
class PinholeCamera : public Sensor {  // PinholeCamera(rows, cols, readout, k1,k2,p1,p2,k3, fx,fy,cx,cy), pinhole_camera.h:253-261
 public:
  PinholeCamera(size_t rows, size_t cols, double readout, double k1, double k2, double p1, double p2, double k3, double fx, double fy, double cx, double cy)
      : rows_(rows), cols_(cols), readout_(readout), fx_(fx), fy_(fy), cx_(cx), cy_(cy), k_{{k1, k2, p1, p2, k3}} {}
  const std::array<double, 5>& distortion_params() const { return k_; }   // applied by the device residuals (lvi_problem_desc.distortion)
  size_t rows() const { return rows_; }
  size_t cols() const { return cols_; }
  double readout() const { return readout_; }
  double fx() const { return fx_; } double fy() const { return fy_; } double cx() const { return cx_; } double cy() const { return cy_; }
 private:
  size_t rows_, cols_;
  double readout_, fx_, fy_, cx_, cy_;
  std::array<double, 5> k_;
};
}  // namespace sensors

namespace sfm {
class Landmark;
class View {
 public:
  View(size_t frame, double t0) : frame_nr_(frame), t0_(t0) {}
  double t0() const { return t0_; }
  size_t frame_nr() const { return frame_nr_; }
 private:
  size_t frame_nr_; double t0_;
};
class Observation {
 public:
  Observation(const Vector2d& uv, std::shared_ptr<Landmark> lm, std::shared_ptr<View> v) : uv_(uv), landmark_(lm), view_(v) {}
  Vector2d uv() const { return uv_; }
  std::shared_ptr<Landmark> landmark() const { return landmark_; }
  std::shared_ptr<View> view() const { return view_; }
 private:
  Vector2d uv_; std::shared_ptr<Landmark> landmark_; std::shared_ptr<View> view_;
};
class Landmark {  // K/sfm/landmark.h:15-50
 public:
  void set_reference(std::shared_ptr<Observation> o) { reference_ = o; }
  std::shared_ptr<Observation> reference() const { return reference_; }
  double inverse_depth() const { return rho_; }
  void set_inverse_depth(double r) { rho_ = r; }
  double* inverse_depth_ptr() { return &rho_; }
  void Lock(bool l) { locked_ = l; }
  bool IsLocked() const { return locked_; }
  std::vector<std::weak_ptr<Observation>>& observations() { return obs_; }
 private:
  double rho_ = 0; bool locked_ = false;
  std::shared_ptr<Observation> reference_;
  std::vector<std::weak_ptr<Observation>> obs_;
};
}  // namespace sfm

namespace measurements {
template <class ImuModel> struct GyroscopeMeasurement {  // (imu, t, w, weight)  gyroscope_measurement.h:19-48
  GyroscopeMeasurement(std::shared_ptr<ImuModel> imu_, double t_, const Vector3d& w_, double weight_ = 1.0) : imu(imu_), t(t_), w(w_), weight(weight_) {}
  std::shared_ptr<ImuModel> imu; double t; Vector3d w; double weight;
};
template <class ImuModel> struct AccelerometerMeasurement {  // (imu, t, a, weight)  accelerometer_measurement.h:20-49
  AccelerometerMeasurement(std::shared_ptr<ImuModel> imu_, double t_, const Vector3d& a_, double weight_ = 1.0) : imu(imu_), t(t_), a(a_), weight(weight_) {}
  std::shared_ptr<ImuModel> imu; double t; Vector3d a; double weight;
};
template <class LiDARModel> struct LiDARSurfelPoint {  // (lidar, point, plane, t, map_time, huber = 5, weight = 1)  lidar_surfel_point.h:18-27
  LiDARSurfelPoint(std::shared_ptr<LiDARModel> lidar_, const Vector3d& point_, double* plane_, double t_, double map_time_, double huber_ = 5.0, double weight_ = 1.0)
      : lidar(lidar_), point(point_), plane(plane_), t(t_), map_time(map_time_), huber(huber_), weight(weight_) {}
  std::shared_ptr<LiDARModel> lidar; Vector3d point; double* plane; double t, map_time, huber, weight;
};
template <class CameraModel> struct StaticRsCameraMeasurement {  // (camera, obs, huber = 5, weight = 1)  static_rscamera_measurement.h:67-74
  StaticRsCameraMeasurement(std::shared_ptr<CameraModel> camera_, std::shared_ptr<sfm::Observation> obs_, double huber_ = 5.0, double weight_ = 1.0)
      : camera(camera_), observation(obs_), huber(huber_), weight(weight_) {}
  std::shared_ptr<CameraModel> camera; std::shared_ptr<sfm::Observation> observation; double huber, weight;
};
template <class CameraModel, class LiDARModel> struct CameraSurfelLandmark {  // camera_surfel_landmark.h:19-26
  CameraSurfelLandmark(std::shared_ptr<CameraModel> camera_, std::shared_ptr<LiDARModel> lidar_, sfm::Landmark* lm_, double* plane_, double t_, double map_time_,
                       double huber_ = 5.0, double weight_ = 1.0)
      : camera(camera_), lidar(lidar_), landmark(lm_), plane(plane_), t(t_), map_time(map_time_), huber(huber_), weight(weight_) {}
  std::shared_ptr<CameraModel> camera; std::shared_ptr<LiDARModel> lidar; sfm::Landmark* landmark; double* plane; double t, map_time, huber, weight;
};
struct OrientationMeasurement {  // (t, q, weight)  orientation_measurement.h:20-22
  OrientationMeasurement(double t_, const Quaterniond& q_, double weight_ = 1.0) : t(t_), q(q_), weight(weight_) {}
  double t; Quaterniond q; double weight;
};
}  // namespace measurements

// kontiki::TrajectoryEstimator<SplitTrajectory> (K/trajectory_estimator.h:19-135).  `so3_only` reproduces
// TrajectoryEstimator<UniformSO3SplineTrajectory> of initialSO3TrajWithGyro (L/src/core/trajectory_manager_lvi.cpp:43-62).
template <class TrajectoryModel = trajectories::SplitTrajectory>
class TrajectoryEstimator {
 public:
  TrajectoryEstimator(lvi_ctx* ctx, std::shared_ptr<TrajectoryModel> trajectory, bool so3_only = false) : ctx_(ctx), traj_(trajectory), so3_only_(so3_only) {}
  std::shared_ptr<TrajectoryModel> trajectory() const { return traj_; }

  template <class I> void AddMeasurement(std::shared_ptr<measurements::GyroscopeMeasurement<I>> m) {
    imu_ = m->imu; gyro_t_.push_back(m->t); push3(gyro_w_, m->w); gyro_wt_.push_back(m->weight); keep_.push_back(m);
  }
  template <class I> void AddMeasurement(std::shared_ptr<measurements::AccelerometerMeasurement<I>> m) {
    imu_ = m->imu; acc_t_.push_back(m->t); push3(acc_a_, m->a); acc_wt_.push_back(m->weight); keep_.push_back(m);
  }
  template <class Li> void AddMeasurement(std::shared_ptr<measurements::LiDARSurfelPoint<Li>> m) {
    lidar_ = m->lidar; sf_t_.push_back(m->t); sf_tm_.push_back(m->map_time); push3(sf_p_, m->point); sf_plane_.push_back(plane_id(m->plane));
    sf_wt_.push_back(m->weight); sf_hb_.push_back(m->huber); keep_.push_back(m);
  }
  template <class C> void AddMeasurement(std::shared_ptr<measurements::StaticRsCameraMeasurement<C>> m) {
    cam_ = m->camera;
    auto lm = m->observation->landmark();
    auto ref = lm->reference();
    if (!ref) throw std::runtime_error("landmark has no reference observation");
    cam_tr_.push_back(ref->view()->t0()); cam_to_.push_back(m->observation->view()->t0());
    cam_uvr_.push_back(ref->uv().x); cam_uvr_.push_back(ref->uv().y); cam_uvo_.push_back(m->observation->uv().x); cam_uvo_.push_back(m->observation->uv().y);
    cam_lm_.push_back(landmark_id(lm.get())); cam_wt_.push_back(m->weight); cam_hb_.push_back(m->huber); keep_.push_back(m);
  }
  template <class C, class Li> void AddMeasurement(std::shared_ptr<measurements::CameraSurfelLandmark<C, Li>> m) {
    cam_ = m->camera; lidar_ = m->lidar;
    auto ref = m->landmark->reference();
    if (!ref) throw std::runtime_error("landmark has no reference observation");
    cs_t_.push_back(m->t); cs_tm_.push_back(m->map_time); cs_uv_.push_back(ref->uv().x); cs_uv_.push_back(ref->uv().y);
    cs_lm_.push_back(landmark_id(m->landmark)); cs_plane_.push_back(plane_id(m->plane)); cs_wt_.push_back(m->weight); cs_hb_.push_back(m->huber); keep_.push_back(m);
  }
  void AddMeasurement(std::shared_ptr<measurements::OrientationMeasurement> m) {
    or_t_.push_back(m->t); or_q_.insert(or_q_.end(), {m->q.x, m->q.y, m->q.z, m->q.w}); or_wt_.push_back(m->weight); keep_.push_back(m);
  }

  size_t num_residual_blocks() const { return gyro_t_.size() + acc_t_.size() + sf_t_.size() + cam_tr_.size() + cs_t_.size() + or_t_.size(); }

  // the flat problem the C-ABI takes; pointers stay valid until the next AddMeasurement
  lvi_problem_desc Describe() {
    lvi_problem_desc d{};
    d.t0 = traj_->t0(); d.dt = traj_->dt(); d.n_knots = static_cast<int32_t>(traj_->NumKnots());
    d.r3_knots = so3_only_ ? nullptr : traj_->r3_data().data(); d.so3_knots = traj_->so3_data().data();
    d.lock_r3 = d.lock_so3 = traj_->IsLocked() ? 1 : 0;
    d.lock_lidar_q = d.lock_lidar_p = d.lock_cam_q = d.lock_cam_p = d.lock_acc_bias = d.lock_gyr_bias = 1;
    if (lidar_) { d.lidar_q = lidar_->q_.data(); d.lidar_p = lidar_->p_.data(); d.lidar_toff = lidar_->time_offset();
                  d.lock_lidar_q = lidar_->RelativeOrientationIsLocked(); d.lock_lidar_p = lidar_->RelativePositionIsLocked(); }
    if (cam_) { d.cam_q = cam_->q_.data(); d.cam_p = cam_->p_.data(); d.cam_toff = cam_->time_offset();
                d.lock_cam_q = cam_->RelativeOrientationIsLocked(); d.lock_cam_p = cam_->RelativePositionIsLocked();
                d.fx = cam_->fx(); d.fy = cam_->fy(); d.cx = cam_->cx(); d.cy = cam_->cy(); d.readout = cam_->readout();
                for (int q = 0; q < 5; ++q) d.distortion[q] = cam_->distortion_params()[q];
                d.cam_rows = static_cast<int32_t>(cam_->rows()); d.cam_cols = static_cast<int32_t>(cam_->cols()); }
    if (imu_) { d.gravity = imu_->g_.data(); d.acc_bias = imu_->ba_.data(); d.gyr_bias = imu_->bg_.data(); d.imu_toff = imu_->time_offset();
                d.lock_acc_bias = imu_->AccelerometerBiasIsLocked(); d.lock_gyr_bias = imu_->GyroscopeBiasIsLocked(); }
    else { d.gravity = dummy_g_; d.acc_bias = dummy_b_; d.gyr_bias = dummy_b_ + 3; }
    planes_flat_.resize(3 * planes_.size());
    for (size_t k = 0; k < planes_.size(); ++k) std::memcpy(&planes_flat_[3 * k], planes_[k], 24);
    d.n_planes = static_cast<int32_t>(planes_.size()); d.planes = planes_flat_.data();
    rho_.resize(landmarks_.size()); rho_locked_.resize(landmarks_.size());
    for (size_t l = 0; l < landmarks_.size(); ++l) { rho_[l] = landmarks_[l]->inverse_depth(); rho_locked_[l] = landmarks_[l]->IsLocked(); }
    d.n_landmarks = static_cast<int32_t>(landmarks_.size()); d.rho = rho_.data(); d.rho_locked = rho_locked_.data();
    d.n_gyro = static_cast<int32_t>(gyro_t_.size()); d.gyro_t = gyro_t_.data(); d.gyro_w = gyro_w_.data(); d.gyro_weight = gyro_wt_.data();
    d.n_accel = static_cast<int32_t>(acc_t_.size()); d.accel_t = acc_t_.data(); d.accel_a = acc_a_.data(); d.accel_weight = acc_wt_.data();
    d.n_surfel = static_cast<int32_t>(sf_t_.size()); d.surfel_t = sf_t_.data(); d.surfel_tmap = sf_tm_.data(); d.surfel_point = sf_p_.data();
    d.surfel_plane = sf_plane_.data(); d.surfel_weight = sf_wt_.data(); d.surfel_huber = sf_hb_.data();
    d.n_cam = static_cast<int32_t>(cam_tr_.size()); d.cam_t0_ref = cam_tr_.data(); d.cam_t0_obs = cam_to_.data(); d.cam_uv_ref = cam_uvr_.data();
    d.cam_uv_obs = cam_uvo_.data(); d.cam_landmark = cam_lm_.data(); d.cam_weight = cam_wt_.data(); d.cam_huber = cam_hb_.data();
    d.n_camsurf = static_cast<int32_t>(cs_t_.size()); d.cs_t = cs_t_.data(); d.cs_tmap = cs_tm_.data(); d.cs_uv = cs_uv_.data();
    d.cs_landmark = cs_lm_.data(); d.cs_plane = cs_plane_.data(); d.cs_weight = cs_wt_.data(); d.cs_huber = cs_hb_.data();
    d.n_orient = static_cast<int32_t>(or_t_.size()); d.orient_t = or_t_.data(); d.orient_q = or_q_.data(); d.orient_weight = or_wt_.data();
    return d;
  }

  // Solve(max_iterations = 30, progress = true, num_threads = -1): num_threads is meaningless on the device and ignored
  ceres_like::Summary Solve(int max_iterations = 30, bool progress = true, int /*num_threads*/ = -1) {
    lvi_problem_desc d = Describe();
    lvi_problem* p = nullptr;
    throw_status(lvi_problem_create(ctx_, &d, &p));
    lvi_solve_options o;
    lvi_solve_options_default(&o);
    o.max_num_iterations = max_iterations; o.verbose = progress ? 1 : 0;
    ceres_like::Summary s;
    const int rc = lvi_problem_solve(p, &o, &s.raw);
    lvi_problem_destroy(p);
    throw_status(rc);
    for (size_t l = 0; l < landmarks_.size(); ++l) landmarks_[l]->set_inverse_depth(rho_[l]);  // everything else was updated in place
    s.initial_cost = s.raw.initial_cost; s.final_cost = s.raw.final_cost; s.total_time_in_seconds = s.raw.time_total_ms * 1e-3;
    s.num_successful_steps = s.raw.num_successful_steps; s.num_unsuccessful_steps = s.raw.num_unsuccessful_steps;
    return s;
  }

 private:
  static void push3(std::vector<double>& v, const Vector3d& a) { v.insert(v.end(), {a.x, a.y, a.z}); }
  int32_t plane_id(double* plane) {  // plane pointers alias closest_point_vec_ elements (trajectory_manager_lvi.cpp:566-578)
    auto it = plane_ids_.find(plane);
    if (it != plane_ids_.end()) return it->second;
    const int32_t id = static_cast<int32_t>(planes_.size());
    plane_ids_[plane] = id; planes_.push_back(plane);
    return id;
  }
  int32_t landmark_id(sfm::Landmark* lm) {
    auto it = landmark_ids_.find(lm);
    if (it != landmark_ids_.end()) return it->second;
    const int32_t id = static_cast<int32_t>(landmarks_.size());
    landmark_ids_[lm] = id; landmarks_.push_back(lm);
    return id;
  }
  lvi_ctx* ctx_;
  std::shared_ptr<TrajectoryModel> traj_;
  bool so3_only_;
  std::shared_ptr<sensors::ConstantBiasImu> imu_;
  std::shared_ptr<sensors::VLP16LiDAR> lidar_;
  std::shared_ptr<sensors::PinholeCamera> cam_;
  std::vector<std::shared_ptr<void>> keep_;  // measurements must outlive the estimator in the reference; here they are simply retained
  std::map<double*, int32_t> plane_ids_; std::vector<double*> planes_; std::vector<double> planes_flat_;
  std::map<sfm::Landmark*, int32_t> landmark_ids_; std::vector<sfm::Landmark*> landmarks_; std::vector<double> rho_; std::vector<uint8_t> rho_locked_;
  std::vector<double> gyro_t_, gyro_w_, gyro_wt_, acc_t_, acc_a_, acc_wt_, sf_t_, sf_tm_, sf_p_, sf_wt_, sf_hb_;
  std::vector<int32_t> sf_plane_, cam_lm_, cs_lm_, cs_plane_;
  std::vector<double> cam_tr_, cam_to_, cam_uvr_, cam_uvo_, cam_wt_, cam_hb_, cs_t_, cs_tm_, cs_uv_, cs_wt_, cs_hb_, or_t_, or_q_, or_wt_;
  double dummy_g_[2] = {0.01, 0.01}, dummy_b_[6] = {0, 0, 0, 0, 0, 0};
};

}  // namespace kontiki
