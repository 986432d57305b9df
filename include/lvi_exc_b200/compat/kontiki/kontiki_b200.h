// kontiki_b200.h — the Kontiki surface of the calibration hot path, signature for signature, over the C-ABI of the CUDA
// library (include/lvi_exc_b200.h).  The headers next to this one carry the reference's own include paths and forward here.
//
//   kontiki::trajectories::{UniformR3SplineTrajectory, UniformSO3SplineTrajectory, SplitTrajectory, TrajectoryEvaluation, EvaluationFlags}
//        K/kontiki/trajectories/{trajectory.h:17-133, spline_base.h:340-378, uniform_r3_spline_trajectory.h, uniform_so3_spline_trajectory.h,
//        split_trajectory.h:87-131}
//   kontiki::sensors::{ConstantBiasImu, VLP16LiDAR, PinholeCamera}        K/kontiki/sensors/{sensors.h:36-167, imu.h:41-142,
//        constant_bias_imu.h:33-119, vlp16_lidar.h:39-45, pinhole_camera.h:54-124,253-261}
//   kontiki::sfm::{Landmark, View, Observation}                            K/kontiki/sfm/{landmark.h:15-50, view.h:15-33, observation.h:14-36}
//   kontiki::measurements::{GyroscopeMeasurement, AccelerometerMeasurement, LiDARSurfelPoint, StaticRsCameraMeasurement,
//        CameraSurfelLandmark, OrientationMeasurement}                     K/kontiki/measurements/*.h (constructor argument order kept)
//   kontiki::TrajectoryEstimator<Traj>::{TrajectoryEstimator(shared_ptr<Traj>), trajectory, AddMeasurement<M>, Solve, problem, AddCallback}
//        K/kontiki/trajectory_estimator.h:19-135
//
// How it differs inside: AddMeasurement<M> does not allocate a ceres cost function and a parameter-pointer list per measurement; it
// appends one row to the flat table of its type (and registers the blocks it touches with the estimator's ceres::Problem, which is
// bookkeeping only).  Solve lowers the tables to one lvi_problem_desc, runs the device LM (lvi_problem_solve) and the optimum lands in the
// SAME parameter memory the objects expose through Eigen::Map, as Ceres updates Kontiki's DynamicParameterStore in place.  Evaluations
// of single poses go through lvi_trajectory_evaluate_full (device).  There is no CPU path: without a CUDA device the first call that
// needs the context throws.
#ifndef LVI_EXC_B200_COMPAT_KONTIKI_H
#define LVI_EXC_B200_COMPAT_KONTIKI_H
#include <array>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include <Eigen/Dense>
#include <ceres/ceres.h>

#include "../../../lvi_exc_b200.h"

#include "../context.h"

namespace kontiki {

namespace trajectories {

enum EvaluationFlags { EvalPosition = 1, EvalVelocity = 2, EvalAcceleration = 4, EvalOrientation = 8, EvalAngularVelocity = 16 };

template <typename T> struct TrajectoryEvaluation {
  explicit TrajectoryEvaluation(int flags) : needs(flags) {}
  using Vector3 = Eigen::Vector3d;
  Vector3 position, velocity, acceleration;
  Eigen::Quaterniond orientation;
  Vector3 angular_velocity;
  struct Needs {
    explicit Needs(int f) : flags(f) {}
    bool Position() const { return flags & EvalPosition; }
    bool Velocity() const { return flags & EvalVelocity; }
    bool Acceleration() const { return flags & EvalAcceleration; }
    bool Orientation() const { return flags & EvalOrientation; }
    bool AngularVelocity() const { return flags & EvalAngularVelocity; }
    bool AnyLinear() const { return FlagsLinear(); }
    bool AnyRotation() const { return FlagsRotation(); }
    int FlagsLinear() const { return flags & (EvalPosition | EvalVelocity | EvalAcceleration); }
    int FlagsRotation() const { return flags & (EvalOrientation | EvalAngularVelocity); }
   protected:
    int flags;
  } needs;
};

namespace detail {
// one device evaluation (lvi_trajectory_evaluate_full) of the splines given as raw control-point arrays
inline std::unique_ptr<TrajectoryEvaluation<double>> evaluate(double t0, double dt, size_t n, const double* r3, const double* so3, double t, int flags) {
  lvi_problem_desc d{};
  std::vector<double> ident;
  d.t0 = t0; d.dt = dt; d.n_knots = static_cast<int32_t>(n);
  d.r3_knots = const_cast<double*>(r3);
  if (!so3) { ident.assign(4 * n, 0.0); for (size_t i = 0; i < n; ++i) ident[4 * i + 3] = 1.0; so3 = ident.data(); }
  d.so3_knots = const_cast<double*>(so3);
  double p[3], v[3], a[3], q[4], w[3];
  uint8_t valid = 0;
  lvi_exc_b200::throw_status(lvi_trajectory_evaluate_full(lvi_exc_b200::DefaultContext(), &d, &t, 1, p, v, a, q, w, &valid));
  if (!valid) throw std::range_error("t is out of range for spline");   // spline_base.h:221
  auto r = std::make_unique<TrajectoryEvaluation<double>>(flags);
  r->position = Eigen::Vector3d(p[0], p[1], p[2]); r->velocity = Eigen::Vector3d(v[0], v[1], v[2]); r->acceleration = Eigen::Vector3d(a[0], a[1], a[2]);
  r->orientation = Eigen::Quaterniond(q[3], q[0], q[1], q[2]); r->angular_velocity = Eigen::Vector3d(w[0], w[1], w[2]);
  return r;
}
}  // namespace detail

// uniform cubic B-spline with control points of DIM doubles each (3: R3, 4: unit quaternions x,y,z,w)
template <int DIM> class UniformSplineBase {
 public:
  using Result = std::unique_ptr<TrajectoryEvaluation<double>>;
  explicit UniformSplineBase(double dt = 1.0, double t0 = 0.0) : dt_(dt), t0_(t0) {}
  double dt() const { return dt_; }
  double t0() const { return t0_; }
  size_t NumKnots() const { return cps_.size() / DIM; }
  double MinTime() const { if (NumKnots() < 4) throw std::range_error("Spline had too few control points"); return t0_; }
  double MaxTime() const { if (NumKnots() < 4) throw std::range_error("Spline had too few control points"); return t0_ + (static_cast<double>(NumKnots()) - 3) * dt_; }
  std::pair<double, double> ValidTime() const { return std::make_pair(MinTime(), MaxTime()); }
  void Lock(bool lock) { locked_ = lock; }
  bool IsLocked() const { return locked_; }
  std::vector<double>& data() { return cps_; }
  const std::vector<double>& data() const { return cps_; }
 protected:
  double dt_, t0_;
  bool locked_ = false;
  std::vector<double> cps_;
};

class UniformR3SplineTrajectory : public UniformSplineBase<3> {
 public:
  using UniformSplineBase<3>::UniformSplineBase;
  using ControlPointType = Eigen::Vector3d;
  void AppendKnot(const Eigen::Vector3d& cp) { cps_.insert(cps_.end(), {cp(0), cp(1), cp(2)}); }
  void ExtendTo(double t, const Eigen::Vector3d& fill_value) { while (NumKnots() < 4 || MaxTime() < t) AppendKnot(fill_value); }   // spline_base.h:374-378
  Eigen::Map<Eigen::Vector3d> MutableControlPoint(size_t i) { return Eigen::Map<Eigen::Vector3d>(&cps_.at(3 * i)); }
  Eigen::Vector3d ControlPoint(size_t i) const { return Eigen::Vector3d(cps_.at(3 * i), cps_.at(3 * i + 1), cps_.at(3 * i + 2)); }
  Result Evaluate(double t, int flags) const { return detail::evaluate(t0_, dt_, NumKnots(), cps_.data(), nullptr, t, flags); }
  Eigen::Vector3d Position(double t) const { return Evaluate(t, EvalPosition)->position; }
  Eigen::Vector3d Velocity(double t) const { return Evaluate(t, EvalVelocity)->velocity; }
  Eigen::Vector3d Acceleration(double t) const { return Evaluate(t, EvalAcceleration)->acceleration; }
};

class UniformSO3SplineTrajectory : public UniformSplineBase<4> {
 public:
  using UniformSplineBase<4>::UniformSplineBase;
  using ControlPointType = Eigen::Quaterniond;
  void AppendKnot(const Eigen::Quaterniond& cp) {
    if (std::fabs(cp.norm() - 1.0) > 1e-5) throw std::domain_error("Control points must be unit quaternions!");   // uniform_so3_spline_trajectory.h:23-27
    cps_.insert(cps_.end(), {cp.x(), cp.y(), cp.z(), cp.w()});
  }
  void ExtendTo(double t, const Eigen::Quaterniond& fill_value) { while (NumKnots() < 4 || MaxTime() < t) AppendKnot(fill_value); }
  Eigen::Map<Eigen::Quaterniond> MutableControlPoint(size_t i) { return Eigen::Map<Eigen::Quaterniond>(&cps_.at(4 * i)); }
  Eigen::Quaterniond ControlPoint(size_t i) const { return Eigen::Quaterniond(cps_.at(4 * i + 3), cps_.at(4 * i), cps_.at(4 * i + 1), cps_.at(4 * i + 2)); }
  Result Evaluate(double t, int flags) const { return detail::evaluate(t0_, dt_, NumKnots(), nullptr, cps_.data(), t, flags); }
  Eigen::Quaterniond Orientation(double t) const { return Evaluate(t, EvalOrientation)->orientation; }
  Eigen::Vector3d AngularVelocity(double t) const { return Evaluate(t, EvalAngularVelocity)->angular_velocity; }
};

class SplitTrajectory {
 public:
  using Result = std::unique_ptr<TrajectoryEvaluation<double>>;
  SplitTrajectory(std::shared_ptr<UniformR3SplineTrajectory> r3, std::shared_ptr<UniformSO3SplineTrajectory> so3) : r3_(std::move(r3)), so3_(std::move(so3)) {}
  SplitTrajectory(double r3_dt, double so3_dt, double r3_t0, double so3_t0)
      : SplitTrajectory(std::make_shared<UniformR3SplineTrajectory>(r3_dt, r3_t0), std::make_shared<UniformSO3SplineTrajectory>(so3_dt, so3_t0)) {}
  SplitTrajectory(double r3_dt, double so3_dt) : SplitTrajectory(r3_dt, so3_dt, 0.0, 0.0) {}
  SplitTrajectory() : SplitTrajectory(1.0, 1.0) {}
  SplitTrajectory(const SplitTrajectory& rhs)
      : SplitTrajectory(std::make_shared<UniformR3SplineTrajectory>(*rhs.r3_), std::make_shared<UniformSO3SplineTrajectory>(*rhs.so3_)) {}
  std::shared_ptr<UniformR3SplineTrajectory> R3Spline() const { return r3_; }
  std::shared_ptr<UniformSO3SplineTrajectory> SO3Spline() const { return so3_; }
  bool IsLocked() const {
    if (r3_->IsLocked() != so3_->IsLocked()) throw std::runtime_error("R3 and SO3 trajectories have different lock status!");   // split_trajectory.h:101-109
    return r3_->IsLocked();
  }
  void Lock(bool lock) { r3_->Lock(lock); so3_->Lock(lock); }
  double MinTime() const { return std::max(r3_->MinTime(), so3_->MinTime()); }   // split_trajectory.h:60-65
  double MaxTime() const { return std::min(r3_->MaxTime(), so3_->MaxTime()); }
  std::pair<double, double> ValidTime() const { return std::make_pair(MinTime(), MaxTime()); }
  // SplitView::Evaluate (split_trajectory.h:41-58): linear part from the R3 spline, rotational part from the SO3 spline
  Result Evaluate(double t, int flags) const {
    require_common_grid();
    return detail::evaluate(r3_->t0(), r3_->dt(), std::min(r3_->NumKnots(), so3_->NumKnots()), r3_->data().data(), so3_->data().data(), t, flags);
  }
  Eigen::Vector3d Position(double t) const { return Evaluate(t, EvalPosition)->position; }
  Eigen::Vector3d Velocity(double t) const { return Evaluate(t, EvalVelocity)->velocity; }
  Eigen::Vector3d Acceleration(double t) const { return Evaluate(t, EvalAcceleration)->acceleration; }
  Eigen::Quaterniond Orientation(double t) const { return Evaluate(t, EvalOrientation)->orientation; }
  Eigen::Vector3d AngularVelocity(double t) const { return Evaluate(t, EvalAngularVelocity)->angular_velocity; }
  Eigen::Vector3d FromWorld(Eigen::Vector3d& Xw, double t) { Result r = Evaluate(t, EvalPosition | EvalOrientation); return r->orientation.conjugate() * (Xw - r->position); }
  Eigen::Vector3d ToWorld(Eigen::Vector3d& Xt, double t) { Result r = Evaluate(t, EvalPosition | EvalOrientation); return r->orientation * Xt + r->position; }
  // The device problem keeps both splines on ONE knot grid, as every caller on the hot path builds them
  // (L/include/core/trajectory_manager_lvi.h:120-124: SplitTrajectory(knot_distance, knot_distance, t0, t0))
  void require_common_grid() const {
    if (r3_->dt() != so3_->dt() || r3_->t0() != so3_->t0()) throw std::invalid_argument("lvi_exc_b200: SplitTrajectory needs equal dt and t0 for both splines");
  }
 private:
  std::shared_ptr<UniformR3SplineTrajectory> r3_;
  std::shared_ptr<UniformSO3SplineTrajectory> so3_;
};
}  // namespace trajectories

namespace sensors {
// SensorEntity (K/kontiki/sensors/sensors.h:93-167): parameter blocks [q_rel(4), p_rel(3), t_off(1)], locked by default
class SensorBase {
 public:
  SensorBase() { q_[3] = 1.0; }
  Eigen::Map<Eigen::Quaterniond> relative_orientation() const { return Eigen::Map<Eigen::Quaterniond>(const_cast<double*>(q_)); }
  void set_relative_orientation(const Eigen::Quaterniond& q) { std::memcpy(q_, q.c, 32); }
  Eigen::Map<Eigen::Vector3d> relative_position() const { return Eigen::Map<Eigen::Vector3d>(const_cast<double*>(p_)); }
  void set_relative_position(const Eigen::Vector3d& p) { p_[0] = p(0); p_[1] = p(1); p_[2] = p(2); }
  double& time_offset() const { return const_cast<double&>(toff_); }
  void set_time_offset(double d) {
    if (std::abs(d) <= max_time_offset()) { toff_ = d; return; }
    std::stringstream ss; ss << "Time offset |" << d << "| > " << max_time_offset();
    throw std::range_error(ss.str());
  }
  double max_time_offset() const { return max_toff_; }
  void set_max_time_offset(double m) { max_toff_ = m; }
  Eigen::Vector3d FromTrajectory(const Eigen::Vector3d& X) const { return relative_orientation() * X + Eigen::Vector3d(relative_position()); }
  Eigen::Vector3d ToTrajectory(const Eigen::Vector3d& X) const { return relative_orientation().conjugate() * (X - Eigen::Vector3d(relative_position())); }
  bool RelativeOrientationIsLocked() const { return lock_q_; }
  void LockRelativeOrientation(bool f) { lock_q_ = f; }
  bool RelativePositionIsLocked() const { return lock_p_; }
  void LockRelativePosition(bool f) { lock_p_ = f; }
  bool TimeOffsetIsLocked() const { return lock_t_; }
  // Time-offset optimisation (f-4, off in L/cfg/lvi.yaml:32) makes every evaluation time a variable; the device problem takes the offsets
  // as constants, so an unlocked offset is refused at Solve() instead of being silently ignored.
  void LockTimeOffset(bool f) { lock_t_ = f; }
  double* q_data() { return q_; }
  double* p_data() { return p_; }
 protected:
  double q_[4] = {0, 0, 0, 1}, p_[3] = {0, 0, 0};
  double toff_ = 0.0, max_toff_ = 0.01;   // sensors.h:112
  bool lock_q_ = true, lock_p_ = true, lock_t_ = true;
};
class VLP16LiDAR : public SensorBase {};

// ImuEntity + ConstantBiasImuEntity (imu.h:41-142, constant_bias_imu.h:33-119): [g_roll, g_pitch] never locked (Q5), biases locked by default
class ConstantBiasImu : public SensorBase {
 public:
  ConstantBiasImu() {}
  ConstantBiasImu(const Eigen::Vector3d& abias, const Eigen::Vector3d& gbias) { set_accelerometer_bias(abias); set_gyroscope_bias(gbias); }
  Eigen::Map<Eigen::Vector3d> accelerometer_bias() const { return Eigen::Map<Eigen::Vector3d>(const_cast<double*>(ba_)); }
  void set_accelerometer_bias(const Eigen::Vector3d& b) { ba_[0] = b(0); ba_[1] = b(1); ba_[2] = b(2); }
  Eigen::Map<Eigen::Vector3d> gyroscope_bias() const { return Eigen::Map<Eigen::Vector3d>(const_cast<double*>(bg_)); }
  void set_gyroscope_bias(const Eigen::Vector3d& b) { bg_[0] = b(0); bg_[1] = b(1); bg_[2] = b(2); }
  bool GyroscopeBiasIsLocked() const { return lock_bg_; }
  bool AccelerometerBiasIsLocked() const { return lock_ba_; }
  void LockGyroscopeBias(bool l) { lock_bg_ = l; }
  void LockAccelerometerBias(bool l) { lock_ba_ = l; }
  double& gravity_orientation_roll() const { return const_cast<double&>(g_[0]); }
  double& gravity_orientation_pitch() const { return const_cast<double&>(g_[1]); }
  void set_gravity_orientation_roll(double r) { g_[0] = r; }
  void set_gravity_orientation_pitch(double p) { g_[1] = p; }
  Eigen::Vector3d refined_gravity() const {   // imu.h:61-70 with G = -9.79 (imu.h:25)
    const double G = -9.79, cr = std::cos(g_[0]), sr = std::sin(g_[0]), cp = std::cos(g_[1]), sp = std::sin(g_[1]);
    return Eigen::Vector3d(-sp * cr * G, sr * G, -cr * cp * G);
  }
  double* ba_data() { return ba_; }
  double* bg_data() { return bg_; }
  double* g_data() { return g_; }
 private:
  double ba_[3] = {0, 0, 0}, bg_[3] = {0, 0, 0}, g_[2] = {0.01, 0.01};   // imu.h:127
  bool lock_ba_ = true, lock_bg_ = true;   // constant_bias_imu.h:70-71
};

// PinholeCamera(rows, cols, readout, k1, k2, p1, p2, k3, fx, fy, cx, cy) (pinhole_camera.h:253-281) with the radial-tangential distortion
// model: Project / spaceToPlane distort the normalised point (:217-240), Unproject inverts with 8 fixed-point steps (liftProjective,
// :131-190); the model is active when |k1|, |k2| or |p1| exceeds 1e-5 (:78).  The device residuals apply the same model (lvi_problem_desc.distortion).
class PinholeCamera : public SensorBase {
 public:
  using CameraMatrix = Eigen::Matrix3d;
  PinholeCamera(size_t rows, size_t cols, double readout, double k1, double k2, double p1, double p2, double k3, double fx, double fy, double cx, double cy)
      : rows_(rows), cols_(cols), readout_(readout), k1_(k1), k2_(k2), p1_(p1), p2_(p2), k3_(k3), fx_(fx), fy_(fy), cx_(cx), cy_(cy) {}
  size_t rows() const { return rows_; }
  size_t cols() const { return cols_; }
  void set_rows(size_t r) { rows_ = r; }
  void set_cols(size_t c) { cols_ = c; }
  double readout() const { return readout_; }
  void set_readout(double r) { readout_ = r; }
  double fx() const { return fx_; } double fy() const { return fy_; } double cx() const { return cx_; } double cy() const { return cy_; }
  bool do_distortion() const { return std::fabs(k1_) > 1e-5 || std::fabs(k2_) > 1e-5 || std::fabs(p1_) > 1e-5; }   // :78 (p1 twice, never p2 / k3)
  std::vector<double> distortion_params() const { return {k1_, k2_, p1_, p2_, k3_}; }
  void distortion(const Eigen::Vector2d& p_u, Eigen::Vector2d& d_u) const {   // :199-215
    const double x = p_u(0), y = p_u(1), x2 = x * x, y2 = y * y, xy = x * y, r2 = x2 + y2;
    const double rad = k1_ * r2 + k2_ * r2 * r2 + k3_ * r2 * r2 * r2;
    d_u = Eigen::Vector2d(x * rad + 2.0 * p1_ * xy + p2_ * (r2 + 2.0 * x2), y * rad + 2.0 * p2_ * xy + p1_ * (r2 + 2.0 * y2));
  }
  void liftProjective(const Eigen::Vector2d& p, Eigen::Vector3d& P) const {   // :131-190
    const double mx_d = (p(0) - cx_) / fx_, my_d = (p(1) - cy_) / fy_;
    double mx_u = mx_d, my_u = my_d;
    if (do_distortion()) {
      Eigen::Vector2d d_u;
      distortion(Eigen::Vector2d(mx_d, my_d), d_u);
      mx_u = mx_d - d_u(0); my_u = my_d - d_u(1);
      for (int i = 1; i < 8; ++i) { distortion(Eigen::Vector2d(mx_u, my_u), d_u); mx_u = mx_d - d_u(0); my_u = my_d - d_u(1); }
    }
    P = Eigen::Vector3d(mx_u, my_u, 1.0);
  }
  CameraMatrix camera_matrix() const { CameraMatrix K = CameraMatrix::Identity(); K(0, 0) = fx_; K(1, 1) = fy_; K(0, 2) = cx_; K(1, 2) = cy_; return K; }
  // PinholeView::Project / Unproject / spaceToPlane (pinhole_camera.h:96-124, 217-240)
  Eigen::Vector2d Project(const Eigen::Vector3d& X) const {
    const double z = 1e-32 + X(2);
    Eigen::Vector2d p_d(X(0) / z, X(1) / z);
    if (do_distortion()) { Eigen::Vector2d d_u; distortion(p_d, d_u); p_d = p_d + d_u; }
    return Eigen::Vector2d(fx_ * p_d(0) + cx_, fy_ * p_d(1) + cy_);
  }
  Eigen::Vector2d spaceToPlane(const Eigen::Vector3d& X) const { return Project(X); }
  Eigen::Vector3d Unproject(const Eigen::Vector2d& y) const { Eigen::Vector3d P; liftProjective(y, P); return P; }
 private:
  size_t rows_, cols_;
  double readout_, k1_, k2_, p1_, p2_, k3_, fx_, fy_, cx_, cy_;
};
}  // namespace sensors

namespace sfm {
class Landmark;
class View;
class Observation {   // observation.h:14-36
 public:
  Observation(const Eigen::Vector2d& uv, std::shared_ptr<Landmark> landmark, std::shared_ptr<View> view) : uv_(uv), landmark_(landmark), view_(view) {}
  std::shared_ptr<Landmark> landmark() const { return landmark_; }
  std::shared_ptr<View> view() const { return view_.lock(); }
  bool IsReference() const;
  Eigen::Vector2d uv() const { return uv_; }
  void set_uv(const Eigen::Vector2d& uv) { uv_ = uv; }
  double u() const { return uv_(0); }
  double v() const { return uv_(1); }
 private:
  Eigen::Vector2d uv_;
  std::shared_ptr<Landmark> landmark_;
  std::weak_ptr<View> view_;
};
class Landmark {   // landmark.h:15-50: one inverse-depth block, observations held weakly, unlocked by default
 public:
  Landmark() : id_(next_id()++) {}
  size_t id() const { return id_; }
  void set_reference(std::shared_ptr<Observation> o) { reference_ = o; }
  std::shared_ptr<Observation> reference() const { return reference_.lock(); }
  std::vector<std::shared_ptr<Observation>> observations() const { std::vector<std::shared_ptr<Observation>> out; for (auto& w : observations_) if (auto s = w.lock()) out.push_back(s); return out; }
  double inverse_depth() const { return inverse_depth_; }
  void set_inverse_depth(double x) { inverse_depth_ = x; }
  double* inverse_depth_ptr() { return &inverse_depth_; }
  void Lock(bool flag) { locked_ = flag; }
  bool IsLocked() const { return locked_; }
  void AddObservation(std::shared_ptr<Observation> o) { observations_.push_back(o); }
 private:
  static size_t& next_id() { static size_t n = 0; return n; }
  size_t id_;
  double inverse_depth_ = 0.0;
  bool locked_ = false;
  std::weak_ptr<Observation> reference_;
  std::vector<std::weak_ptr<Observation>> observations_;
};
class View : public std::enable_shared_from_this<View> {   // view.h:15-33
 public:
  View(size_t frame, double t0) : frame_nr_(frame), t0_(t0) {}
  size_t frame_nr() const { return frame_nr_; }
  void set_frame_nr(size_t fnr) { frame_nr_ = fnr; }
  double t0() const { return t0_; }
  void set_t0(double t0) { t0_ = t0; }
  std::vector<std::shared_ptr<Observation>> observations() const { return observations_; }
  std::shared_ptr<Observation> CreateObservation(std::shared_ptr<Landmark> landmark, const Eigen::Vector2d& uv) {   // view_impl.cc:46-55
    auto obs = std::make_shared<Observation>(uv, landmark, shared_from_this());
    observations_.push_back(obs);
    landmark->AddObservation(obs);
    return obs;
  }
 private:
  size_t frame_nr_;
  double t0_;
  std::vector<std::shared_ptr<Observation>> observations_;
};
inline bool Observation::IsReference() const { return landmark_->reference().get() == this; }
}  // namespace sfm

template <class TrajectoryModel> class TrajectoryEstimator;

namespace measurements {
namespace detail {
// single-measurement evaluation on the device: a problem with this one measurement, evaluated at the current parameters
template <class M, class TrajectoryModel> std::vector<double> residual_of(const M& m, const TrajectoryModel& traj);
}
// GyroscopeMeasurement(imu, t, w, weight) — gyroscope_measurement.h:19-48
template <typename ImuModel> class GyroscopeMeasurement {
 public:
  using Vector3 = Eigen::Vector3d;
  GyroscopeMeasurement(std::shared_ptr<ImuModel> imu, double t, const Vector3& w, double weight) : imu_(imu), t(t), w(w), weight(weight) {}
  GyroscopeMeasurement(std::shared_ptr<ImuModel> imu, double t, const Vector3& w) : GyroscopeMeasurement(imu, t, w, 1.0) {}
  template <typename TrajectoryModel> Vector3 Error(const TrajectoryModel& trajectory) const { auto r = detail::residual_of(*this, trajectory); return Vector3(r[0], r[1], r[2]); }
  template <typename TrajectoryModel> Vector3 ErrorRaw(const TrajectoryModel& trajectory) const { return Error(trajectory) / weight; }
  template <typename TrajectoryModel> Vector3 Measure(const TrajectoryModel& trajectory) const { return w - ErrorRaw(trajectory); }
  std::shared_ptr<ImuModel> imu_;
  double t; Vector3 w; double weight;
};
// AccelerometerMeasurement(imu, t, a, weight) — accelerometer_measurement.h:20-49
template <typename ImuModel> class AccelerometerMeasurement {
 public:
  using Vector3 = Eigen::Vector3d;
  AccelerometerMeasurement(std::shared_ptr<ImuModel> imu, double t, const Vector3& a, double weight) : imu_(imu), t(t), a(a), weight(weight) {}
  AccelerometerMeasurement(std::shared_ptr<ImuModel> imu, double t, const Vector3& a) : AccelerometerMeasurement(imu, t, a, 1.0) {}
  template <typename TrajectoryModel> Vector3 Error(const TrajectoryModel& trajectory) const { auto r = detail::residual_of(*this, trajectory); return Vector3(r[0], r[1], r[2]); }
  template <typename TrajectoryModel> Vector3 ErrorRaw(const TrajectoryModel& trajectory) const { return Error(trajectory) / weight; }
  template <typename TrajectoryModel> Vector3 Measure(const TrajectoryModel& trajectory) const { return a - ErrorRaw(trajectory); }
  std::shared_ptr<ImuModel> imu_;
  double t; Vector3 a; double weight;
};
// LiDARSurfelPoint(lidar, point, plane, timestamp, map_time, huber_loss, weight) — lidar_surfel_point.h:18-27
template <typename LiDARModel> class LiDARSurfelPoint {
 public:
  using Vector3 = Eigen::Vector3d;
  LiDARSurfelPoint(std::shared_ptr<LiDARModel> lidar, Vector3 lidar_point, double* plane, double timestamp, double map_time, double huber_loss, double weight)
      : lidar_(lidar), lidar_point_(lidar_point), plane_(plane), timestamp_(timestamp), map_time_(map_time), loss_function_(huber_loss), weight(weight) {}
  LiDARSurfelPoint(std::shared_ptr<LiDARModel> lidar, Vector3 lidar_point, double* plane, double timestamp, double map_time, double huber_loss)
      : LiDARSurfelPoint(lidar, lidar_point, plane, timestamp, map_time, huber_loss, 1.0) {}
  LiDARSurfelPoint(std::shared_ptr<LiDARModel> lidar, Vector3 lidar_point, double* plane, double timestamp, double map_time)
      : LiDARSurfelPoint(lidar, lidar_point, plane, timestamp, map_time, 5.) {}
  // point2plane (lidar_surfel_point.h:85-89): the weighted point-to-plane distance, before the loss
  template <typename TrajectoryModel> Eigen::VecN<1> point2plane(const TrajectoryModel& trajectory) const { Eigen::VecN<1> r; r(0) = detail::residual_of(*this, trajectory)[0]; return r; }
  template <typename TrajectoryModel> Eigen::VecN<1> Error(const TrajectoryModel& trajectory) const { return point2plane(trajectory); }
  std::shared_ptr<LiDARModel> lidar_;
  Vector3 lidar_point_;
  double* plane_;
  double timestamp_, map_time_;
  ceres::HuberLoss loss_function_;
  double weight;
};
// StaticRsCameraMeasurement(camera, obs, huber_loss, weight) — static_rscamera_measurement.h:67-74
template <typename CameraModel> class StaticRsCameraMeasurement {
 public:
  using Vector2 = Eigen::Vector2d;
  StaticRsCameraMeasurement(std::shared_ptr<CameraModel> camera, std::shared_ptr<sfm::Observation> obs, double huber_loss, double weight)
      : camera(camera), observation(obs), loss_function_(huber_loss), weight(weight) {}
  StaticRsCameraMeasurement(std::shared_ptr<CameraModel> camera, std::shared_ptr<sfm::Observation> obs, double huber_loss) : StaticRsCameraMeasurement(camera, obs, huber_loss, 1.0) {}
  StaticRsCameraMeasurement(std::shared_ptr<CameraModel> camera, std::shared_ptr<sfm::Observation> obs) : StaticRsCameraMeasurement(camera, obs, 5.) {}
  template <typename TrajectoryModel> Vector2 Error(const TrajectoryModel& trajectory) const { auto r = detail::residual_of(*this, trajectory); return Vector2(r[0], r[1]); }
  template <typename TrajectoryModel> Vector2 Project(const TrajectoryModel& trajectory) const { return observation->uv() - Error(trajectory) / weight; }
  template <typename TrajectoryModel> Vector2 Measure(const TrajectoryModel& trajectory) const { return Project(trajectory); }
  std::shared_ptr<CameraModel> camera;
  std::shared_ptr<sfm::Observation> observation;
  ceres::HuberLoss loss_function_;
  double weight;
};
// CameraSurfelLandmark(camera, lidar, landmark, plane, timestamp, map_time, huber_loss, weight) — camera_surfel_landmark.h:19-26
template <typename CameraModel, typename LiDARModel> class CameraSurfelLandmark {
 public:
  CameraSurfelLandmark(std::shared_ptr<CameraModel> camera, std::shared_ptr<LiDARModel> lidar, sfm::Landmark* lm, double* plane, double timestamp, double map_time,
                       double huber_loss, double weight)
      : camera_(camera), lidar_(lidar), landmark_(lm), plane_(plane), timestamp_(timestamp), map_time_(map_time), loss_function_(huber_loss), weight(weight) {}
  CameraSurfelLandmark(std::shared_ptr<CameraModel> camera, std::shared_ptr<LiDARModel> lidar, sfm::Landmark* lm, double* plane, double timestamp, double map_time)
      : CameraSurfelLandmark(camera, lidar, lm, plane, timestamp, map_time, 5., 1.0) {}
  template <typename TrajectoryModel> Eigen::VecN<1> point2plane(const TrajectoryModel& trajectory) const { Eigen::VecN<1> r; r(0) = detail::residual_of(*this, trajectory)[0]; return r; }
  template <typename TrajectoryModel> Eigen::VecN<1> Error(const TrajectoryModel& trajectory) const { return point2plane(trajectory); }
  std::shared_ptr<CameraModel> camera_;
  std::shared_ptr<LiDARModel> lidar_;
  sfm::Landmark* landmark_;
  double* plane_;
  double timestamp_, map_time_;
  ceres::HuberLoss loss_function_;
  double weight;
};
// OrientationMeasurement(t, q, weight) — orientation_measurement.h:20-22
class OrientationMeasurement {
 public:
  OrientationMeasurement(double t, const Eigen::Quaterniond& q, double weight) : t(t), q(q), weight(weight) {}
  OrientationMeasurement(double t, const Eigen::Quaterniond& q) : OrientationMeasurement(t, q, 1.0) {}
  template <typename TrajectoryModel> double Error(const TrajectoryModel& trajectory) const { return detail::residual_of(*this, trajectory)[0]; }
  double t; Eigen::Quaterniond q; double weight;
};
}  // namespace measurements

// ---- TrajectoryEstimator ------------------------------------------------------------------------------------------------------------
namespace detail {
// the spline pair behind a trajectory model: SplitTrajectory (both), UniformSO3SplineTrajectory (S0: SO3 only)
inline void splines_of(const trajectories::SplitTrajectory& t, trajectories::UniformR3SplineTrajectory*& r3, trajectories::UniformSO3SplineTrajectory*& so3) {
  t.require_common_grid(); r3 = t.R3Spline().get(); so3 = t.SO3Spline().get();
}
inline void splines_of(const trajectories::UniformSO3SplineTrajectory& t, trajectories::UniformR3SplineTrajectory*& r3, trajectories::UniformSO3SplineTrajectory*& so3) {
  r3 = nullptr; so3 = const_cast<trajectories::UniformSO3SplineTrajectory*>(&t);
}
}  // namespace detail

template <class TrajectoryModel>
class TrajectoryEstimator {
 public:
  using time_span_t = std::pair<double, double>;
  using time_init_t = std::vector<time_span_t>;
  explicit TrajectoryEstimator(std::shared_ptr<TrajectoryModel> trajectory) : trajectory_(trajectory), problem_(DefaultProblemOptions()) { register_trajectory(); }
  static ceres::Problem::Options DefaultProblemOptions() {   // trajectory_estimator.h:22-27
    ceres::Problem::Options options;
    options.loss_function_ownership = ceres::DO_NOT_TAKE_OWNERSHIP;
    options.local_parameterization_ownership = ceres::DO_NOT_TAKE_OWNERSHIP;
    return options;
  }
  auto trajectory() const { return trajectory_; }
  ceres::Problem& problem() { return problem_; }
  // the context the solve runs on (default: lvi_exc_b200::DefaultContext())
  void set_context(lvi_ctx* ctx) { ctx_ = ctx; }

  // trajectory_estimator.h:88-94
  void AddCallback(std::unique_ptr<ceres::IterationCallback> callback, bool needs_state = false) {
    callbacks_.push_back(std::move(callback));
    callback_needs_state_ = callback_needs_state_ || needs_state;
  }
  // trajectory_estimator.h:96-109 — the checks of CheckTimeSpans (:111-130); the blocks themselves are bookkeeping here
  bool AddTrajectoryForTimes(const time_init_t& times) {
    CheckTimeSpans(times);
    return true;
  }
  void CheckTimeSpans(const time_init_t& times) const {
    double t1 = 0, t2 = 0;
    bool first = true;
    for (auto& tt : times) {
      if (tt.first > tt.second) throw std::range_error("At least one time span begins before it ends");
      if (first) { t1 = tt.first; t2 = tt.second; first = false; } else { t1 = std::min(t1, tt.first); t2 = std::max(t2, tt.second); }
    }
    if (!first && (t1 < trajectory_->MinTime() || t2 >= trajectory_->MaxTime())) throw std::range_error("Time span out of range for trajectory");
  }

  template <class MeasurementType> void AddMeasurement(std::shared_ptr<MeasurementType> m) { Add(m); }

  // Solve(max_iterations = 30, progress = true, num_threads = -1) — trajectory_estimator.h:38-68: TRUST_REGION / LEVENBERG_MARQUARDT /
  // SPARSE_SCHUR with Ceres' defaults; num_threads has no meaning on the device and is ignored
  ceres::Solver::Summary Solve(int max_iterations = 30, bool progress = true, int num_threads = -1) {
    (void)num_threads;
    lvi_problem_desc d = Describe();
    lvi_problem* p = nullptr;
    lvi_ctx* ctx = ctx_ ? ctx_ : lvi_exc_b200::DefaultContext();
    lvi_exc_b200::throw_status(lvi_problem_create(ctx, &d, &p));
    lvi_solve_options o;
    lvi_solve_options_default(&o);
    o.max_num_iterations = max_iterations; o.verbose = progress ? 1 : 0;
    lvi_solve_summary raw;
    ceres::Solver::Summary s;
    struct Ctx { TrajectoryEstimator* self; ceres::Solver::Summary* s; } cctx{this, &s};
    auto trampoline = [](const lvi_iteration_summary* it, void* user) -> int {
      Ctx* c = static_cast<Ctx*>(user);
      ceres::IterationSummary is;
      is.iteration = it->iteration; is.step_is_successful = it->step_is_successful != 0; is.cost = it->cost; is.cost_change = it->cost_change;
      is.gradient_max_norm = it->gradient_max_norm; is.step_norm = it->step_norm; is.trust_region_radius = it->trust_region_radius;
      if (c->self->callback_needs_state_) c->self->sync_landmarks();
      int verdict = 0;
      for (auto& cb : c->self->callbacks_) {
        const ceres::CallbackReturnType r = (*cb)(is);
        if (r == ceres::SOLVER_ABORT) verdict = 1; else if (r == ceres::SOLVER_TERMINATE_SUCCESSFULLY && verdict == 0) verdict = 2;
      }
      return verdict;
    };
    const int rc = callbacks_.empty() ? lvi_problem_solve(p, &o, &raw)
                                      : lvi_problem_solve_cb(p, &o, &raw, trampoline, &cctx, callback_needs_state_ ? 1 : 0);
    lvi_problem_destroy(p);
    lvi_exc_b200::throw_status(rc);
    sync_landmarks();   // everything else was updated in place
    static const ceres::TerminationType term[] = {ceres::CONVERGENCE, ceres::NO_CONVERGENCE, ceres::FAILURE, ceres::USER_SUCCESS, ceres::USER_FAILURE};
    s.termination_type = term[raw.termination_type];
    s.initial_cost = raw.initial_cost; s.final_cost = raw.final_cost; s.fixed_cost = raw.fixed_cost;
    s.num_successful_steps = raw.num_successful_steps; s.num_unsuccessful_steps = raw.num_unsuccessful_steps;
    s.num_residual_blocks = raw.num_residual_blocks; s.num_residuals = raw.num_residuals; s.num_effective_parameters = raw.num_effective_parameters;
    s.num_parameter_blocks = problem_.NumParameterBlocks(); s.num_parameters = problem_.NumParameters();
    s.total_time_in_seconds = raw.time_total_ms * 1e-3; s.jacobian_evaluation_time_in_seconds = raw.time_jacobian_ms * 1e-3;
    s.linear_solver_time_in_seconds = raw.time_linear_solve_ms * 1e-3;
    for (int k = 0; k < raw.n_log; ++k) {
      ceres::IterationSummary is;
      is.iteration = k; is.cost = raw.log_cost[k]; is.cost_change = raw.log_cost_change[k]; is.gradient_max_norm = raw.log_gradient_max_norm[k];
      is.step_norm = raw.log_step_norm[k]; is.trust_region_radius = raw.log_radius[k]; is.step_is_successful = raw.log_successful[k] != 0;
      s.iterations.push_back(is);
    }
    return s;
  }

  // One evaluation of every residual at the current parameters on the device (Problem::Evaluate): cost and the corrected residuals in
  // table order gyro, accel, surfel, camera, camera-surfel, orientation (used by TrajectoryManagerLVI::printErrorStatistics)
  double Evaluate(std::vector<double>* residuals = nullptr) {
    lvi_problem_desc d = Describe();
    lvi_problem* p = nullptr;
    lvi_ctx* ctx = ctx_ ? ctx_ : lvi_exc_b200::DefaultContext();
    lvi_exc_b200::throw_status(lvi_problem_create(ctx, &d, &p));
    double cost = 0;
    if (residuals) residuals->assign(static_cast<size_t>(lvi_problem_num_residuals(p)), 0.0);
    const int rc = lvi_problem_evaluate(p, &cost, residuals ? residuals->data() : nullptr, nullptr);
    lvi_problem_destroy(p);
    lvi_exc_b200::throw_status(rc);
    return cost;
  }
  size_t count(int table) const { return table == 0 ? gyro_t_.size() : table == 1 ? acc_t_.size() : table == 2 ? sf_t_.size() : table == 3 ? cam_tr_.size() : table == 4 ? cs_t_.size() : or_t_.size(); }

  // the flat problem the C-ABI takes; pointers stay valid until the next AddMeasurement
  lvi_problem_desc Describe() {
    lvi_problem_desc d{};
    trajectories::UniformR3SplineTrajectory* r3 = nullptr;
    trajectories::UniformSO3SplineTrajectory* so3 = nullptr;
    detail::splines_of(*trajectory_, r3, so3);
    d.t0 = so3->t0(); d.dt = so3->dt();
    d.n_knots = static_cast<int32_t>(r3 ? std::min(r3->NumKnots(), so3->NumKnots()) : so3->NumKnots());
    d.r3_knots = r3 ? r3->data().data() : nullptr; d.so3_knots = so3->data().data();
    d.lock_r3 = (r3 && r3->IsLocked()) ? 1 : 0; d.lock_so3 = so3->IsLocked() ? 1 : 0;
    d.lock_lidar_q = d.lock_lidar_p = d.lock_cam_q = d.lock_cam_p = d.lock_acc_bias = d.lock_gyr_bias = 1;
    auto refuse_toff = [](const sensors::SensorBase& s, const char* who) {
      if (!s.TimeOffsetIsLocked()) throw std::invalid_argument(std::string("lvi_exc_b200: ") + who + " time offset is unlocked; time-offset optimisation is not built "
                                                               "(SURVEY §8 f-4; L/cfg/lvi.yaml:32 optimize_time_offset = false)");
    };
    if (lidar_) { refuse_toff(*lidar_, "LiDAR"); d.lidar_q = lidar_->q_data(); d.lidar_p = lidar_->p_data(); d.lidar_toff = lidar_->time_offset();
                  d.lock_lidar_q = lidar_->RelativeOrientationIsLocked(); d.lock_lidar_p = lidar_->RelativePositionIsLocked(); }
    if (cam_) { refuse_toff(*cam_, "camera");
                d.cam_q = cam_->q_data(); d.cam_p = cam_->p_data(); d.cam_toff = cam_->time_offset();
                d.lock_cam_q = cam_->RelativeOrientationIsLocked(); d.lock_cam_p = cam_->RelativePositionIsLocked();
                d.fx = cam_->fx(); d.fy = cam_->fy(); d.cx = cam_->cx(); d.cy = cam_->cy(); d.readout = cam_->readout();
                { const std::vector<double> k = cam_->distortion_params(); for (int q = 0; q < 5; ++q) d.distortion[q] = k[q]; }
                d.cam_rows = static_cast<int32_t>(cam_->rows()); d.cam_cols = static_cast<int32_t>(cam_->cols()); }
    if (imu_) { d.gravity = imu_->g_data(); d.acc_bias = imu_->ba_data(); d.gyr_bias = imu_->bg_data(); d.imu_toff = imu_->time_offset();
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

 private:
  // ---- one Add per measurement type: a table row + the bookkeeping AddToEstimator does on the ceres::Problem
  template <class I> void Add(std::shared_ptr<measurements::GyroscopeMeasurement<I>> m) {
    AddTrajectoryForTimes({{m->t, m->t}});   // gyroscope_measurement.h:88-91
    use_imu(m->imu_); gyro_t_.push_back(m->t); push3(gyro_w_, m->w); gyro_wt_.push_back(m->weight); keep_.push_back(m);
    problem_.CountResidualBlocks(1, 3);
  }
  template <class I> void Add(std::shared_ptr<measurements::AccelerometerMeasurement<I>> m) {
    AddTrajectoryForTimes({{m->t, m->t}});
    use_imu(m->imu_); acc_t_.push_back(m->t); push3(acc_a_, m->a); acc_wt_.push_back(m->weight); keep_.push_back(m);
    problem_.CountResidualBlocks(1, 3);
  }
  template <class Li> void Add(std::shared_ptr<measurements::LiDARSurfelPoint<Li>> m) {
    // lidar_surfel_point.h:149-167: spans {map_time, map_time}, {t, t} (widened by the max time offset when it is unlocked)
    AddTrajectoryForTimes({{m->map_time_, m->map_time_}, {m->timestamp_, m->timestamp_}});
    use_lidar(m->lidar_);
    sf_t_.push_back(m->timestamp_); sf_tm_.push_back(m->map_time_); push3(sf_p_, m->lidar_point_); sf_plane_.push_back(plane_id(m->plane_));
    sf_wt_.push_back(m->weight); sf_hb_.push_back(m->loss_function_.a()); keep_.push_back(m);
    problem_.CountResidualBlocks(1, 1);
  }
  template <class C> void Add(std::shared_ptr<measurements::StaticRsCameraMeasurement<C>> m) {
    auto lm = m->observation->landmark();
    auto ref = lm->reference();
    if (!ref) throw std::runtime_error("landmark has no reference observation");
    const double t1 = ref->view()->t0(), t2 = m->observation->view()->t0(), margin = 1e-3, ro = m->camera->readout();   // static_rscamera_measurement.h:165-172
    AddTrajectoryForTimes({{std::min(t1, t2) - margin, std::min(t1, t2) + ro + margin}, {std::max(t1, t2) - margin, std::max(t1, t2) + ro + margin}});
    use_camera(m->camera);
    cam_tr_.push_back(t1); cam_to_.push_back(t2);
    cam_uvr_.push_back(ref->u()); cam_uvr_.push_back(ref->v()); cam_uvo_.push_back(m->observation->u()); cam_uvo_.push_back(m->observation->v());
    cam_lm_.push_back(landmark_id(lm.get())); cam_wt_.push_back(m->weight); cam_hb_.push_back(m->loss_function_.a()); keep_.push_back(m);
    problem_.CountResidualBlocks(1, 2);
  }
  template <class C, class Li> void Add(std::shared_ptr<measurements::CameraSurfelLandmark<C, Li>> m) {
    auto ref = m->landmark_->reference();
    if (!ref) throw std::runtime_error("landmark has no reference observation");
    AddTrajectoryForTimes({{m->map_time_, m->map_time_}, {m->timestamp_, m->timestamp_}});
    use_camera(m->camera_); use_lidar(m->lidar_);
    cs_t_.push_back(m->timestamp_); cs_tm_.push_back(m->map_time_); cs_uv_.push_back(ref->u()); cs_uv_.push_back(ref->v());
    cs_lm_.push_back(landmark_id(m->landmark_)); cs_plane_.push_back(plane_id(m->plane_)); cs_wt_.push_back(m->weight); cs_hb_.push_back(m->loss_function_.a());
    keep_.push_back(m);
    problem_.CountResidualBlocks(1, 1);
  }
  void Add(std::shared_ptr<measurements::OrientationMeasurement> m) {
    AddTrajectoryForTimes({{m->t, m->t}});
    or_t_.push_back(m->t); or_q_.insert(or_q_.end(), {m->q.x(), m->q.y(), m->q.z(), m->q.w()}); or_wt_.push_back(m->weight); keep_.push_back(m);
    problem_.CountResidualBlocks(1, 1);
  }

  // ---- ceres::Problem bookkeeping: what SplineEntity / SensorEntity / ImuEntity::AddToProblem register
  void register_trajectory() {
    trajectories::UniformR3SplineTrajectory* r3 = nullptr;
    trajectories::UniformSO3SplineTrajectory* so3 = nullptr;
    detail::splines_of(*trajectory_, r3, so3);
    if (r3) for (size_t i = 0; i < r3->NumKnots(); ++i) { problem_.AddParameterBlock(&r3->data()[3 * i], 3); if (r3->IsLocked()) problem_.SetParameterBlockConstant(&r3->data()[3 * i]); }
    for (size_t i = 0; i < so3->NumKnots(); ++i) {
      problem_.AddParameterBlock(&so3->data()[4 * i], 4, &quat_param_);
      if (so3->IsLocked()) problem_.SetParameterBlockConstant(&so3->data()[4 * i]);
    }
  }
  void register_sensor(sensors::SensorBase& s) {   // sensors.h:137-167
    problem_.AddParameterBlock(s.q_data(), 4, &quat_param_);
    if (s.RelativeOrientationIsLocked()) problem_.SetParameterBlockConstant(s.q_data()); else problem_.SetParameterBlockVariable(s.q_data());
    problem_.AddParameterBlock(s.p_data(), 3);
    if (s.RelativePositionIsLocked()) problem_.SetParameterBlockConstant(s.p_data()); else problem_.SetParameterBlockVariable(s.p_data());
    problem_.AddParameterBlock(&s.time_offset(), 1);
    problem_.SetParameterLowerBound(&s.time_offset(), 0, -s.max_time_offset());
    problem_.SetParameterUpperBound(&s.time_offset(), 0, s.max_time_offset());
    if (s.TimeOffsetIsLocked()) problem_.SetParameterBlockConstant(&s.time_offset());
  }
  template <class I> void use_imu(std::shared_ptr<I> imu) {
    imu_ = imu; register_sensor(*imu);
    problem_.AddParameterBlock(&imu->gravity_orientation_roll(), 1); problem_.AddParameterBlock(&imu->gravity_orientation_pitch(), 1);   // imu.h:129-142 (never locked, Q5)
    problem_.AddParameterBlock(imu->ba_data(), 3); problem_.AddParameterBlock(imu->bg_data(), 3);                                       // constant_bias_imu.h:104-118
    if (imu->AccelerometerBiasIsLocked()) problem_.SetParameterBlockConstant(imu->ba_data()); else problem_.SetParameterBlockVariable(imu->ba_data());
    if (imu->GyroscopeBiasIsLocked()) problem_.SetParameterBlockConstant(imu->bg_data()); else problem_.SetParameterBlockVariable(imu->bg_data());
  }
  template <class L> void use_lidar(std::shared_ptr<L> lidar) { lidar_ = lidar; register_sensor(*lidar); }
  template <class C> void use_camera(std::shared_ptr<C> cam) { cam_ = cam; register_sensor(*cam); }

  static void push3(std::vector<double>& v, const Eigen::Vector3d& a) { v.insert(v.end(), {a(0), a(1), a(2)}); }
  int32_t plane_id(double* plane) {   // plane pointers alias closest_point_vec_ elements (L/src/core/trajectory_manager_lvi.cpp:566-578)
    auto it = plane_ids_.find(plane);
    if (it != plane_ids_.end()) return it->second;
    const int32_t id = static_cast<int32_t>(planes_.size());
    plane_ids_[plane] = id; planes_.push_back(plane);
    problem_.AddParameterBlock(plane, 3); problem_.SetParameterBlockConstant(plane);   // lidar_surfel_point.h:185-190
    return id;
  }
  int32_t landmark_id(sfm::Landmark* lm) {
    auto it = landmark_ids_.find(lm);
    if (it != landmark_ids_.end()) return it->second;
    const int32_t id = static_cast<int32_t>(landmarks_.size());
    landmark_ids_[lm] = id; landmarks_.push_back(lm);
    problem_.AddParameterBlock(lm->inverse_depth_ptr(), 1);                            // static_rscamera_measurement.h:183-189
    problem_.SetParameterLowerBound(lm->inverse_depth_ptr(), 0, 0.);
    if (lm->IsLocked()) problem_.SetParameterBlockConstant(lm->inverse_depth_ptr());
    return id;
  }
  void sync_landmarks() { for (size_t l = 0; l < landmarks_.size() && l < rho_.size(); ++l) landmarks_[l]->set_inverse_depth(rho_[l]); }

  std::shared_ptr<TrajectoryModel> trajectory_;
  ceres::Problem problem_;
  ceres::EigenQuaternionParameterization quat_param_;
  lvi_ctx* ctx_ = nullptr;
  std::vector<std::unique_ptr<ceres::IterationCallback>> callbacks_;
  bool callback_needs_state_ = false;
  std::shared_ptr<sensors::ConstantBiasImu> imu_;
  std::shared_ptr<sensors::VLP16LiDAR> lidar_;
  std::shared_ptr<sensors::PinholeCamera> cam_;
  std::vector<std::shared_ptr<void>> keep_;   // measurements must outlive the estimator in the reference (Residual holds a const M&); here they are retained
  std::map<double*, int32_t> plane_ids_; std::vector<double*> planes_; std::vector<double> planes_flat_;
  std::map<sfm::Landmark*, int32_t> landmark_ids_; std::vector<sfm::Landmark*> landmarks_; std::vector<double> rho_; std::vector<uint8_t> rho_locked_;
  std::vector<double> gyro_t_, gyro_w_, gyro_wt_, acc_t_, acc_a_, acc_wt_, sf_t_, sf_tm_, sf_p_, sf_wt_, sf_hb_;
  std::vector<int32_t> sf_plane_, cam_lm_, cs_lm_, cs_plane_;
  std::vector<double> cam_tr_, cam_to_, cam_uvr_, cam_uvo_, cam_wt_, cam_hb_, cs_t_, cs_tm_, cs_uv_, cs_wt_, cs_hb_, or_t_, or_q_, or_wt_;
  double dummy_g_[2] = {0.01, 0.01}, dummy_b_[6] = {0, 0, 0, 0, 0, 0};
};

namespace measurements {
namespace detail {
// the weighted residual of ONE measurement at the current parameters, before the loss (what Error / point2plane return in the reference):
// a one-row estimator, evaluated on the device.  The loss corrector scales a residual only beyond the Huber threshold; Error() wants the
// raw weighted residual, so the row is evaluated with the threshold lifted out of reach.
template <class M> inline void lift_loss(M&) {}
template <class L> inline void lift_loss(LiDARSurfelPoint<L>& m) { m.loss_function_ = ceres::HuberLoss(1e150); }
template <class C> inline void lift_loss(StaticRsCameraMeasurement<C>& m) { m.loss_function_ = ceres::HuberLoss(1e150); }
template <class C, class L> inline void lift_loss(CameraSurfelLandmark<C, L>& m) { m.loss_function_ = ceres::HuberLoss(1e150); }
template <class M, class TrajectoryModel> std::vector<double> residual_of(const M& m, const TrajectoryModel& traj) {
  auto tshared = std::shared_ptr<TrajectoryModel>(const_cast<TrajectoryModel*>(&traj), [](TrajectoryModel*) {});
  TrajectoryEstimator<TrajectoryModel> est(tshared);
  auto mcopy = std::make_shared<M>(m);
  lift_loss(*mcopy);
  est.template AddMeasurement<M>(mcopy);
  std::vector<double> r;
  est.Evaluate(&r);
  return r;
}
}  // namespace detail
}  // namespace measurements

}  // namespace kontiki
#endif
