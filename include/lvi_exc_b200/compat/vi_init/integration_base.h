// vi_init/integration_base.h — mid-point IMU pre-integration between two sensor frames, under the reference's include path
// (L/include/vi_init/integration_base.h:17-323).  Host code, as in the reference (the initial-guess stage is small dense
// algebra, nothing data-parallel).  What the initialisation reads is kept: delta_p / delta_q / delta_v / sum_dt and the seeding rule of
// push_back; the 15x15 Jacobian / covariance propagation is not (VisualIMUAlignment has the gyroscope-bias step commented out,
// L/src/vi_init/initial_aligment.cpp:521, so nothing on this path reads them).
#pragma once
#include <vector>

#include <Eigen/Dense>

struct IMUNoise { double ACC_N = 0, GYR_N = 0, ACC_W = 0, GYR_W = 0; };   // kept for the constructor's signature (covariance only)

class IntegrationBase {
 public:
  IntegrationBase() = delete;
  IntegrationBase(const Eigen::Vector3d& _linearized_ba, const Eigen::Vector3d& _linearized_bg, const IMUNoise& = IMUNoise())
      : dt(-1.0), linearized_ba(_linearized_ba), linearized_bg(_linearized_bg), sum_dt(0.0) {}

  double deltaTij() { return sum_dt; }
  Eigen::Vector3d deltaPij() { return delta_p; }
  Eigen::Vector3d deltaVij() { return delta_v; }
  Eigen::Matrix3d deltaRij() { return delta_q.toRotationMatrix(); }
  Eigen::Quaterniond deltaQij() { return delta_q; }

  // :103-117 — the first sample of an integrator only seeds the mid-point rule
  void push_back(double _dt, const Eigen::Vector3d& _acc, const Eigen::Vector3d& _gyr) {
    if (dt < 0.) {
      dt = 1e-6;
      acc_0 = _acc; gyr_0 = _gyr;
      linearized_acc = acc_0; linearized_gyr = gyr_0;
      return;
    }
    dt_buf.push_back(_dt); acc_buf.push_back(_acc); gyr_buf.push_back(_gyr);
    propagate(_dt, _acc, _gyr);
  }

  // :134-149
  void repropagate(const Eigen::Vector3d& _linearized_ba, const Eigen::Vector3d& _linearized_bg) {
    sum_dt = 0.0;
    acc_0 = linearized_acc; gyr_0 = linearized_gyr;
    delta_p = Eigen::Vector3d::Zero(); delta_q = Eigen::Quaterniond::Identity(); delta_v = Eigen::Vector3d::Zero();
    linearized_ba = _linearized_ba; linearized_bg = _linearized_bg;
    for (size_t i = 0; i < dt_buf.size(); ++i) propagate(dt_buf[i], acc_buf[i], gyr_buf[i]);
  }

  // :258-284 (midPointIntegration :151-175 inlined)
  void propagate(double _dt, const Eigen::Vector3d& _acc_1, const Eigen::Vector3d& _gyr_1) {
    dt = _dt;
    const Eigen::Vector3d un_acc_0 = rotate_raw(delta_q, acc_0 - linearized_ba);
    const Eigen::Vector3d un_gyr = 0.5 * (gyr_0 + _gyr_1) - linearized_bg;
    Eigen::Quaterniond result_delta_q = delta_q * Eigen::Quaterniond(1, un_gyr(0) * _dt / 2, un_gyr(1) * _dt / 2, un_gyr(2) * _dt / 2);
    const Eigen::Vector3d un_acc_1 = rotate_raw(result_delta_q, _acc_1 - linearized_ba);   // with the not yet normalised coefficients, as Eigen's q * v does
    const Eigen::Vector3d un_acc = 0.5 * (un_acc_0 + un_acc_1);
    delta_p = delta_p + delta_v * _dt + 0.5 * un_acc * _dt * _dt;
    delta_v = delta_v + un_acc * _dt;
    delta_q = result_delta_q;
    delta_q.normalize();
    sum_dt += dt;
    acc_0 = _acc_1; gyr_0 = _gyr_1;
  }

  // Eigen's Quaternion * Vector3 on raw coefficients: v + w (2 u x v) + u x (2 u x v)
  static Eigen::Vector3d rotate_raw(const Eigen::Quaterniond& q, const Eigen::Vector3d& v) {
    const Eigen::Vector3d u(q.x(), q.y(), q.z());
    Eigen::Vector3d uv = u.cross(v);
    uv += uv;
    return v + q.w() * uv + u.cross(uv);
  }

  double dt;
  Eigen::Vector3d acc_0, gyr_0;
  Eigen::Vector3d linearized_acc, linearized_gyr;
  Eigen::Vector3d linearized_ba, linearized_bg;
  double sum_dt;
  Eigen::Vector3d delta_p;
  Eigen::Quaterniond delta_q;
  Eigen::Vector3d delta_v;
  std::vector<double> dt_buf;
  std::vector<Eigen::Vector3d> acc_buf, gyr_buf;
};
