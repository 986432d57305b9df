// vi_init/initial_alignment.h — gravity / velocities / translation extrinsic (and scale) by linear alignment of the pre-integrations with
// the sensor's odometry, under the reference's include path (L/include/vi_init/initial_alignment.h:13-35,
// L/src/vi_init/initial_aligment.cpp:48-66,128-171,175-423,518-526).  Reference behaviour that is kept: the loops stop two frames before
// the end; the velocity blocks of both gravity systems go to columns 0..5 for EVERY frame pair (`A.block<3, 6>(i * 3, 0)`, :167,301), so
// only v_0 and v_1 are observed and x.segment<3>(3 i) is zero for i >= 2; |g| = 9.7964; the gyroscope-bias step is disabled.
#pragma once
#include <cmath>
#include <deque>
#include <vector>

#include <Eigen/Dense>

#include "vi_init/dense_small.h"
#include "vi_init/integration_base.h"

class ImageFrame {
 public:
  ImageFrame() {}
  double t = 0;
  Eigen::Matrix3d R;
  Eigen::Vector3d T;
  IntegrationBase* pre_integration = nullptr;
  bool is_key_frame = false;
};

namespace lvi_init {
inline void tangent_basis(const Eigen::Vector3d& g0, double lxly[3][2]) {   // initial_aligment.cpp:48-61
  const Eigen::Vector3d a = g0.normalized();
  Eigen::Vector3d tmp(0, 0, 1);
  if (a(0) == 0 && a(1) == 0 && a(2) == 1) tmp = Eigen::Vector3d(1, 0, 0);
  const Eigen::Vector3d b = (tmp - a * a.dot(tmp)).normalized();
  const Eigen::Vector3d c = a.cross(b);
  for (int r = 0; r < 3; ++r) { lxly[r][0] = b(r); lxly[r][1] = c(r); }
}
// the velocity / gravity system of LinearAlignment (free = 3: g itself) and of RefineGravity (free = 2: its tangent coordinates)
inline std::vector<double> gravity_system(const std::deque<ImageFrame>& f, const double (*lxly)[2], const Eigen::Vector3d* g0) {
  const int n = static_cast<int>(f.size()), free_dims = lxly ? 2 : 3;
  MatX A((n - 1) * 3, n * 3 + free_dims);
  std::vector<double> b((n - 1) * 3, 0.0);
  for (int i = 0; i < n - 2; ++i) {
    const ImageFrame& fi = f[i];
    const ImageFrame& fj = f[i + 1];
    const double dt = fj.pre_integration->sum_dt;
    const Eigen::Matrix3d RiT = fi.R.transpose(), RiRj = RiT * fj.R;
    for (int r = 0; r < 3; ++r) {
      A(3 * i + r, r) += -1.0;                                      // columns 0..5, not 3 i .. 3 i + 5 (kept)
      for (int c = 0; c < 3; ++c) A(3 * i + r, 3 + c) += RiRj(r, c);
      for (int c = 0; c < free_dims; ++c) {
        double v = 0;
        if (lxly) { for (int k = 0; k < 3; ++k) v += RiT(r, k) * dt * lxly[k][c]; } else v = RiT(r, c) * dt;
        A(3 * i + r, 3 * n + c) += v;
      }
    }
    Eigen::Vector3d rhs = fj.pre_integration->delta_v;
    if (g0) rhs = rhs - RiT * ((*g0) * dt);
    for (int r = 0; r < 3; ++r) b[3 * i + r] += rhs(r);
  }
  return solve_semidefinite(gram(A, 1000.0), atb(A, b, 1000.0));
}
}  // namespace lvi_init

inline void RefineGravity(const std::deque<ImageFrame>& all_image_frame, Eigen::Vector3d& g, Eigen::VectorXd& x) {
  Eigen::Vector3d g0 = g.normalized() * 9.7964;
  const int n = static_cast<int>(all_image_frame.size());
  for (int k = 0; k < 4; k++) {
    double lxly[3][2];
    lvi_init::tangent_basis(g0, lxly);
    const std::vector<double> sol = lvi_init::gravity_system(all_image_frame, lxly, &g0);
    x = Eigen::VectorXd(sol);
    Eigen::Vector3d step;
    for (int r = 0; r < 3; ++r) step(r) = lxly[r][0] * sol[3 * n] + lxly[r][1] * sol[3 * n + 1];
    g0 = (g0 + step).normalized() * 9.7964;
  }
  g = g0;
}

inline bool LinearAlignment(const std::deque<ImageFrame>& all_image_frame, Eigen::Vector3d& g, Eigen::Vector3d& T_I_C, Eigen::VectorXd& x, bool fix_scale) {
  const int n = static_cast<int>(all_image_frame.size());
  {
    const std::vector<double> sol = lvi_init::gravity_system(all_image_frame, nullptr, nullptr);
    x = Eigen::VectorXd(sol);
    g = Eigen::Vector3d(sol[3 * n], sol[3 * n + 1], sol[3 * n + 2]);
  }
  RefineGravity(all_image_frame, g, x);
  if (!(std::fabs(g.norm() - 9.7964) < 0.5)) return false;
  const int m = fix_scale ? 3 : 4;
  lvi_init::MatX A((n - 1) * 3, m);
  std::vector<double> b((n - 1) * 3, 0.0);
  for (int i = 0; i < n - 2; ++i) {
    const ImageFrame& fi = all_image_frame[i];
    const ImageFrame& fj = all_image_frame[i + 1];
    const double dt = fj.pre_integration->sum_dt;
    const Eigen::Matrix3d RiT = fi.R.transpose(), RiRj = RiT * fj.R;
    const Eigen::Vector3d dT = RiT * (fj.T - fi.T);
    const Eigen::Vector3d v(x(3 * i), x(3 * i + 1), x(3 * i + 2));
    Eigen::Vector3d rhs = fj.pre_integration->delta_p + dt * v - (RiT * g) * (dt * dt / 2);
    if (fix_scale) rhs = rhs - dT;
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) A(3 * i + r, c) += (r == c ? 1.0 : 0.0) - RiRj(r, c);
      if (!fix_scale) A(3 * i + r, 3) += dT(r);
      b[3 * i + r] += rhs(r);
    }
  }
  const std::vector<double> t = lvi_init::solve_semidefinite(lvi_init::gram(A, 1.0), lvi_init::atb(A, b, 1.0));
  T_I_C = Eigen::Vector3d(t[0], t[1], t[2]);
  if (fix_scale) return true;
  x = Eigen::VectorXd(t);
  return t[3] > 0;
}

inline bool VisualIMUAlignment(const std::deque<ImageFrame>& all_image_frame, Eigen::Vector3d* /*Bgs*/, Eigen::Vector3d& g, Eigen::Vector3d& T_I_C,
                               Eigen::VectorXd& x, bool fix_scale = false) {
  return LinearAlignment(all_image_frame, g, T_I_C, x, fix_scale);   // solveGyroscopeBias is commented out in the reference (:521)
}
