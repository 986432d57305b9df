// vi_init/initial_ex_rotation.h — sensor-to-IMU rotation from pairs of relative rotations (quaternion hand-eye equation), under the
// reference's include path (L/include/vi_init/initial_ex_rotation.h:10-38, L/src/vi_init/initial_ex_rotation.cpp:4-77).  The camera
// variant that starts from feature correspondences (CalibrationExRotation, OpenCV essential matrix) is not on the LiDAR / ORB-pose path
// of lvi_init_orb_surfel and is not carried.
#pragma once
#include <algorithm>
#include <vector>

#include <Eigen/Dense>

#include "vi_init/dense_small.h"

class InitialEXRotation {
 public:
  InitialEXRotation() { Reset(); }

  void Reset() {
    frame_count = 0;
    Rc.assign(1, Eigen::Matrix3d::Identity());
    Rc_g.assign(1, Eigen::Matrix3d::Identity());
    Rimu.assign(1, Eigen::Matrix3d::Identity());
    ric = Eigen::Matrix3d::Identity();
  }

  bool CalibrationExRotationLiDAR(Eigen::Matrix3d delta_R_lidar, Eigen::Quaterniond delta_q_imu, Eigen::Matrix3d& calib_ric_result) {
    frame_count++;
    Rc.push_back(delta_R_lidar);
    Rimu.push_back(delta_q_imu.toRotationMatrix());
    Rc_g.push_back(ric.transpose() * delta_q_imu.toRotationMatrix() * ric);   // ric.inverse() * delta_q_imu * ric
    lvi_init::MatX A(frame_count * 4, 4);
    for (int i = 1; i <= frame_count; i++) {
      const Eigen::Quaterniond r1(Rc[i]), r2(Rc_g[i]);
      const double angular_distance = 180 / M_PI * r1.angularDistance(r2);
      const double huber = angular_distance > 5.0 ? 5.0 / angular_distance : 1.0;
      double L[4][4], R[4][4];
      quat_left(r1, L);
      quat_right(Eigen::Quaterniond(Rimu[i]), R);
      for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) A((i - 1) * 4 + r, c) = huber * (L[r][c] - R[r][c]);
    }
    // JacobiSVD(A): V and the singular values from the eigen-decomposition of A^T A (4 x 4)
    std::vector<double> ev;
    lvi_init::MatX V;
    lvi_init::jacobi_eig(lvi_init::gram(A, 1.0), ev, V);
    int order[4] = {0, 1, 2, 3};
    std::sort(order, order + 4, [&](int a, int b) { return ev[a] > ev[b]; });
    const int k = order[3];   // svd.matrixV().col(3): singular vector of the smallest singular value, coefficients x y z w
    Eigen::Quaterniond estimated_R(V(3, k), V(0, k), V(1, k), V(2, k));
    estimated_R.normalize();
    ric = estimated_R.toRotationMatrix().transpose();
    const double second_smallest = std::sqrt(std::max(ev[order[2]], 0.0));   // singularValues().tail<3>()(1)
    if (frame_count * 4 >= 3 && second_smallest > 0.25) {
      calib_ric_result = ric;
      return true;
    }
    return false;
  }

 private:
  static void quat_left(const Eigen::Quaterniond& q, double L[4][4]) {
    const double w = q.w(), v[3] = {q.x(), q.y(), q.z()};
    const double S[3][3] = {{0, -v[2], v[1]}, {v[2], 0, -v[0]}, {-v[1], v[0], 0}};
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) L[r][c] = (r == c ? w : 0.0) + S[r][c]; L[r][3] = v[r]; L[3][r] = -v[r]; }
    L[3][3] = w;
  }
  static void quat_right(const Eigen::Quaterniond& q, double R[4][4]) {
    const double w = q.w(), v[3] = {q.x(), q.y(), q.z()};
    const double S[3][3] = {{0, -v[2], v[1]}, {v[2], 0, -v[0]}, {-v[1], v[0], 0}};
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R[r][c] = (r == c ? w : 0.0) - S[r][c]; R[r][3] = v[r]; R[3][r] = -v[r]; }
    R[3][3] = w;
  }
  int frame_count;
  std::vector<Eigen::Matrix3d> Rc, Rimu, Rc_g;
  Eigen::Matrix3d ric;
};
