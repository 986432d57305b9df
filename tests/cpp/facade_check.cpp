// facade_check.cpp — exercises include/lvi_exc_b200/kontiki_facade.hpp the way TrajectoryManagerLVI drives Kontiki
// (L/src/core/trajectory_manager_lvi.cpp:43-62,464-606).  Built and run by tests/test_cpp_facade.py.
//   facade_check describe   : CPU only — records measurements and prints the lowered table sizes / flags as JSON
//   facade_check solve      : GPU — initialSO3TrajWithGyro-style fit of a constant-rate rotation, prints cost and orientation error
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>

#include "lvi_exc_b200/kontiki_facade.hpp"

using namespace kontiki;
using Traj = trajectories::SplitTrajectory;

int main(int argc, char** argv) {
  const bool solve = argc > 1 && std::strcmp(argv[1], "solve") == 0;
  const double t_start = 5.0037, dt = 0.02, pad = 0.2, dur = 1.0;
  auto traj = std::make_shared<Traj>(dt, dt, t_start - pad, t_start - pad);
  traj->ExtendTo(t_start + dur + pad, Vector3d(0, 0, 0), Quaterniond::Identity());
  auto imu = std::make_shared<sensors::ConstantBiasImu>();
  auto lidar = std::make_shared<sensors::VLP16LiDAR>();
  auto cam = std::make_shared<sensors::PinholeCamera>(720, 1280, 0.0666, 0, 0, 0, 0, 0, 530.175, 530.095, 635.12, 356.522);
  if (!solve) {
    TrajectoryEstimator<Traj> est(nullptr, traj);
    lidar->LockRelativeOrientation(false); lidar->LockRelativePosition(false);
    imu->LockAccelerometerBias(false);
    for (int i = 0; i < 10; ++i) {
      const double t = t_start + 0.00137 + i * 0.005;
      est.AddMeasurement(std::make_shared<measurements::GyroscopeMeasurement<sensors::ConstantBiasImu>>(imu, t, Vector3d(0, 0, 0.3), 28.0));
      est.AddMeasurement(std::make_shared<measurements::AccelerometerMeasurement<sensors::ConstantBiasImu>>(imu, t, Vector3d(0, 0, 9.79), 18.0));
    }
    double planes[2][3] = {{1.0, 0.0, 0.0}, {0.0, 2.0, 0.0}};
    for (int i = 0; i < 6; ++i)  // two distinct plane pointers, used alternately (they alias closest_point_vec_ in the reference)
      est.AddMeasurement(std::make_shared<measurements::LiDARSurfelPoint<sensors::VLP16LiDAR>>(lidar, Vector3d(1, 0.1 * i, 0), planes[i % 2], t_start + 0.3 + 0.01 * i,
                                                                                                t_start, 5.0, 10.0));
    auto lm = std::make_shared<sfm::Landmark>();
    lm->set_inverse_depth(0.25);
    auto v0 = std::make_shared<sfm::View>(0, t_start + 0.00411), v1 = std::make_shared<sfm::View>(4, t_start + 0.20411);
    auto o0 = std::make_shared<sfm::Observation>(Vector2d(600, 300), lm, v0), o1 = std::make_shared<sfm::Observation>(Vector2d(610, 305), lm, v1);
    lm->set_reference(o0);
    est.AddMeasurement(std::make_shared<measurements::StaticRsCameraMeasurement<sensors::PinholeCamera>>(cam, o1, 5.0));
    est.AddMeasurement(std::make_shared<measurements::CameraSurfelLandmark<sensors::PinholeCamera, sensors::VLP16LiDAR>>(cam, lidar, lm.get(), planes[1], v0->t0(),
                                                                                                                           t_start, 5.0, 30.0));
    lvi_problem_desc d = est.Describe();
    std::printf("{\"n_knots\": %d, \"n_gyro\": %d, \"n_accel\": %d, \"n_surfel\": %d, \"n_cam\": %d, \"n_camsurf\": %d, \"n_planes\": %d, \"n_landmarks\": %d, "
                "\"lock_lidar_q\": %d, \"lock_cam_q\": %d, \"lock_acc_bias\": %d, \"lock_gyr_bias\": %d, \"surfel_plane_3\": %d, \"cs_plane\": %d, \"rho0\": %.3f, "
                "\"cam_t0_ref\": %.5f, \"cam_weight\": %.1f, \"plane1_y\": %.1f, \"blocks\": %zu, \"min_time\": %.4f, \"max_time\": %.4f}\n",
                d.n_knots, d.n_gyro, d.n_accel, d.n_surfel, d.n_cam, d.n_camsurf, d.n_planes, d.n_landmarks, d.lock_lidar_q, d.lock_cam_q, d.lock_acc_bias,
                d.lock_gyr_bias, d.surfel_plane[3], d.cs_plane[0], d.rho[0], d.cam_t0_ref[0] - t_start, d.cam_weight[0], d.planes[4], est.num_residual_blocks(),
                traj->MinTime() - t_start, traj->MaxTime() - t_start);
    // a null context must fail loudly, never fall back to a CPU path
    try { est.Solve(1, false); std::printf("{\"error\": \"Solve without a device did not throw\"}\n"); return 2; }
    catch (const std::exception&) {}
    return 0;
  }
  lvi_ctx* ctx = nullptr;
  throw_status(lvi_ctx_create(0, nullptr, 0, 1, &ctx));
  // initialSO3TrajWithGyro: gyro samples of a constant yaw rate + one orientation anchor at MinTime
  const double rate = 0.7;
  TrajectoryEstimator<Traj> est(ctx, traj, /*so3_only=*/true);
  for (double t = traj->MinTime() + 0.00137; t < traj->MaxTime(); t += 0.005)
    est.AddMeasurement(std::make_shared<measurements::GyroscopeMeasurement<sensors::ConstantBiasImu>>(imu, t, Vector3d(0, 0, rate), 28.0));
  est.AddMeasurement(std::make_shared<measurements::OrientationMeasurement>(traj->MinTime(), Quaterniond::Identity(), 28.0));
  auto summary = est.Solve(30, false);
  Quaterniond q;
  const double tq = t_start + 0.6;
  traj->Evaluate(ctx, tq, nullptr, &q);
  const double yaw = 2.0 * std::atan2(q.z, q.w), expect = rate * (tq - traj->MinTime());
  bool threw = false;
  try { traj->Evaluate(ctx, traj->MaxTime() + 1.0, nullptr, &q); } catch (const std::range_error&) { threw = true; }
  std::printf("{\"iterations\": %d, \"initial_cost\": %.6e, \"final_cost\": %.6e, \"yaw\": %.9f, \"expected_yaw\": %.9f, \"usable\": %d, \"range_error\": %d}\n",
              summary.raw.num_iterations, summary.initial_cost, summary.final_cost, yaw, expect, summary.IsSolutionUsable() ? 1 : 0, threw ? 1 : 0);
  std::fprintf(stderr, "%s\n", summary.BriefReport().c_str());
  lvi_ctx_destroy(ctx);
  return 0;
}
