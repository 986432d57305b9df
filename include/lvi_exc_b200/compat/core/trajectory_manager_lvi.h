// core/trajectory_manager_lvi.h — licalib::TrajectoryManagerLVI (L/include/core/trajectory_manager_lvi.h:96-290,
// L/src/core/trajectory_manager_lvi.cpp:31-62,138-351,386-606) for the stages on the calibration hot path, written against the Kontiki
// surface of this tree (kontiki/kontiki_b200.h), so that every Solve() runs on the B200.  Same public names, argument orders, lock policy per
// stage and copy-back into CalibParamManager as the reference; the LOAM-pose / visual-frame-only initialisers (trajInitFromLidarPose,
// trajInitFromVisualFrames) belong to the initial-guess stage and are not declared.
#ifndef LVI_EXC_B200_COMPAT_CORE_TRAJECTORY_MANAGER_LVI_H
#define LVI_EXC_B200_COMPAT_CORE_TRAJECTORY_MANAGER_LVI_H
#include <cassert>
#include <iostream>
#include <map>
#include <memory>
#include <vector>

#include "../kontiki/kontiki_b200.h"
#include "surfel_association.h"

namespace licalib {
namespace IO {
struct IMUData {   // L/include/utils/dataset_reader.h:45-50
  double timestamp;
  Eigen::Vector3d gyro, accel;
};
}  // namespace IO

struct CameraIntrinsic {   // L/include/core/trajectory_manager_lvi.h:57-70 (the distortion-free set the pipeline runs with, Q13)
  int row = 720, col = 1280;
  double readout = 0.0666;
  double k1 = 0, k2 = 0, p1 = 0, p2 = 0, k3 = 0;
  double fx = 530.175, fy = 530.095, cx = 635.12, cy = 356.522;
};

// licalib::CalibParamManager (L/include/core/calibration.hpp:40-120): the calibration state every stage reads and writes
class CalibParamManager {
 public:
  typedef std::shared_ptr<CalibParamManager> Ptr;
  Eigen::Vector3d p_LinI, p_CinI, gravity = Eigen::Vector3d(0, 0, -9.8), gyro_bias, acce_bias;
  Eigen::Quaterniond q_LtoI, q_CtoI;
  double time_offset = 0;
  double global_opt_gyro_weight = 28.5, global_opt_acce_weight = 18.5, global_opt_lidar_weight = 10.0, global_opt_visual_surfel_weight = 200.;
  double global_opt_pos_weight = 1, global_opt_rot_weight = 0.5, global_opt_cam_weight = 1.0;
  void set_q_LtoI(Eigen::Quaterniond q) { q_LtoI = q; }
  void set_p_LinI(Eigen::Vector3d p) { p_LinI = p; }
  void set_q_CtoI(Eigen::Quaterniond q) { q_CtoI = q; }
  void set_p_CinI(Eigen::Vector3d p) { p_CinI = p; }
  void set_gravity(Eigen::Vector3d g) { gravity = g; }
  void set_time_offset(double t) { time_offset = t; }
  void set_gyro_bias(Eigen::Vector3d b) { gyro_bias = b; }
  void set_acce_bias(Eigen::Vector3d b) { acce_bias = b; }
  void showStates(bool verbose = false) const {
    if (!verbose) return;
    std::cout << "P_LinI " << p_LinI << " | q_LtoI " << q_LtoI.x() << " " << q_LtoI.y() << " " << q_LtoI.z() << " " << q_LtoI.w() << "\nP_CinI " << p_CinI << " | q_CtoI "
              << q_CtoI.x() << " " << q_CtoI.y() << " " << q_CtoI.z() << " " << q_CtoI.w() << "\ngravity " << gravity << " | gyro bias " << gyro_bias << " | acce bias "
              << acce_bias << std::endl;
  }
};

class TrajectoryManagerLVI {
  using IMUSensor = kontiki::sensors::ConstantBiasImu;
  using LiDARSensor = kontiki::sensors::VLP16LiDAR;
  using CameraSensor = kontiki::sensors::PinholeCamera;
  using SO3TrajEstimator = kontiki::TrajectoryEstimator<kontiki::trajectories::UniformSO3SplineTrajectory>;
  using R3TrajEstimator = kontiki::TrajectoryEstimator<kontiki::trajectories::UniformR3SplineTrajectory>;
  using SplitTrajEstimator = kontiki::TrajectoryEstimator<kontiki::trajectories::SplitTrajectory>;
  using GyroMeasurement = kontiki::measurements::GyroscopeMeasurement<IMUSensor>;
  using AccelMeasurement = kontiki::measurements::AccelerometerMeasurement<IMUSensor>;
  using SurfMeasurement = kontiki::measurements::LiDARSurfelPoint<LiDARSensor>;
  using CameraMeasurement = kontiki::measurements::StaticRsCameraMeasurement<CameraSensor>;
  using CameraSurfMeasurement = kontiki::measurements::CameraSurfelLandmark<CameraSensor, LiDARSensor>;
  using OrientationMeasurement = kontiki::measurements::OrientationMeasurement;

 public:
  typedef std::shared_ptr<TrajectoryManagerLVI> Ptr;
  using Result = std::unique_ptr<kontiki::trajectories::TrajectoryEvaluation<double>>;

  explicit TrajectoryManagerLVI(const CameraIntrinsic& ci, double start_time, double end_time, double knot_distance, double time_offset_padding)
      : time_offset_padding_(time_offset_padding), map_time_(0), imu_(std::make_shared<IMUSensor>()), lidar_(std::make_shared<LiDARSensor>()),
        calib_param_manager(std::make_shared<CalibParamManager>()) {
    assert(knot_distance > 0 && "knot_distance should be lager than 0");
    camera_ = std::make_shared<CameraSensor>(ci.row, ci.col, ci.readout, ci.k1, ci.k2, ci.p1, ci.p2, ci.k3, ci.fx, ci.fy, ci.cx, ci.cy);
    camera_->set_max_time_offset(0.001);
    lidar_->set_max_time_offset(0.001);
    const double traj_start_time = start_time - time_offset_padding, traj_end_time = end_time + time_offset_padding;
    traj_ = std::make_shared<kontiki::trajectories::SplitTrajectory>(knot_distance, knot_distance, traj_start_time, traj_start_time);
    initialTrajTo(traj_end_time);
  }

  void initialTrajTo(double max_time) {
    traj_->R3Spline()->ExtendTo(max_time, Eigen::Vector3d(0, 0, 0));
    traj_->SO3Spline()->ExtendTo(max_time, Eigen::Quaterniond::Identity());
  }
  void feedIMUData(const IO::IMUData& data) { imu_data_.emplace_back(data); }

  // S0 (:43-62): SO3 spline alone against the gyro samples, anchored by one orientation measurement at MinTime
  void initialSO3TrajWithGyro() {
    assert(imu_data_.size() > 0 && "[initialSO3TrajWithGyro]: There's NO imu data for initialization.");
    auto estimator_SO3 = std::make_shared<SO3TrajEstimator>(traj_->SO3Spline());
    addGyroscopeMeasurements(estimator_SO3);
    const double weight_t0 = calib_param_manager->global_opt_gyro_weight, t0 = traj_->SO3Spline()->MinTime();
    const Eigen::AngleAxisd rotation_vector(0.0001, Eigen::Vector3d(0, 0, 1));
    const Eigen::Quaterniond q0(rotation_vector.matrix());
    auto m_q0 = std::make_shared<OrientationMeasurement>(t0, q0, weight_t0);
    estimator_SO3->AddMeasurement<OrientationMeasurement>(m_q0);
    last_summary_ = estimator_SO3->Solve(30, false);
    std::cout << last_summary_.BriefReport() << std::endl;
  }

  // S1-S3 (:311-351): IMU + LiDAR-surfel residuals, LiDAR extrinsics and IMU biases free, camera locked
  void trajInitFromSurfel(SurfelAssociation::Ptr surfels_association, bool opt_time_offset_ = false) {
    lidar_->set_relative_orientation(calib_param_manager->q_LtoI);
    lidar_->set_relative_position(calib_param_manager->p_LinI);
    lidar_->LockRelativeOrientation(false);
    lidar_->LockRelativePosition(false);
    camera_->LockRelativeOrientation(true);
    camera_->LockRelativePosition(true);
    lidar_->LockTimeOffset(!(opt_time_offset_ && time_offset_padding_ > 0));
    imu_->LockGyroscopeBias(false);
    imu_->LockAccelerometerBias(false);

    auto estimator_split = std::make_shared<SplitTrajEstimator>(traj_);
    addGyroscopeMeasurements(estimator_split);
    addAccelerometerMeasurement(estimator_split);
    addSurfMeasurement(estimator_split, surfels_association);

    last_summary_ = estimator_split->Solve(30, false);
    std::cout << last_summary_.BriefReport() << std::endl;
    copyBack(/*camera=*/false, lidar_->time_offset());
  }

  // S4 (:138-194) and S5 (:196-257): camera residuals on top, with (S5) the landmark-to-surfel residuals
  void trajInitFromLVIdata(const std::map<int64_t, std::shared_ptr<kontiki::sfm::View>>& frames, SurfelAssociation::Ptr surfels_association,
                           bool opt_time_offset_ = false, bool lock_traj_and_lidar = false) {
    static const std::map<kontiki::sfm::Landmark*, size_t> none;
    solveLVI(frames, surfels_association, none, false, opt_time_offset_, lock_traj_and_lidar);
  }
  void trajInitFromLVIdata(const std::map<int64_t, std::shared_ptr<kontiki::sfm::View>>& frames, SurfelAssociation::Ptr surfels_association,
                           const std::map<kontiki::sfm::Landmark*, size_t>& lm_surfel, bool opt_time_offset_ = false, bool lock_traj_and_lidar = false) {
    solveLVI(frames, surfels_association, lm_surfel, true, opt_time_offset_, lock_traj_and_lidar);
  }

  bool evaluateIMUPose(double imu_time, int flags, Result& result) const {   // :386-392
    if (traj_->MinTime() > imu_time || traj_->MaxTime() <= imu_time) return false;
    result = traj_->Evaluate(imu_time, flags);
    return true;
  }
  bool evaluateLidarPose(double lidar_time, Eigen::Quaterniond& q_LtoG, Eigen::Vector3d& p_LinG) const {   // :394-404
    return sensorPose(lidar_time + lidar_->time_offset(), calib_param_manager->q_LtoI, calib_param_manager->p_LinI, q_LtoG, p_LinG);
  }
  bool evaluateCameraPose(double camera_time, Eigen::Quaterniond& q_CtoG, Eigen::Vector3d& p_CinG) const {   // :427-437
    return sensorPose(camera_time + camera_->time_offset(), calib_param_manager->q_CtoI, calib_param_manager->p_CinI, q_CtoG, p_CinG);
  }
  bool evaluateLidarRelativeRotation(double lidar_time1, double lidar_time2, Eigen::Quaterniond& q_L2toL1) const {   // :406-424
    return relativeRotation(lidar_time1 + lidar_->time_offset(), lidar_time2 + lidar_->time_offset(), calib_param_manager->q_LtoI, q_L2toL1);
  }
  bool evaluateCameraRelativeRotation(double camera_time1, double camera_time2, Eigen::Quaterniond& q_C2toC1) const {   // :439-461
    return relativeRotation(camera_time1 + camera_->time_offset(), camera_time2 + camera_->time_offset(), calib_param_manager->q_CtoI, q_C2toC1);
  }

  CalibParamManager::Ptr getCalibParamManager() const { return calib_param_manager; }
  double get_map_time() const { return map_time_; }
  std::shared_ptr<kontiki::trajectories::SplitTrajectory> getTrajectory() const { return traj_; }
  std::shared_ptr<CameraSensor> getCameraModel() const { return camera_; }
  std::shared_ptr<LiDARSensor> getLidarModel() const { return lidar_; }
  std::shared_ptr<IMUSensor> getIMUModel() const { return imu_; }
  const ceres::Solver::Summary& lastSummary() const { return last_summary_; }

  // the flat trajectory description the C-ABI's de-skew / pose calls take (lvi_undistort, lvi_trajectory_evaluate)
  lvi_problem_desc trajectoryDesc() const {
    lvi_problem_desc d{};
    traj_->require_common_grid();
    d.t0 = traj_->R3Spline()->t0(); d.dt = traj_->R3Spline()->dt();
    d.n_knots = static_cast<int32_t>(std::min(traj_->R3Spline()->NumKnots(), traj_->SO3Spline()->NumKnots()));
    d.r3_knots = traj_->R3Spline()->data().data(); d.so3_knots = traj_->SO3Spline()->data().data();
    d.lidar_q = calib_param_manager->q_LtoI.c; d.lidar_p = calib_param_manager->p_LinI.v; d.lidar_toff = lidar_->time_offset();
    return d;
  }

 private:
  void solveLVI(const std::map<int64_t, std::shared_ptr<kontiki::sfm::View>>& frames, SurfelAssociation::Ptr surfels_association,
                const std::map<kontiki::sfm::Landmark*, size_t>& lm_surfel, bool with_lm_surfel, bool opt_time_offset_, bool lock_traj_and_lidar) {
    camera_->set_relative_orientation(calib_param_manager->q_CtoI);
    camera_->set_relative_position(calib_param_manager->p_CinI);
    camera_->LockRelativeOrientation(false);
    camera_->LockRelativePosition(false);
    lidar_->set_relative_orientation(calib_param_manager->q_LtoI);
    lidar_->set_relative_position(calib_param_manager->p_LinI);
    traj_->Lock(lock_traj_and_lidar);
    lidar_->LockRelativeOrientation(lock_traj_and_lidar);
    lidar_->LockRelativePosition(lock_traj_and_lidar);
    const bool free_offsets = opt_time_offset_ && time_offset_padding_ > 0;
    camera_->LockTimeOffset(!free_offsets);
    lidar_->LockTimeOffset(!free_offsets);
    imu_->LockGyroscopeBias(false);
    imu_->LockAccelerometerBias(false);

    auto estimator_split = std::make_shared<SplitTrajEstimator>(traj_);
    addGyroscopeMeasurements(estimator_split);
    addAccelerometerMeasurement(estimator_split);
    addSurfMeasurement(estimator_split, surfels_association);
    addVisualObservation(estimator_split, frames);
    if (with_lm_surfel) addVisualLidarMeasurement(estimator_split, lm_surfel);

    last_summary_ = estimator_split->Solve(80, false);
    std::cout << last_summary_.BriefReport() << std::endl;
    copyBack(/*camera=*/true, camera_->time_offset());
  }

  void copyBack(bool camera, double time_offset) {
    if (camera) {
      calib_param_manager->set_p_CinI(camera_->relative_position());
      calib_param_manager->set_q_CtoI(camera_->relative_orientation());
    }
    calib_param_manager->set_p_LinI(lidar_->relative_position());
    calib_param_manager->set_q_LtoI(lidar_->relative_orientation());
    calib_param_manager->set_time_offset(time_offset);
    calib_param_manager->set_gravity(imu_->refined_gravity());
    calib_param_manager->set_gyro_bias(imu_->gyroscope_bias());
    calib_param_manager->set_acce_bias(imu_->accelerometer_bias());
    calib_param_manager->showStates(false);
  }

  bool sensorPose(double traj_time, const Eigen::Quaterniond& q_StoI, const Eigen::Vector3d& p_SinI, Eigen::Quaterniond& q_StoG, Eigen::Vector3d& p_SinG) const {
    if (traj_->MinTime() > traj_time || traj_->MaxTime() <= traj_time) return false;
    Result result = traj_->Evaluate(traj_time, kontiki::trajectories::EvalOrientation | kontiki::trajectories::EvalPosition);
    q_StoG = result->orientation * q_StoI;
    p_SinG = result->orientation * p_SinI + result->position;
    return true;
  }
  bool relativeRotation(double traj_time1, double traj_time2, const Eigen::Quaterniond& q_StoI, Eigen::Quaterniond& q_S2toS1) const {
    assert(traj_time1 <= traj_time2);
    if (traj_->MinTime() > traj_time1 || traj_->MaxTime() <= traj_time2) return false;
    Result result1 = traj_->Evaluate(traj_time1, kontiki::trajectories::EvalOrientation), result2 = traj_->Evaluate(traj_time2, kontiki::trajectories::EvalOrientation);
    const Eigen::Quaterniond q_I2toI1 = result1->orientation.conjugate() * result2->orientation;
    q_S2toS1 = q_StoI.conjugate() * q_I2toI1 * q_StoI;
    return true;
  }

  // ---- the measurement feeders (:464-606): same filters, same constructor arguments
  template <typename TrajectoryModel> void addGyroscopeMeasurements(std::shared_ptr<kontiki::TrajectoryEstimator<TrajectoryModel>> estimator) {
    gyro_list_.clear();
    const double weight = calib_param_manager->global_opt_gyro_weight;
    const double min_time = estimator->trajectory()->MinTime(), max_time = estimator->trajectory()->MaxTime();
    for (const auto& v : imu_data_) {
      if (min_time > v.timestamp || max_time <= v.timestamp) continue;
      auto mg = std::make_shared<GyroMeasurement>(imu_, v.timestamp, v.gyro, weight);
      gyro_list_.push_back(mg);
      estimator->template AddMeasurement<GyroMeasurement>(mg);
    }
  }
  template <typename TrajectoryModel> void addAccelerometerMeasurement(std::shared_ptr<kontiki::TrajectoryEstimator<TrajectoryModel>> estimator) {
    accel_list_.clear();
    const double weight = calib_param_manager->global_opt_acce_weight;
    const double min_time = estimator->trajectory()->MinTime(), max_time = estimator->trajectory()->MaxTime();
    for (auto const& v : imu_data_) {
      if (min_time > v.timestamp || max_time <= v.timestamp) continue;
      auto ma = std::make_shared<AccelMeasurement>(imu_, v.timestamp, v.accel, weight);
      accel_list_.push_back(ma);
      estimator->template AddMeasurement<AccelMeasurement>(ma);
    }
  }
  template <typename TrajectoryModel> void addSurfMeasurement(std::shared_ptr<kontiki::TrajectoryEstimator<TrajectoryModel>> estimator, const SurfelAssociation::Ptr surfel_association) {
    const double weight = calib_param_manager->global_opt_lidar_weight;
    surfelpoint_list_.clear();
    closest_point_vec_.clear();
    for (auto const& v : surfel_association->get_surfel_planes()) closest_point_vec_.push_back(v.Pi);
    map_time_ = surfel_association->get_maptime();
    for (auto const& spoint : surfel_association->get_surfel_points()) {
      auto msp = std::make_shared<SurfMeasurement>(lidar_, spoint.point, closest_point_vec_.at(spoint.plane_id).data(), spoint.timestamp, map_time_, 5.0, weight);
      surfelpoint_list_.push_back(msp);
      estimator->template AddMeasurement<SurfMeasurement>(msp);
    }
  }
  template <typename TrajectoryModel> void addVisualObservation(std::shared_ptr<kontiki::TrajectoryEstimator<TrajectoryModel>> estimator,
                                                                const std::map<int64_t, std::shared_ptr<kontiki::sfm::View>>& frames) {
    landmark_list_.clear();
    const double min_time = estimator->trajectory()->MinTime(), max_time = estimator->trajectory()->MaxTime();
    const double weight = calib_param_manager->global_opt_cam_weight;
    for (auto iter = frames.begin(); iter != frames.end(); iter++) {
      if (min_time > iter->second->t0() || max_time <= iter->second->t0()) continue;
      for (const auto& obs : iter->second->observations()) {
        if (obs->landmark()->observations().size() <= 5) continue;
        if (obs->landmark()->inverse_depth() <= 0) continue;
        auto ms_obs = std::make_shared<CameraMeasurement>(camera_, obs, weight);   // 3-argument form: `weight` is the Huber threshold (Q3)
        landmark_list_.push_back(ms_obs);
        estimator->template AddMeasurement<CameraMeasurement>(ms_obs);
      }
    }
  }
  template <typename TrajectoryModel> void addVisualLidarMeasurement(std::shared_ptr<kontiki::TrajectoryEstimator<TrajectoryModel>> estimator,
                                                                     const std::map<kontiki::sfm::Landmark*, size_t>& lm_surfel) {
    const double weight = calib_param_manager->global_opt_visual_surfel_weight;
    surf_landmark_list_.clear();
    for (const auto& ls : lm_surfel) {
      const double time = ls.first->reference()->view()->t0();
      auto lsm = std::make_shared<CameraSurfMeasurement>(camera_, lidar_, ls.first, closest_point_vec_.at(ls.second).data(), time, map_time_, 5.0, weight);
      surf_landmark_list_.push_back(lsm);
      estimator->template AddMeasurement<CameraSurfMeasurement>(lsm);
    }
  }

  double time_offset_padding_, map_time_;
  std::shared_ptr<kontiki::trajectories::SplitTrajectory> traj_;
  std::shared_ptr<IMUSensor> imu_;
  std::shared_ptr<LiDARSensor> lidar_;
  std::shared_ptr<CameraSensor> camera_;
  CalibParamManager::Ptr calib_param_manager;
  Eigen::aligned_vector<IO::IMUData> imu_data_;
  Eigen::aligned_vector<Eigen::Vector3d> closest_point_vec_;
  std::vector<std::shared_ptr<GyroMeasurement>> gyro_list_;
  std::vector<std::shared_ptr<AccelMeasurement>> accel_list_;
  std::vector<std::shared_ptr<SurfMeasurement>> surfelpoint_list_;
  std::vector<std::shared_ptr<CameraMeasurement>> landmark_list_;
  std::vector<std::shared_ptr<CameraSurfMeasurement>> surf_landmark_list_;
  ceres::Solver::Summary last_summary_;
};

// SurfelAssociation::associateVisualPointsWithPlanes (L/src/core/surfel_association.cpp:161-214): the 3-D position of every landmark with
// inverse depth >= 0.05 in the L0 (map) frame from the camera pose at its reference view, then the box + 2 x radius plane test over every
// surfel on the device; the LAST matching plane wins (:206).
inline void SurfelAssociation::associateVisualPointsWithPlanes(std::shared_ptr<TrajectoryManagerLVI> traj_manager, const Eigen::Quaterniond& q_LtoC,
                                                               const Eigen::Vector3d& t_LinC,
                                                               const std::map<int64_t, std::shared_ptr<kontiki::sfm::Landmark>>& landmarks,
                                                               std::map<kontiki::sfm::Landmark*, size_t>& lm_splane) {
  Eigen::Quaterniond q_CtoG, q_L0_G;
  Eigen::Vector3d p_CinG, t_L0_G;
  if (!traj_manager->evaluateCameraPose(map_timestamp_, q_CtoG, p_CinG)) return;
  q_L0_G = q_CtoG * q_LtoC;
  t_L0_G = q_CtoG * t_LinC + p_CinG;
  auto camera = traj_manager->getCameraModel();
  std::vector<double> pts;
  std::vector<kontiki::sfm::Landmark*> who;
  for (auto it = landmarks.begin(); it != landmarks.end(); it++) {
    if (!it->second->reference()) continue;
    if (it->second->inverse_depth() < 0.05) continue;   // beyond 20 m
    Eigen::Vector3d p3d_C = camera->Unproject(it->second->reference()->uv());
    p3d_C = p3d_C / it->second->inverse_depth();
    const double timestamp = it->second->reference()->view()->t0();
    if (!traj_manager->evaluateCameraPose(timestamp, q_CtoG, p_CinG)) continue;
    const Eigen::Vector3d p3d_G = q_CtoG * p3d_C + p_CinG;
    const Eigen::Vector3d p3d_L0 = q_L0_G.inverse() * (p3d_G - t_L0_G);
    pts.insert(pts.end(), {p3d_L0[0], p3d_L0[1], p3d_L0[2]});
    who.push_back(it->second.get());
  }
  if (who.empty() || !set_) return;
  std::vector<int32_t> plane(who.size(), -1);
  lvi_exc_b200::throw_status(lvi_associate_landmarks(lvi_exc_b200::DefaultContext(), set_, pts.data(), static_cast<int64_t>(who.size()), associated_radius_, plane.data()));
  for (size_t k = 0; k < who.size(); ++k)
    if (plane[k] >= 0) lm_splane[who[k]] = static_cast<size_t>(plane[k]);
}
}  // namespace licalib
#endif
