// drop_in_check.cpp — the reference's own call sequences, compiled against include/lvi_exc_b200/compat under the reference's include paths.
// It drives the calibration the way LIinitializer does (T:539-708,1169-1300): S0 initialSO3TrajWithGyro -> Mapping (undistortScan, feedScan with
// LOAM poses, key scans into the NDT target) -> setSurfelMap / per-scan getAssociation / averageTimeDownSmaple -> S1 trajInitFromSurfel ->
// S4 trajInitFromLVIdata -> associateVisualPointsWithPlanes -> S5 trajInitFromLVIdata with the landmark-surfel residuals, every solve and every map
// step on the GPU.  Input: a binary blob written by tests/test_cpp_drop_in.py (synthetic sequence); output: one JSON line that the test compares
// with the CPU oracle running the same stage sequence.
#include <pclomp/ndt_omp.h>
#include <kontiki/trajectory_estimator.h>
#include <kontiki/trajectories/split_trajectory.h>
#include <kontiki/sfm/sfm.h>
#include <core/lidar_odometry.h>
#include <core/surfel_association.h>
#include <core/trajectory_manager_lvi.h>

#include <cstdio>
#include <cstring>
#include <fstream>

using namespace licalib;

namespace {
struct Blob {
  int32_t S, H, W, n_imu, n_views, n_obs, n_lm;
  double map_time, end_time;
  double q_LtoI[4], p_LinI[3], q_CtoI[4], p_CinI[3];
  std::vector<TPoint> raw;
  std::vector<double> poses, imu_t, gyro, accel, view_t0, obs_uv, lm_rho;
  std::vector<int32_t> obs_view, obs_lm, lm_ref;
};
template <class T> void rd(std::ifstream& f, T* p, size_t n) { f.read(reinterpret_cast<char*>(p), static_cast<std::streamsize>(sizeof(T) * n)); }
template <class T> void rdv(std::ifstream& f, std::vector<T>& v, size_t n) { v.resize(n); if (n) rd(f, v.data(), n); }
Blob load(const char* path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open input blob");
  Blob b;
  int32_t hdr[8];
  rd(f, hdr, 8);
  if (hdr[0] != 0x4C564931) throw std::runtime_error("bad blob magic");
  b.S = hdr[1]; b.H = hdr[2]; b.W = hdr[3]; b.n_imu = hdr[4]; b.n_views = hdr[5]; b.n_obs = hdr[6]; b.n_lm = hdr[7];
  rd(f, &b.map_time, 1); rd(f, &b.end_time, 1);
  rd(f, b.q_LtoI, 4); rd(f, b.p_LinI, 3); rd(f, b.q_CtoI, 4); rd(f, b.p_CinI, 3);
  rdv(f, b.raw, static_cast<size_t>(b.S) * b.H * b.W);
  rdv(f, b.poses, 16 * static_cast<size_t>(b.S));
  rdv(f, b.imu_t, b.n_imu); rdv(f, b.gyro, 3 * static_cast<size_t>(b.n_imu)); rdv(f, b.accel, 3 * static_cast<size_t>(b.n_imu));
  rdv(f, b.view_t0, b.n_views);
  rdv(f, b.obs_view, b.n_obs); rdv(f, b.obs_lm, b.n_obs); rdv(f, b.obs_uv, 2 * static_cast<size_t>(b.n_obs));
  rdv(f, b.lm_ref, b.n_lm); rdv(f, b.lm_rho, b.n_lm);
  if (!f) throw std::runtime_error("short input blob");
  return b;
}
std::string g_json;   // the result line is assembled here and printed last (the stages themselves print to stdout, as in the reference)
template <class... A> void jprintf(const char* fmt, A... a) {
  char buf[4096];
  std::snprintf(buf, sizeof(buf), fmt, a...);
  g_json += buf;
}
void stage_json(const char* name, const ceres::Solver::Summary& s, bool last = false) {
  jprintf("{\"name\": \"%s\", \"iterations\": %d, \"initial_cost\": %.17g, \"final_cost\": %.17g, \"residuals\": %d}%s", name, static_cast<int>(s.iterations.size()) - 1,
              s.initial_cost, s.final_cost, s.num_residuals, last ? "" : ", ");
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: drop_in_check <blob> [per_scan]\n"); return 2; }
  const bool per_scan = argc > 2 && std::strcmp(argv[2], "per_scan") == 0;
  try {
    const Blob b = load(argv[1]);
    const double ndt_resolution = 0.5, associated_radius = 0.05, knot_distance = 0.02, time_offset_padding = 0.2;
    double plane_lambda = 0.6;

    // ---- LIinitializer setup: trajectory manager, IMU feed, initial extrinsics (stand-in for the out-of-scope initial-guess stage)
    auto traj_manager = std::make_shared<TrajectoryManagerLVI>(CameraIntrinsic(), b.map_time, b.end_time, knot_distance, time_offset_padding);
    auto calib = traj_manager->getCalibParamManager();
    calib->set_q_LtoI(Eigen::Quaterniond(b.q_LtoI[3], b.q_LtoI[0], b.q_LtoI[1], b.q_LtoI[2]));
    calib->set_p_LinI(Eigen::Vector3d(b.p_LinI[0], b.p_LinI[1], b.p_LinI[2]));
    calib->set_q_CtoI(Eigen::Quaterniond(b.q_CtoI[3], b.q_CtoI[0], b.q_CtoI[1], b.q_CtoI[2]));
    calib->set_p_CinI(Eigen::Vector3d(b.p_CinI[0], b.p_CinI[1], b.p_CinI[2]));
    for (int i = 0; i < b.n_imu; ++i) {
      IO::IMUData d;
      d.timestamp = b.imu_t[i];
      d.gyro = Eigen::Vector3d(b.gyro[3 * i], b.gyro[3 * i + 1], b.gyro[3 * i + 2]);
      d.accel = Eigen::Vector3d(b.accel[3 * i], b.accel[3 * i + 1], b.accel[3 * i + 2]);
      traj_manager->feedIMUData(d);
    }
    g_json += "{\"stages\": [";

    // ---- S0
    traj_manager->initialSO3TrajWithGyro();
    stage_json("S0_so3", traj_manager->lastSummary());

    // ---- Mapping(): ScanUndistortion::undistortScan (rotation-only de-skew to each scan's own stamp) + LiDAROdometry::feedScan with LOAM poses
    const size_t npts = static_cast<size_t>(b.H) * b.W;
    std::vector<double> own_stamp(b.S);
    for (int s = 0; s < b.S; ++s) own_stamp[s] = b.raw[s * npts].timestamp;
    std::vector<VPoint> deskewed(b.raw.size());
    {
      lvi_problem_desc traj = traj_manager->trajectoryDesc();
      int32_t bad = 0;
      lvi_exc_b200::throw_status(lvi_undistort(lvi_exc_b200::DefaultContext(), &traj, reinterpret_cast<const lvi_point_xyzit*>(b.raw.data()), b.S, static_cast<int64_t>(npts),
                                               own_stamp.data(), /*correct_position=*/0, deskewed.data(), &bad));
      if (bad) throw std::runtime_error("scan stamps outside the trajectory");
    }
    auto lidar_odom = std::make_shared<LiDAROdometry>(ndt_resolution);
    std::vector<VPointCloud::Ptr> scans_in_map(b.S);
    std::vector<TPointCloud::Ptr> scans_raw(b.S);
    for (int s = 0; s < b.S; ++s) {
      VPointCloud::Ptr scan(new VPointCloud());
      scan->points.assign(deskewed.begin() + s * npts, deskewed.begin() + (s + 1) * npts);
      scan->width = b.W; scan->height = b.H; scan->is_dense = false;
      Eigen::Matrix4d pose;
      for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) pose(r, c) = b.poses[16 * s + 4 * r + c];
      lidar_odom->feedScan(own_stamp[s], scan, pose, /*update_map=*/true, /*using_loam=*/true);
      // ScanUndistortion::undistortScanInMap(odom_data_map): the de-skewed scan in the map frame
      scans_in_map[s] = VPointCloud::Ptr(new VPointCloud());
      pcl::transformPointCloud(*scan, *scans_in_map[s], pose);
      scans_raw[s] = TPointCloud::Ptr(new TPointCloud());
      scans_raw[s]->points.assign(b.raw.begin() + s * npts, b.raw.begin() + (s + 1) * npts);
      scans_raw[s]->width = b.W; scans_raw[s]->height = b.H; scans_raw[s]->is_dense = false;
    }

    // ---- DataAssociation() (T:1169-1210)
    auto surfel_association = std::make_shared<SurfelAssociation>(associated_radius, plane_lambda);
    surfel_association->setSurfelMap(lidar_odom->getNDTPtr(), b.map_time);
    if (per_scan)
      for (int s = 0; s < b.S; ++s) surfel_association->getAssociation(scans_in_map[s], scans_raw[s], 2);
    else
      surfel_association->getAssociationBatch(scans_in_map, scans_raw, 2);
    surfel_association->averageTimeDownSmaple();
    const size_t n_leaves = lidar_odom->getNDTPtr()->getTargetCells().getLeaves().size();

    // ---- S1 BatchOptimization()
    traj_manager->trajInitFromSurfel(surfel_association, false);
    stage_json("S1_surfel", traj_manager->lastSummary());

    // ---- ORB track: views, landmarks, observations (T:340-520 builds these from the ORB-SLAM text files)
    std::map<int64_t, std::shared_ptr<kontiki::sfm::View>> frames;
    std::map<int64_t, std::shared_ptr<kontiki::sfm::Landmark>> landmarks;
    std::vector<std::shared_ptr<kontiki::sfm::View>> views(b.n_views);
    std::vector<std::shared_ptr<kontiki::sfm::Landmark>> lms(b.n_lm);
    std::vector<std::shared_ptr<kontiki::sfm::Observation>> obs(b.n_obs);
    for (int v = 0; v < b.n_views; ++v) { views[v] = std::make_shared<kontiki::sfm::View>(v, b.view_t0[v]); frames[v] = views[v]; }
    for (int l = 0; l < b.n_lm; ++l) { lms[l] = std::make_shared<kontiki::sfm::Landmark>(); lms[l]->set_inverse_depth(b.lm_rho[l]); landmarks[l] = lms[l]; }
    for (int o = 0; o < b.n_obs; ++o) obs[o] = views[b.obs_view[o]]->CreateObservation(lms[b.obs_lm[o]], Eigen::Vector2d(b.obs_uv[2 * o], b.obs_uv[2 * o + 1]));
    for (int l = 0; l < b.n_lm; ++l) if (b.lm_ref[l] >= 0) lms[l]->set_reference(obs[b.lm_ref[l]]);

    // ---- S4: IMU + surfel + camera residuals, everything free
    traj_manager->trajInitFromLVIdata(frames, surfel_association, false, false);
    stage_json("S4_lvi", traj_manager->lastSummary());

    // ---- landmark -> surfel association, then S5 with the trajectory and the LiDAR locked
    const Eigen::Quaterniond q_LtoC = calib->q_CtoI.conjugate() * calib->q_LtoI;
    const Eigen::Vector3d t_LinC = calib->q_CtoI.conjugate() * (calib->p_LinI - calib->p_CinI);
    std::map<kontiki::sfm::Landmark*, size_t> lm_surfel;
    std::map<int64_t, std::shared_ptr<kontiki::sfm::Landmark>> with_ref;
    for (auto& kv : landmarks) if (kv.second->reference()) with_ref.insert(kv);
    surfel_association->associateVisualPointsWithPlanes(traj_manager, q_LtoC, t_LinC, with_ref, lm_surfel);
    traj_manager->trajInitFromLVIdata(frames, surfel_association, lm_surfel, false, true);
    stage_json("S5_lvi_surfel", traj_manager->lastSummary(), true);

    // one measurement's evaluators, as printErrorStatistics uses them (L/src/core/trajectory_manager_lvi.cpp:621-697)
    double p2p = 0;
    {
      const auto& sp = surfel_association->get_surfel_points().front();
      Eigen::Vector3d Pi = surfel_association->get_surfel_planes().at(sp.plane_id).Pi;
      auto lidar = traj_manager->getLidarModel();
      kontiki::measurements::LiDARSurfelPoint<kontiki::sensors::VLP16LiDAR> m(lidar, sp.point, Pi.data(), sp.timestamp, b.map_time, 5.0, 1.0);
      p2p = m.point2plane<kontiki::trajectories::SplitTrajectory>(*traj_manager->getTrajectory())(0);
    }
    Eigen::Quaterniond q_LtoG; Eigen::Vector3d p_LinG;
    const bool pose_ok = traj_manager->evaluateLidarPose(b.map_time + 0.5, q_LtoG, p_LinG), pose_out = traj_manager->evaluateLidarPose(b.end_time + 10.0, q_LtoG, p_LinG);

    jprintf("], \"n_leaves\": %zu, \"n_planes\": %zu, \"n_all\": %zu, \"n_surfel_points\": %zu, \"n_lm_plane\": %zu, \"point2plane\": %.9g, \"pose_ok\": %d, \"pose_out\": %d, "
                "\"q_LtoI\": [%.17g, %.17g, %.17g, %.17g], \"p_LinI\": [%.17g, %.17g, %.17g], \"q_CtoI\": [%.17g, %.17g, %.17g, %.17g], \"p_CinI\": [%.17g, %.17g, %.17g]}\n",
                n_leaves, surfel_association->get_surfel_planes().size(), surfel_association->get_all_surfel_points().size(), surfel_association->get_surfel_points().size(),
                lm_surfel.size(), p2p, pose_ok ? 1 : 0, pose_out ? 1 : 0, calib->q_LtoI.x(), calib->q_LtoI.y(), calib->q_LtoI.z(), calib->q_LtoI.w(), calib->p_LinI(0),
                calib->p_LinI(1), calib->p_LinI(2), calib->q_CtoI.x(), calib->q_CtoI.y(), calib->q_CtoI.z(), calib->q_CtoI.w(), calib->p_CinI(0), calib->p_CinI(1), calib->p_CinI(2));
    std::printf("%s", g_json.c_str());
  } catch (const std::exception& e) {
    std::printf("\n{\"error\": \"%s\"}\n", e.what());
    return 1;
  }
  return 0;
}
