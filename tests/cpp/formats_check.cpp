// Reads the LOAM / ORB text files with include/lvi_exc_b200/compat/io/lvi_files.h and prints what it found (one JSON line), appends a result line.
// usage: formats_check <loam.txt> <orb.txt> <result.csv>
#include <cstdio>

#include <io/lvi_files.h>

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  std::vector<std::pair<double, Eigen::Matrix4d>> loam_poses;
  std::vector<lvi_io::IntegrationFrame> frames_lidar, frames_cam;
  if (!lvi_io::ReadPoseGT(argv[1], loam_poses, frames_lidar)) return 3;
  std::map<int64_t, std::shared_ptr<kontiki::sfm::View>> views_db;
  std::map<int64_t, std::shared_ptr<kontiki::sfm::Landmark>> landmark_db;
  if (!lvi_io::LoadOrbResults(argv[2], 720, 1280, 10, frames_cam, views_db, landmark_db)) return 4;
  printf("{\"n_loam\": %zu, \"key_stamps\": [", loam_poses.size());
  for (size_t i = 0; i < frames_lidar.size(); ++i) printf("%s%lld", i ? ", " : "", static_cast<long long>(frames_lidar[i].timestamp));
  printf("], \"last_pose\": [");
  const Eigen::Matrix4d& T = loam_poses.back().second;
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) printf("%s%.17g", (r || c) ? ", " : "", T(r, c));
  printf("], \"n_frames_cam\": %zu, \"cam_pose1\": [", frames_cam.size());
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) printf("%s%.17g", (r || c) ? ", " : "", frames_cam[1].Tcw(r, c));
  printf("], \"n_views\": %zu, \"landmarks\": [", views_db.size());
  bool first = true;
  for (auto& kv : landmark_db) {
    const auto obs = kv.second->observations();
    printf("%s{\"id\": %lld, \"rho\": %.17g, \"n_obs\": %zu, \"ref_uv\": [%.17g, %.17g], \"ref_t0\": %.17g}", first ? "" : ", ", static_cast<long long>(kv.first),
           kv.second->inverse_depth(), obs.size(), kv.second->reference()->uv()(0), kv.second->reference()->uv()(1), kv.second->reference()->view()->t0());
    first = false;
  }
  printf("]}\n");
  lvi_io::save_result(argv[3], "check", Eigen::Quaterniond(0.9, 0.1, -0.2, 0.3).normalized(), Eigen::Vector3d(0.05, -0.1, 0.08), 0.0, Eigen::Vector3d(0.1, 0.2, -9.7),
                      Eigen::Vector3d(0.002, -0.001, 0.0015), Eigen::Vector3d(0.03, -0.02, 0.01));
  return 0;
}
