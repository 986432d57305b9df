// io/lvi_files.h — the text formats on either side of the hot path, in C++ for hosts of the drop-in boundary: what the
// upstream producers hand to lvi_init_orb_surfel and what it appends to its result file.  Host-side text I/O, no device call.
//
//   LOAM/<bag>.txt   "<stamp_ns> tx ty tz qw qx qy qz"  (aloam/src/laserMapping.cpp:890-900), read by LIinitializer::ReadPoseGT
//                    (lvi_exc/test/lvi_initialize_surfel_orb.cpp:458-516): every pose -> loam_poses; a pose becomes an integration frame when it
//                    turned >= 5 deg or moved >= 0.1 m from the last KEPT one
//   ORB/<bag>.txt    "FramePose <stamp_ns> tx ty tz qx qy qz qw" / "UV <keyframe_stamp_ns> (u v mappoint_id)*" /
//                    "MapPoint <id> x y z <ref_keyframe_stamp_ns>"  (lvi_exc/test/write_orb_slam_results.cpp:138-185), read by
//                    LIinitializer::LoadOrbResults (:337-455) with its filters: unknown reference keyframe, reference observation missing,
//                    border_filter_uv pixels on the REFERENCE observation only; inverse depth 1 / (z + 1e-15); the other observations attached in
//                    ascending keyframe stamp
//   result CSV       CalibParamManager::save_result (lvi_exc/include/core/calibration.hpp:140-153)
// The Python side of the same formats is lvi_exc_b200/formats.py; tests/test_formats.py reads the files of one with the other.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include <Eigen/Dense>

#include "../kontiki/kontiki_b200.h"

namespace lvi_io {

struct IntegrationFrame {   // lvi_initialize_surfel_orb.cpp:59-63
  int64_t timestamp = 0;    // ns
  Eigen::Matrix4d Tcw = Eigen::Matrix4d::Identity();
};

inline std::vector<std::string> SplitString(const std::string& s, char sep) {   // utils/string_utils: empty fields are dropped
  std::vector<std::string> out;
  std::string cur;
  for (char c : s) {
    if (c == sep || c == '\r' || c == '\n') { if (!cur.empty()) out.push_back(cur); cur.clear(); }
    else cur.push_back(c);
  }
  if (!cur.empty()) out.push_back(cur);
  return out;
}
inline int64_t to_i64(const std::string& s) { std::istringstream iss(s); int64_t v = 0; iss >> v; return v; }

inline Eigen::Matrix4d pose_of(const Eigen::Vector3d& t, const Eigen::Quaterniond& q) {
  Eigen::Matrix4d T = Eigen::Matrix4d::Identity();
  const Eigen::Matrix3d R = q.toRotationMatrix();
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T(r, c) = R(r, c); T(r, 3) = t(r); }
  return T;
}

// ReadPoseGT (:458-516).  Returns false when the file cannot be opened.
inline bool ReadPoseGT(const std::string& pose_file_path, std::vector<std::pair<double, Eigen::Matrix4d>>& loam_poses,
                       std::vector<IntegrationFrame>& integration_frames_lidar) {
  std::ifstream ifs(pose_file_path);
  if (!ifs.is_open()) return false;
  std::string line;
  Eigen::Vector3d last_pos;
  Eigen::Quaterniond last_ori;
  while (std::getline(ifs, line)) {
    const std::vector<std::string> words = SplitString(line, ' ');
    if (words.size() != 8) break;
    const int64_t stamp = to_i64(words[0]);
    const Eigen::Vector3d pos(atof(words[1].c_str()), atof(words[2].c_str()), atof(words[3].c_str()));
    const Eigen::Quaterniond rot(atof(words[4].c_str()), atof(words[5].c_str()), atof(words[6].c_str()), atof(words[7].c_str()));   // w x y z
    const Eigen::Matrix4d pose = pose_of(pos, rot);
    loam_poses.push_back({stamp * 1e-9, pose});
    if (!integration_frames_lidar.empty()) {
      if (last_ori.angularDistance(rot) * 180. / M_PI < 5. && (last_pos - pos).norm() < 0.1) continue;
    }
    IntegrationFrame frame;
    frame.timestamp = stamp;
    frame.Tcw = pose;
    integration_frames_lidar.push_back(frame);
    last_pos = pos;
    last_ori = rot;
  }
  return true;
}

// LoadOrbResults (:337-455).  Returns false when the file cannot be opened.
inline bool LoadOrbResults(const std::string& orb_res_path, int cam_rows, int cam_cols, int border_filter_uv, std::vector<IntegrationFrame>& integration_frames_cam,
                           std::map<int64_t, std::shared_ptr<kontiki::sfm::View>>& views_db,
                           std::map<int64_t, std::shared_ptr<kontiki::sfm::Landmark>>& landmark_db) {
  std::ifstream ifs(orb_res_path);
  if (!ifs.is_open()) return false;
  std::map<int64_t, std::map<int64_t, Eigen::Vector2d>> uv_points;
  std::string line;
  while (std::getline(ifs, line)) {
    const std::vector<std::string> tokens = SplitString(line, ' ');
    if (tokens.empty()) break;
    if (tokens[0] == "FramePose") {
      if (tokens.size() != 9) throw std::runtime_error("Wrong line!");
      IntegrationFrame frame;
      frame.timestamp = to_i64(tokens[1]);
      const Eigen::Vector3d t(atof(tokens[2].c_str()), atof(tokens[3].c_str()), atof(tokens[4].c_str()));
      const Eigen::Quaterniond q(atof(tokens[8].c_str()), atof(tokens[5].c_str()), atof(tokens[6].c_str()), atof(tokens[7].c_str()));   // file order x y z w
      frame.Tcw = pose_of(t, q);
      integration_frames_cam.push_back(frame);
    } else if (tokens[0] == "UV") {
      if ((tokens.size() - 2) % 3 != 0) throw std::runtime_error("Wrong line!");
      const int num = static_cast<int>((tokens.size() - 2) / 3);
      std::map<int64_t, Eigen::Vector2d> obs_pair;
      const int64_t frameid = to_i64(tokens[1]);
      for (int i = 0; i < num; ++i) {
        const int idx = 2 + i * 3;
        obs_pair[to_i64(tokens[idx + 2])] = Eigen::Vector2d(atof(tokens[idx].c_str()), atof(tokens[idx + 1].c_str()));
      }
      uv_points[frameid] = obs_pair;
      views_db[frameid] = std::make_shared<kontiki::sfm::View>(frameid, frameid * 1e-9);
    } else if (tokens[0] == "MapPoint") {
      if (tokens.size() != 6) throw std::runtime_error("Wrong line!");
      const int64_t lm_id = to_i64(tokens[1]), ref_id = to_i64(tokens[5]);
      const Eigen::Vector3d pos(atof(tokens[2].c_str()), atof(tokens[3].c_str()), atof(tokens[4].c_str()));   // in the reference keyframe's camera frame
      auto view_ref = views_db.find(ref_id);
      if (view_ref == views_db.end()) continue;
      if (uv_points.find(ref_id) == uv_points.end()) continue;
      auto it_uv = uv_points[ref_id].find(lm_id);
      if (it_uv == uv_points[ref_id].end()) continue;
      const Eigen::Vector2d euv = it_uv->second;
      const int border = border_filter_uv;
      if (euv[0] < border || euv[1] < border || euv[0] > cam_cols - border || euv[1] > cam_rows - border) continue;
      if (landmark_db.find(lm_id) != landmark_db.end()) continue;
      auto lm = std::make_shared<kontiki::sfm::Landmark>();
      lm->set_inverse_depth(1. / (pos[2] + 1e-15));
      auto ref_obs_ptr = view_ref->second->CreateObservation(lm, euv);
      lm->set_reference(ref_obs_ptr);
      landmark_db.insert({lm_id, lm});
      for (auto it = views_db.begin(); it != views_db.end(); it++) {
        if (it == view_ref) continue;
        auto it_obs = uv_points[it->first].find(lm_id);
        if (it_obs != uv_points[it->first].end()) it->second->CreateObservation(lm, it_obs->second);
      }
    }
  }
  return true;
}

// CalibParamManager::save_result (calibration.hpp:140-153): appends one line
inline void save_result(const std::string& filename, const std::string& info, const Eigen::Quaterniond& q_LtoI, const Eigen::Vector3d& p_LinI, double time_offset,
                        const Eigen::Vector3d& gravity, const Eigen::Vector3d& gyro_bias, const Eigen::Vector3d& acce_bias) {
  const Eigen::Quaterniond q_ItoL = q_LtoI.inverse();
  const Eigen::Vector3d p_IinL = q_ItoL * (-p_LinI);
  std::ofstream outfile;
  outfile.open(filename, std::ios::app);
  outfile << info << "," << p_IinL(0) << "," << p_IinL(1) << "," << p_IinL(2) << "," << q_ItoL.x() << "," << q_ItoL.y() << "," << q_ItoL.z() << "," << q_ItoL.w() << ","
          << time_offset << "," << gravity(0) << "," << gravity(1) << "," << gravity(2) << "," << gyro_bias(0) << "," << gyro_bias(1) << "," << gyro_bias(2) << ","
          << acce_bias(0) << "," << acce_bias(1) << "," << acce_bias(2) << "\n";
  outfile.close();
}

}  // namespace lvi_io
