// core/lidar_odometry.h — licalib::LiDAROdometry (L/include/core/lidar_odometry.h:30-110, L/src/core/lidar_odometry.cpp:26-135) as the
// calibration runs it: with using_loam = true (T:1287-1294) feedScan takes the pose as given, key scans are transformed into the map frame
// (pcl::transformPointCloud on the device) and appended to the NDT target, whose voxel grid is built on the B200 when its cells are read.
// NDT registration (using_loam = false) is not on the path and throws.
#ifndef LVI_EXC_B200_COMPAT_CORE_LIDAR_ODOMETRY_H
#define LVI_EXC_B200_COMPAT_CORE_LIDAR_ODOMETRY_H
#include <cmath>
#include <map>
#include <memory>
#include <vector>

#include "../pcl/pcl_b200.h"
#include "../pclomp/ndt_omp.h"

namespace licalib {
class LiDAROdometry {
 public:
  typedef std::shared_ptr<LiDAROdometry> Ptr;
  struct OdomData {
    double timestamp;
    Eigen::Matrix4d pose;   // scan to map
  };
  explicit LiDAROdometry(double ndtResolution = 0.5) : map_cloud_(new VPointCloud()) { ndt_omp_ = ndtInit(ndtResolution); }

  static pclomp::NormalDistributionsTransform<VPoint, VPoint>::Ptr ndtInit(double ndt_resolution) {   // lidar_odometry.cpp:32-43
    auto ndt_omp = pclomp::NormalDistributionsTransform<VPoint, VPoint>::Ptr(new pclomp::NormalDistributionsTransform<VPoint, VPoint>());
    ndt_omp->setResolution(static_cast<float>(ndt_resolution));
    ndt_omp->setNumThreads(4);
    ndt_omp->setNeighborhoodSearchMethod(pclomp::DIRECT7);
    ndt_omp->setTransformationEpsilon(1e-3);
    ndt_omp->setStepSize(0.01);
    ndt_omp->setMaximumIterations(50);
    return ndt_omp;
  }

  void feedScan(double timestamp, VPointCloud::Ptr cur_scan, Eigen::Matrix4d pose_predict = Eigen::Matrix4d::Identity(), const bool update_map = true,
                const bool using_loam = false) {
    if (!using_loam) throw std::logic_error("lvi_exc_b200: LiDAROdometry::feedScan with NDT registration is not on the calibration hot path (using_loam = true)");
    OdomData odom_cur;
    odom_cur.timestamp = timestamp;
    odom_cur.pose = pose_predict;
    odom_data_.push_back(odom_cur);
    odom_data_map_[static_cast<int64_t>(timestamp * 1e9)] = odom_cur;
    if (update_map) updateKeyScan(cur_scan, odom_cur);
  }
  void clearOdomData() { key_frame_index_.clear(); odom_data_.clear(); odom_data_map_.clear(); }
  void setTargetMap(VPointCloud::Ptr map_cloud_in) {
    map_cloud_->clear();
    pcl::copyPointCloud(*map_cloud_in, *map_cloud_);
    for (size_t i = 0; i < map_cloud_->size(); ++i) map_cloud_->points[i].intensity = map_cloud_in->points[i].intensity;
    ndt_omp_->setInputTarget(map_cloud_);
  }
  const VPointCloud::Ptr getTargetMap() { return map_cloud_; }
  const pclomp::NormalDistributionsTransform<VPoint, VPoint>::Ptr& getNDTPtr() const { return ndt_omp_; }
  const Eigen::aligned_vector<OdomData>& get_odom_data() const { return odom_data_; }
  const std::map<int64_t, OdomData>& get_odom_data_map() const { return odom_data_map_; }
  const std::vector<size_t>& getKeyFrameIndex() const { return key_frame_index_; }
  static inline double normalize_angle(double ang_degree) {   // lidar_odometry.h:95-102 (one wrap each way)
    if (ang_degree > 180) ang_degree -= 360;
    if (ang_degree < -180) ang_degree += 360;
    return ang_degree;
  }

 private:
  void updateKeyScan(const VPointCloud::Ptr& cur_scan, const OdomData& odom_data) {   // :89-105
    if (!checkKeyScan(odom_data)) return;
    VPointCloud::Ptr scan_in_target(new VPointCloud());
    pcl::transformPointCloud(*cur_scan, *scan_in_target, odom_data.pose);
    *map_cloud_ += *scan_in_target;
    ndt_omp_->setInputTarget(map_cloud_);   // deferred: the grid is built once, over the final map (pclomp/ndt_omp.h)
    key_frame_index_.push_back(odom_data_.size());
  }
  // :107-128 — first scan, or moved more than 0.2 m, or yaw / pitch / roll (degrees, mathutils::R2ypr) changed by more than 5
  bool checkKeyScan(const OdomData& odom_data) {
    const Eigen::Matrix4d& T = odom_data.pose;
    const Eigen::Vector3d position_now(T(0, 3), T(1, 3), T(2, 3));
    const double dist = (position_now - position_last_).norm();
    const double n0 = T(0, 0), n1 = T(1, 0), n2 = T(2, 0), o0 = T(0, 1), o1 = T(1, 1), a0 = T(0, 2), a1 = T(1, 2);
    const double y = std::atan2(n1, n0), p = std::atan2(-n2, n0 * std::cos(y) + n1 * std::sin(y));
    const double r = std::atan2(a0 * std::sin(y) - a1 * std::cos(y), -o0 * std::sin(y) + o1 * std::cos(y));
    const double kDeg = 180.0 / M_PI;
    const Eigen::Vector3d ypr(y * kDeg, p * kDeg, r * kDeg);
    bool turned = false;
    for (int i = 0; i < 3; ++i) {
      double d = ypr(i) - ypr_last_(i);
      d = normalize_angle(d);
      if (std::fabs(d) > 5.0) turned = true;
    }
    if (key_frame_index_.empty() || dist > 0.2 || turned) {
      position_last_ = position_now;
      ypr_last_ = ypr;
      return true;
    }
    return false;
  }

  pclomp::NormalDistributionsTransform<VPoint, VPoint>::Ptr ndt_omp_;
  VPointCloud::Ptr map_cloud_;
  std::vector<size_t> key_frame_index_;
  Eigen::aligned_vector<OdomData> odom_data_;
  std::map<int64_t, OdomData> odom_data_map_;
  Eigen::Vector3d position_last_, ypr_last_;   // function-level statics in the reference
};
}  // namespace licalib
#endif
