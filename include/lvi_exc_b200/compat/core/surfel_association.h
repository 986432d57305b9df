// core/surfel_association.h — licalib::SurfelAssociation (L/include/core/surfel_association.h:36-140, L/src/core/surfel_association.cpp:42-331)
// with the reference's public interface, over the C-ABI of the CUDA library:
//   setSurfelMap(ndt, t)          -> lvi_surfel_extract on the NDT target cells that are already on the device (planarity test, RANSAC plane,
//                                    box of the leaf's points), lvi_surfel_export fills surfel_planes_
//   getAssociation(inM, raw, k)   -> lvi_associate for ONE organised scan (the reference's per-scan call, T:1191-1197);
//                                    getAssociationBatch takes all scans in one call (what a maintainer switches the loop to)
//   averageTimeDownSmaple / averageDownSmaple -> host-side index arithmetic on the points that came back (:216-244)
//   associateVisualPointsWithPlanes -> landmark positions from the trajectory manager's pose queries, lvi_associate_landmarks for the box +
//                                    distance test (last matching plane wins, :196-210)
// SurfelPlane::cloud / cloud_inlier (visualisation copies of the leaf's points) are not materialised; boxMin / boxMax / p4 / Pi are.
#ifndef LVI_EXC_B200_COMPAT_CORE_SURFEL_ASSOCIATION_H
#define LVI_EXC_B200_COMPAT_CORE_SURFEL_ASSOCIATION_H
#include <algorithm>
#include <iostream>
#include <map>
#include <memory>
#include <vector>

#include "../kontiki/kontiki_b200.h"
#include "../pcl/pcl_b200.h"
#include "../pclomp/ndt_omp.h"

namespace licalib {
class TrajectoryManagerLVI;

class SurfelAssociation {
 public:
  typedef std::shared_ptr<SurfelAssociation> Ptr;
  struct SurfelPoint {
    double timestamp;
    Eigen::Vector3d point;          // raw data
    Eigen::Vector3d point_in_map;
    size_t plane_id;
  };
  struct SurfelPlane {
    Eigen::Vector4d p4;
    Eigen::Vector3d Pi;             // closest-point parameterisation
    Eigen::Vector3d boxMin, boxMax;
    VPointCloud cloud, cloud_inlier;   // left empty (see the header comment)
    int n_inliers = 0;
    int64_t leaf_key = 0;
  };

  explicit SurfelAssociation(double associated_radius = 0.05, double plane_lambda = 0.7)
      : associated_radius_(associated_radius), p_lambda_(plane_lambda), map_timestamp_(0) {}
  ~SurfelAssociation() { release(); }
  SurfelAssociation(const SurfelAssociation&) = delete;
  SurfelAssociation& operator=(const SurfelAssociation&) = delete;

  void setPlaneLambda(double lambda) { p_lambda_ = lambda; }

  void setSurfelMap(const pclomp::NormalDistributionsTransform<VPoint, VPoint>::Ptr& ndtPtr, double timestamp = 0) {
    clearSurfelMap();
    map_timestamp_ = timestamp;
    ndt_ = ndtPtr;
    lvi_ctx* ctx = lvi_exc_b200::DefaultContext();
    // leaf.nr_points >= 10, planarity >= p_lambda_, RANSAC plane at 0.05 m with >= 20 inliers (:63-72, 276-290)
    lvi_exc_b200::throw_status(lvi_surfel_extract(ctx, ndtPtr->device_map(), p_lambda_, 10, 0.05f, 20, &set_));
    const int64_t P = lvi_surfel_count(set_);
    std::vector<double> p4(4 * P), Pi(3 * P), bmin(3 * P), bmax(3 * P);
    std::vector<int64_t> key(P);
    std::vector<int32_t> ninl(P);
    if (P) lvi_exc_b200::throw_status(lvi_surfel_export(ctx, set_, p4.data(), Pi.data(), bmin.data(), bmax.data(), key.data(), ninl.data()));
    surfel_planes_.resize(P);
    for (int64_t k = 0; k < P; ++k) {
      SurfelPlane& s = surfel_planes_[k];
      s.p4 = Eigen::Vector4d(p4[4 * k], p4[4 * k + 1], p4[4 * k + 2], p4[4 * k + 3]);
      s.Pi = Eigen::Vector3d(Pi[3 * k], Pi[3 * k + 1], Pi[3 * k + 2]);
      s.boxMin = Eigen::Vector3d(bmin[3 * k], bmin[3 * k + 1], bmin[3 * k + 2]);
      s.boxMax = Eigen::Vector3d(bmax[3 * k], bmax[3 * k + 1], bmax[3 * k + 2]);
      s.n_inliers = ninl[k]; s.leaf_key = key[k];
    }
    spoint_per_surfel_.resize(surfel_planes_.size());
    std::cout << "Plane number: " << surfel_planes_.size() << std::endl;
  }

  void getAssociation(const VPointCloud::Ptr& scan_inM, const TPointCloud::Ptr& scan_raw, size_t selected_num_per_ring = 2) {
    associate(scan_inM->points.data(), scan_raw->points.data(), 1, static_cast<int32_t>(scan_raw->width), static_cast<int32_t>(scan_raw->height), selected_num_per_ring);
  }
  // every scan of the sequence in one device call; scans are organised W x H clouds of equal size, concatenated in time order
  void getAssociationBatch(const std::vector<VPointCloud::Ptr>& scans_inM, const std::vector<TPointCloud::Ptr>& scans_raw, size_t selected_num_per_ring = 2) {
    if (scans_inM.empty()) return;
    const uint32_t W = scans_raw[0]->width, H = scans_raw[0]->height;
    std::vector<VPoint> in_map;
    std::vector<TPoint> raw;
    in_map.reserve(scans_inM.size() * W * H); raw.reserve(in_map.capacity());
    for (size_t s = 0; s < scans_inM.size(); ++s) {
      if (scans_raw[s]->width != W || scans_raw[s]->height != H || scans_inM[s]->size() != static_cast<size_t>(W) * H) throw std::invalid_argument("getAssociationBatch: scans differ in size");
      in_map.insert(in_map.end(), scans_inM[s]->points.begin(), scans_inM[s]->points.end());
      raw.insert(raw.end(), scans_raw[s]->points.begin(), scans_raw[s]->points.end());
    }
    associate(in_map.data(), raw.data(), static_cast<int32_t>(scans_inM.size()), static_cast<int32_t>(W), static_cast<int32_t>(H), selected_num_per_ring);
  }

  // :161-214; defined after TrajectoryManagerLVI (core/trajectory_manager_lvi.h)
  inline void associateVisualPointsWithPlanes(std::shared_ptr<TrajectoryManagerLVI> traj_manager, const Eigen::Quaterniond& q_LtoC, const Eigen::Vector3d& t_LinC,
                                              const std::map<int64_t, std::shared_ptr<kontiki::sfm::Landmark>>& landmarks,
                                              std::map<kontiki::sfm::Landmark*, size_t>& lm_splane);

  void averageDownSmaple(int num_points_max = 5) {   // :227-238
    for (const auto& v : spoint_per_surfel_) {
      if (v.size() < 20) continue;
      const int d_step = static_cast<int>(v.size()) / num_points_max;
      const int step = d_step > 1 ? d_step : 1;
      for (size_t i = 0; i < v.size(); i += step) spoint_downsampled_.push_back(v.at(i));
    }
  }
  void averageTimeDownSmaple(int step = 10) {   // :240-244
    for (size_t idx = 0; idx < spoints_all_.size(); idx += step) spoint_downsampled_.push_back(spoints_all_.at(idx));
  }

  const Eigen::aligned_vector<SurfelPlane>& get_surfel_planes() const { return surfel_planes_; }
  const Eigen::aligned_vector<SurfelPoint>& get_surfel_points() const { return spoint_downsampled_; }
  const Eigen::aligned_vector<SurfelPoint>& get_all_surfel_points() const { return spoints_all_; }
  double get_maptime() const { return map_timestamp_; }
  lvi_surfel_set* device_surfels() const { return set_; }

 private:
  void release() { if (set_) { lvi_surfel_destroy(set_); set_ = nullptr; } }
  void clearSurfelMap() {
    release();
    surfel_planes_.clear(); spoint_per_surfel_.clear(); spoint_downsampled_.clear(); spoints_all_.clear();
  }
  void associate(const VPoint* in_map, const TPoint* raw, int32_t n_scans, int32_t W, int32_t H, size_t k) {
    if (!set_) throw std::logic_error("SurfelAssociation::getAssociation before setSurfelMap");
    static_assert(sizeof(TPoint) == sizeof(lvi_point_xyzit), "raw point layout");
    lvi_ctx* ctx = lvi_exc_b200::DefaultContext();
    std::vector<lvi_surfel_point> out(std::max<size_t>(1, std::min<size_t>(static_cast<size_t>(n_scans) * W * H, size_t(1) << 20)));
    int64_t n_out = 0, n_all = 0;
    // time_step 1: every associated point comes back in the reference's emission order; the decimation is averageTimeDownSmaple's
    int rc = lvi_associate(ctx, ndt_->device_map(), set_, in_map, sizeof(VPoint), reinterpret_cast<const lvi_point_xyzit*>(raw), n_scans, W, H, associated_radius_,
                           static_cast<int32_t>(k), 1, out.data(), static_cast<int64_t>(out.size()), &n_out, &n_all);
    if (rc == LVI_OK && n_out > static_cast<int64_t>(out.size())) {   // n_out is the full count even when it did not fit
      out.resize(static_cast<size_t>(n_out));
      rc = lvi_associate(ctx, ndt_->device_map(), set_, in_map, sizeof(VPoint), reinterpret_cast<const lvi_point_xyzit*>(raw), n_scans, W, H, associated_radius_,
                         static_cast<int32_t>(k), 1, out.data(), static_cast<int64_t>(out.size()), &n_out, &n_all);
    }
    lvi_exc_b200::throw_status(rc);
    for (int64_t i = 0; i < n_out; ++i) {
      SurfelPoint sp;
      sp.timestamp = out[i].timestamp;
      sp.point = Eigen::Vector3d(out[i].point[0], out[i].point[1], out[i].point[2]);
      sp.point_in_map = Eigen::Vector3d(out[i].point_in_map[0], out[i].point_in_map[1], out[i].point_in_map[2]);
      sp.plane_id = static_cast<size_t>(out[i].plane_id);
      spoint_per_surfel_.at(sp.plane_id).push_back(sp);
      spoints_all_.push_back(sp);
    }
  }

  double associated_radius_, p_lambda_, map_timestamp_;
  pclomp::NormalDistributionsTransform<VPoint, VPoint>::Ptr ndt_;
  lvi_surfel_set* set_ = nullptr;
  Eigen::aligned_vector<SurfelPlane> surfel_planes_;
  Eigen::aligned_vector<SurfelPoint> spoints_all_;
  Eigen::aligned_vector<Eigen::aligned_vector<SurfelPoint>> spoint_per_surfel_;
  Eigen::aligned_vector<SurfelPoint> spoint_downsampled_;
};
}  // namespace licalib
#endif
