// pclomp/ndt_omp.h — pclomp::NormalDistributionsTransform / VoxelGridCovariance as the hot path uses them
// (N/include/pclomp/ndt_omp.h:117-122,275-282 setInputTarget -> init -> VoxelGridCovariance::filter ; N/include/pclomp/voxel_grid_covariance_omp.h:92-300 ;
// callers L/src/core/lidar_odometry.cpp:32-43,102 and L/src/core/surfel_association.cpp:60-71).  setInputTarget hands the cloud to the
// CUDA voxel build (lvi_voxel_build); the leaves stay on the device and getLeaves() materialises the std::map view only when asked.
// The registration half (setInputSource / align / derivatives) is not on the path (using_loam = true, T:1287-1294) and throws.
#ifndef LVI_EXC_B200_COMPAT_PCLOMP_NDT_OMP_H
#define LVI_EXC_B200_COMPAT_PCLOMP_NDT_OMP_H
#include <map>
#include <memory>
#include <stdexcept>

#include "../kontiki/kontiki_b200.h"
#include "../pcl/pcl_b200.h"

namespace pclomp {
enum NeighborSearchMethod { KDTREE, DIRECT26, DIRECT7, DIRECT1 };

template <class PointT> class VoxelGridCovariance {
 public:
  struct Leaf {   // voxel_grid_covariance_omp.h:92-190
    int nr_points = 0;
    Eigen::Vector3d mean_;
    Eigen::Matrix3d cov_ = Eigen::Matrix3d::Identity(), icov_, evecs_ = Eigen::Matrix3d::Identity();
    Eigen::Vector3d evals_;
    pcl::PointCloud<PointT> pointList_;
    int getPointCount() const { return nr_points; }
    Eigen::Vector3d getMean() const { return mean_; }
    Eigen::Matrix3d getCov() const { return cov_; }
    Eigen::Matrix3d getInverseCov() const { return icov_; }
    Eigen::Matrix3d getEvecs() const { return evecs_; }
    Eigen::Vector3d getEvals() const { return evals_; }
  };
  using Map = std::map<size_t, Leaf>;
  ~VoxelGridCovariance() { reset(); }
  void reset() { if (map_) { lvi_voxel_destroy(map_); map_ = nullptr; } leaves_.clear(); leaves_valid_ = false; }
  void build(const typename pcl::PointCloud<PointT>::ConstPtr& cloud, float leaf) {
    reset();
    cloud_ = cloud; leaf_ = leaf;
    lvi_exc_b200::throw_status(lvi_voxel_build(lvi_exc_b200::DefaultContext(), cloud->points.data(), sizeof(PointT), static_cast<int64_t>(cloud->size()), leaf,
                                               min_points_per_voxel_, min_covar_eigvalue_mult_, &map_));
  }
  lvi_voxel_map* device_map() const { return map_; }
  Eigen::Vector3d getLeafSize() const { return Eigen::Vector3d(leaf_, leaf_, leaf_); }
  void setMinPointPerVoxel(int n) { min_points_per_voxel_ = n > 2 ? n : 3; }   // voxel_grid_covariance_omp.h:233-246
  int getMinPointPerVoxel() const { return min_points_per_voxel_; }
  void setCovEigValueInflationRatio(double r) { min_covar_eigvalue_mult_ = r; }
  // std::map<size_t, Leaf> in ascending voxel-index order, pointList_ in cloud order (voxel_grid_covariance_omp.h:388-392)
  const Map& getLeaves() const {
    if (leaves_valid_ || !map_) return leaves_;
    const int64_t L = lvi_voxel_num_leaves(map_), np = lvi_voxel_num_points(map_);
    std::vector<int64_t> keys(L), start(L + 1);
    std::vector<int32_t> npts(L), pidx(np);
    std::vector<double> mean(3 * L), cov(9 * L), evals(3 * L), evecs(9 * L), icov(9 * L);
    lvi_exc_b200::throw_status(lvi_voxel_export(lvi_exc_b200::DefaultContext(), map_, keys.data(), npts.data(), mean.data(), cov.data(), evals.data(), evecs.data(),
                                                icov.data(), start.data(), pidx.data()));
    for (int64_t l = 0; l < L; ++l) {
      Leaf& lf = leaves_[static_cast<size_t>(keys[l])];
      lf.nr_points = npts[l];
      lf.mean_ = Eigen::Vector3d(mean[3 * l], mean[3 * l + 1], mean[3 * l + 2]);
      lf.evals_ = Eigen::Vector3d(evals[3 * l], evals[3 * l + 1], evals[3 * l + 2]);
      for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { lf.cov_(r, c) = cov[9 * l + 3 * r + c]; lf.icov_(r, c) = icov[9 * l + 3 * r + c]; lf.evecs_(r, c) = evecs[9 * l + 3 * r + c]; }
      for (int64_t k = start[l]; k < start[l + 1]; ++k) lf.pointList_.push_back(cloud_->points[pidx[k]]);
    }
    leaves_valid_ = true;
    return leaves_;
  }
 private:
  lvi_voxel_map* map_ = nullptr;
  typename pcl::PointCloud<PointT>::ConstPtr cloud_;
  float leaf_ = 1.0f;
  int min_points_per_voxel_ = 6;             // voxel_grid_covariance_omp.h:205
  double min_covar_eigvalue_mult_ = 0.01;    // :206
  mutable Map leaves_;
  mutable bool leaves_valid_ = false;
};

template <class PointSource, class PointTarget> class NormalDistributionsTransform {
 public:
  using Ptr = std::shared_ptr<NormalDistributionsTransform<PointSource, PointTarget>>;
  using ConstPtr = std::shared_ptr<const NormalDistributionsTransform<PointSource, PointTarget>>;
  using PointCloudTargetConstPtr = typename pcl::PointCloud<PointTarget>::ConstPtr;
  using TargetGrid = VoxelGridCovariance<PointTarget>;
  void setResolution(float r) { if (resolution_ != r) { resolution_ = r; dirty_ = target_ != nullptr; } }
  float getResolution() const { return resolution_; }
  void setNumThreads(int n) { num_threads_ = n; }
  void setNeighborhoodSearchMethod(NeighborSearchMethod m) { search_method_ = m; }
  void setTransformationEpsilon(double e) { transformation_epsilon_ = e; }
  void setStepSize(double s) { step_size_ = s; }
  void setMaximumIterations(int n) { max_iterations_ = n; }
  // ndt_omp.h:117-122: the voxel grid is (re)built over the whole target.  LiDAROdometry calls this once per key scan with a growing
  // map (L/src/core/lidar_odometry.cpp:102); only the LAST target is ever read, so the build is deferred until the cells are asked for.
  void setInputTarget(const PointCloudTargetConstPtr& cloud) { target_ = cloud; dirty_ = true; }
  const TargetGrid& getTargetCells() const { ensure(); return target_cells_; }
  lvi_voxel_map* device_map() const { ensure(); return target_cells_.device_map(); }
  PointCloudTargetConstPtr getInputTarget() const { return target_; }
  template <class Cloud> void setInputSource(const Cloud&) { throw std::logic_error("lvi_exc_b200: NDT registration (setInputSource / align) is not on the calibration hot path (using_loam = true)"); }
 private:
  void ensure() const {
    if (!dirty_) return;
    if (!target_ || target_->empty()) throw std::runtime_error("pclomp::NormalDistributionsTransform: no input target");
    target_cells_.build(target_, resolution_);
    dirty_ = false;
  }
  float resolution_ = 1.0f;
  int num_threads_ = 1, max_iterations_ = 35;
  NeighborSearchMethod search_method_ = DIRECT7;
  double transformation_epsilon_ = 0.1, step_size_ = 0.1;
  PointCloudTargetConstPtr target_;
  mutable TargetGrid target_cells_;
  mutable bool dirty_ = false;
};
}  // namespace pclomp
#endif
