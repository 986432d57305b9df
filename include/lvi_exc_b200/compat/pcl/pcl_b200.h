// pcl_b200.h — the slice of PCL the hot path's call sites touch (point layouts, PointCloud container, transformPointCloud, getMinMax3D),
// for hosts without PCL.  Layouts are PCL's: pcl::PointXYZI is 32 bytes with x,y,z at offset 0 and intensity at offset 16;
// licalib::PointXYZIT (L/include/utils/pcl_utils.h:39-58) adds a double timestamp at offset 24.  Heavy work on clouds goes to the device
// through include/lvi_exc_b200.h; what is here is container code.
#ifndef LVI_EXC_B200_COMPAT_PCL_H
#define LVI_EXC_B200_COMPAT_PCL_H
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <vector>

#include <Eigen/Dense>

#include "../context.h"

namespace pcl {
struct alignas(16) PointXYZI {
  float x = 0, y = 0, z = 0, _pad = 1.f;
  float intensity = 0, _pad2[3] = {0, 0, 0};
};
static_assert(sizeof(PointXYZI) == 32, "pcl::PointXYZI layout");

template <class PointT> class PointCloud {
 public:
  using Ptr = std::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
  std::vector<PointT> points;
  uint32_t width = 0, height = 0;
  bool is_dense = true;
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); width = height = 0; }
  void resize(size_t n) { points.resize(n); if (width * height != n) { width = static_cast<uint32_t>(n); height = 1; } }
  void push_back(const PointT& p) { points.push_back(p); width = static_cast<uint32_t>(points.size()); height = 1; }
  PointT& at(int column, int row) { return points.at(static_cast<size_t>(row) * width + column); }
  const PointT& at(int column, int row) const { return points.at(static_cast<size_t>(row) * width + column); }
  PointT& operator[](size_t i) { return points[i]; }
  const PointT& operator[](size_t i) const { return points[i]; }
  typename std::vector<PointT>::iterator begin() { return points.begin(); }
  typename std::vector<PointT>::iterator end() { return points.end(); }
  typename std::vector<PointT>::const_iterator begin() const { return points.begin(); }
  typename std::vector<PointT>::const_iterator end() const { return points.end(); }
  PointCloud& operator+=(const PointCloud& o) {
    points.insert(points.end(), o.points.begin(), o.points.end());
    width = static_cast<uint32_t>(points.size()); height = 1;
    if (!o.is_dense) is_dense = false;
    return *this;
  }
  Ptr makeShared() const { return Ptr(new PointCloud<PointT>(*this)); }
  bool isOrganized() const { return height > 1; }
};

template <class PointT> inline void getMinMax3D(const PointCloud<PointT>& cloud, PointT& mn, PointT& mx) {
  const float big = std::numeric_limits<float>::max();
  mn.x = mn.y = mn.z = big; mx.x = mx.y = mx.z = -big;
  for (const auto& p : cloud.points) {
    if (!cloud.is_dense && !(std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z))) continue;
    mn.x = std::min(mn.x, p.x); mn.y = std::min(mn.y, p.y); mn.z = std::min(mn.z, p.z);
    mx.x = std::max(mx.x, p.x); mx.y = std::max(mx.y, p.y); mx.z = std::max(mx.z, p.z);
  }
}
template <class A, class B> inline void copyPointCloud(const PointCloud<A>& in, PointCloud<B>& out) {
  out.points.resize(in.points.size());
  for (size_t i = 0; i < in.points.size(); ++i) { out.points[i].x = in.points[i].x; out.points[i].y = in.points[i].y; out.points[i].z = in.points[i].z; }
  out.width = in.width; out.height = in.height; out.is_dense = in.is_dense;
}
}  // namespace pcl

namespace licalib {
struct alignas(16) PointXYZIT {   // L/include/utils/pcl_utils.h:39-58
  float x = 0, y = 0, z = 0, _pad = 1.f;
  float intensity = 0, _pad2 = 0;
  double timestamp = 0;
};
static_assert(sizeof(PointXYZIT) == sizeof(lvi_point_xyzit), "licalib::PointXYZIT layout");
typedef pcl::PointXYZI VPoint;
typedef pcl::PointCloud<VPoint> VPointCloud;
typedef PointXYZIT TPoint;
typedef pcl::PointCloud<TPoint> TPointCloud;
}  // namespace licalib

namespace pcl {
// pcl::transformPointCloud with the pose cast to a float 4x4 (the reference passes Eigen::Matrix4d; PCL computes in float): on the device
// (lvi_transform_scans), bit-exact with Eigen's float evaluation order
inline void transformPointCloud(const PointCloud<PointXYZI>& in, PointCloud<PointXYZI>& out, const Eigen::Matrix4d& pose) {
  double row_major[16];
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) row_major[4 * r + c] = pose(r, c);
  PointCloud<PointXYZI> result;
  result.points.resize(in.points.size());
  result.width = in.width; result.height = in.height; result.is_dense = in.is_dense;
  if (!in.points.empty())
    lvi_exc_b200::throw_status(lvi_transform_scans(lvi_exc_b200::DefaultContext(), in.points.data(), 1, static_cast<int64_t>(in.points.size()), row_major, result.points.data()));
  out = std::move(result);   // in and out may be the same cloud
}
}
#endif
