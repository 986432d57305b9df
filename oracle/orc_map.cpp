// oracle/orc_map.cpp — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product library).
//
// CPU restatement of the LiDAR-map half of the LVI-ExC hot path:
//   * pclomp::VoxelGridCovariance::applyFilter        N/voxel_grid_covariance_omp_impl.hpp:49-374
//   * SurfelAssociation::setSurfelMap/checkPlaneType   L/src/core/surfel_association.cpp:50-108,246-266
//   * SurfelAssociation::fitPlane (pcl::SACSegmentation, NOT in the tree)  L/src/core/surfel_association.cpp:268-294
//   * SurfelAssociation::getAssociation / associateScanToSurfel / averageTimeDownSmaple  :111-159,240-244,305-331
// (paths relative to /root/reference/src, N/=ndt_omp/include/pclomp, L/=lvi_exc).
//
// PARITY STATUS: unpinned by upstream — the reference ships no tests or golden vectors and its
// PCL/Eigen dependencies are absent here.  Third-party arithmetic is restated as follows and documented in
// Eigen::SelfAdjointEigenSolver<Matrix3d> -> cyclic Jacobi (fp64, ascending); pcl RANSAC plane
// (PCL >= 1.7, `find_package(PCL 1.7)` N/../CMakeLists.txt:16) -> same loop structure (50 iterations max,
// p=0.99 adaptive k, threshold test `< thr`, PCA refinement + inlier re-selection) but with a counter-based
// RNG and double-accumulated, float-rounded PCA so that the result is independent of summation order.
// Compile with -ffp-contract=off: the float/double operation order below is part of the specification the
// CUDA kernels must reproduce bit-exactly.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <vector>

namespace {

struct Leaf {  // N/voxel_grid_covariance_omp.h:92-190
  int nr_points = 0;
  double mean[3] = {0, 0, 0};
  double cov[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // Q1: seeded with Identity (:97-106)
  double icov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  double evecs[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double evals[3] = {0, 0, 0};
  std::vector<int32_t> points;  // pointList_ as indices into the input cloud
};

struct VoxelMap {
  std::map<size_t, Leaf> leaves;  // N/voxel_grid_covariance_omp.h:198
  int min_b[3] = {0, 0, 0}, max_b[3] = {0, 0, 0}, div_b[3] = {0, 0, 0}, divb_mul[3] = {0, 0, 0};
  float inv_leaf = 0;
  const float* pts = nullptr;
  size_t stride = 0;  // in floats
  int64_t n = 0, n_binned = 0;
  int status = 0;
};

struct Plane {
  double p4[4], Pi[3], bmin[3], bmax[3];
  int64_t key;
  int32_t n_inliers;
};
struct SurfelSet { std::vector<Plane> planes; };

inline const float* P(const VoxelMap& m, int64_t i) { return m.pts + i * m.stride; }

// Cyclic Jacobi for a symmetric 3x3 given by its LOWER triangle (SelfAdjointEigenSolver reads the lower
// triangle, N/voxel_grid_covariance_omp_impl.hpp:337).  Eigenvalues ascending; evecs row-major with
// eigenvectors in columns.  Pure + - * / sqrt in a fixed order => reproducible bit-for-bit on the GPU.
void jacobi3(const double A[9], double evals[3], double evecs[9]) {
  double a00 = A[0], a11 = A[4], a22 = A[8], a01 = A[3], a02 = A[6], a12 = A[7];
  double V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int sweep = 0; sweep < 12; ++sweep) {
    if (a01 == 0.0 && a02 == 0.0 && a12 == 0.0) break;
    for (int pq = 0; pq < 3; ++pq) {
      double app, aqq, apq;
      if (pq == 0) { app = a00; aqq = a11; apq = a01; }
      else if (pq == 1) { app = a00; aqq = a22; apq = a02; }
      else { app = a11; aqq = a22; apq = a12; }
      if (apq == 0.0) continue;
      if (std::fabs(apq) < 1e-300) {  // flush denormal-scale couplings
        if (pq == 0) a01 = 0; else if (pq == 1) a02 = 0; else a12 = 0;
        continue;
      }
      const double theta = (aqq - app) / (2.0 * apq);
      const double at = std::fabs(theta);
      double t = 1.0 / (at + std::sqrt(theta * theta + 1.0));
      if (theta < 0.0) t = -t;
      const double c = 1.0 / std::sqrt(t * t + 1.0);
      const double s = t * c;
      const double napp = app - t * apq;
      const double naqq = aqq + t * apq;
      if (pq == 0) {  // rotate (0,1); third index r = 2
        const double ar_p = a02, ar_q = a12;
        a00 = napp; a11 = naqq; a01 = 0.0;
        a02 = c * ar_p - s * ar_q;
        a12 = s * ar_p + c * ar_q;
      } else if (pq == 1) {  // rotate (0,2); r = 1
        const double ar_p = a01, ar_q = a12;
        a00 = napp; a22 = naqq; a02 = 0.0;
        a01 = c * ar_p - s * ar_q;
        a12 = s * ar_p + c * ar_q;
      } else {  // rotate (1,2); r = 0
        const double ar_p = a01, ar_q = a02;
        a11 = napp; a22 = naqq; a12 = 0.0;
        a01 = c * ar_p - s * ar_q;
        a02 = s * ar_p + c * ar_q;
      }
      const int p = (pq == 2) ? 1 : 0, q = (pq == 0) ? 1 : 2;
      for (int r = 0; r < 3; ++r) {
        const double vp = V[r * 3 + p], vq = V[r * 3 + q];
        V[r * 3 + p] = c * vp - s * vq;
        V[r * 3 + q] = s * vp + c * vq;
      }
    }
  }
  double d[3] = {a00, a11, a22};
  int idx[3] = {0, 1, 2};
  // stable insertion sort ascending
  for (int i = 1; i < 3; ++i)
    for (int j = i; j > 0 && d[idx[j]] < d[idx[j - 1]]; --j) std::swap(idx[j], idx[j - 1]);
  for (int k = 0; k < 3; ++k) {
    evals[k] = d[idx[k]];
    for (int r = 0; r < 3; ++r) evecs[r * 3 + k] = V[r * 3 + idx[k]];
  }
}

bool inv3(const double M[9], double out[9]) {  // cofactor inverse (Eigen's fixed-size 3x3 path)
  const double c00 = M[4] * M[8] - M[5] * M[7];
  const double c01 = M[5] * M[6] - M[3] * M[8];
  const double c02 = M[3] * M[7] - M[4] * M[6];
  const double det = M[0] * c00 + M[1] * c01 + M[2] * c02;
  const double id = 1.0 / det;
  out[0] = c00 * id;
  out[1] = (M[2] * M[7] - M[1] * M[8]) * id;
  out[2] = (M[1] * M[5] - M[2] * M[4]) * id;
  out[3] = c01 * id;
  out[4] = (M[0] * M[8] - M[2] * M[6]) * id;
  out[5] = (M[2] * M[3] - M[0] * M[5]) * id;
  out[6] = c02 * id;
  out[7] = (M[1] * M[6] - M[0] * M[7]) * id;
  out[8] = (M[0] * M[4] - M[1] * M[3]) * id;
  return det != 0.0;
}

// ------------------------------------------------------------------------------------------------------------
// a-1  applyFilter (no distance-field filtering: filter_field_name_ is empty on this path)
VoxelMap* voxel_build(const float* pts, size_t stride_floats, int64_t n, float leaf_size, int min_points,
                      double eig_mult) {
  VoxelMap* m = new VoxelMap();
  m->pts = pts; m->stride = stride_floats; m->n = n;
  const float inv = 1.0f / leaf_size;  // VoxelGrid::setLeafSize: inverse_leaf_size_ = 1/leaf (Array4f)
  m->inv_leaf = inv;
  // pcl::getMinMax3D on a non-dense cloud: skip non-finite points  (impl.hpp:72)
  float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
  float mx[3] = {-mn[0], -mn[1], -mn[2]};
  for (int64_t i = 0; i < n; ++i) {
    const float* p = P(*m, i);
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    for (int k = 0; k < 3; ++k) { mn[k] = std::min(mn[k], p[k]); mx[k] = std::max(mx[k], p[k]); }
  }
  if (mn[0] > mx[0]) { m->status = 1; return m; }  // no finite point
  // overflow guard (impl.hpp:75-84)
  const int64_t dx = static_cast<int64_t>((mx[0] - mn[0]) * inv) + 1;
  const int64_t dy = static_cast<int64_t>((mx[1] - mn[1]) * inv) + 1;
  const int64_t dz = static_cast<int64_t>((mx[2] - mn[2]) * inv) + 1;
  if (static_cast<double>(dx) * static_cast<double>(dy) * static_cast<double>(dz) > 2147483647.0 ||
      dx * dy * dz > static_cast<int64_t>(std::numeric_limits<int32_t>::max())) { m->status = 2; return m; }
  for (int k = 0; k < 3; ++k) {  // impl.hpp:87-92
    m->min_b[k] = static_cast<int>(std::floor(mn[k] * inv));
    m->max_b[k] = static_cast<int>(std::floor(mx[k] * inv));
    m->div_b[k] = m->max_b[k] - m->min_b[k] + 1;
  }
  m->divb_mul[0] = 1; m->divb_mul[1] = m->div_b[0]; m->divb_mul[2] = m->div_b[0] * m->div_b[1];  // :103
  // first pass (impl.hpp:211-269)
  for (int64_t cp = 0; cp < n; ++cp) {
    const float* p = P(*m, cp);
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    const int ijk0 = static_cast<int>(std::floor(p[0] * inv) - static_cast<float>(m->min_b[0]));
    const int ijk1 = static_cast<int>(std::floor(p[1] * inv) - static_cast<float>(m->min_b[1]));
    const int ijk2 = static_cast<int>(std::floor(p[2] * inv) - static_cast<float>(m->min_b[2]));
    const int idx = ijk0 * m->divb_mul[0] + ijk1 * m->divb_mul[1] + ijk2 * m->divb_mul[2];
    Leaf& leaf = m->leaves[static_cast<size_t>(idx)];
    const double x = p[0], y = p[1], z = p[2];
    leaf.mean[0] += x; leaf.mean[1] += y; leaf.mean[2] += z;
    const double v[3] = {x, y, z};
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) leaf.cov[r * 3 + c] += v[r] * v[c];
    ++leaf.nr_points;
    leaf.points.push_back(static_cast<int32_t>(cp));
    ++m->n_binned;
  }
  // second pass (impl.hpp:285-371)
  for (auto& kv : m->leaves) {
    Leaf& leaf = kv.second;
    const double pt_sum[3] = {leaf.mean[0], leaf.mean[1], leaf.mean[2]};
    for (int k = 0; k < 3; ++k) leaf.mean[k] /= leaf.nr_points;
    if (leaf.nr_points < min_points) continue;
    const double nn = leaf.nr_points;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c)  // :333  (cov - 2*(pt_sum*mean^T))/n + mean*mean^T
        leaf.cov[r * 3 + c] = (leaf.cov[r * 3 + c] - 2 * (pt_sum[r] * leaf.mean[c])) / nn + leaf.mean[r] * leaf.mean[c];
    for (int k = 0; k < 9; ++k) leaf.cov[k] *= (nn - 1.0) / nn;  // :334
    double ev[3];
    jacobi3(leaf.cov, ev, leaf.evecs);  // :337-339
    if (ev[0] < 0 || ev[1] < 0 || ev[2] <= 0) { leaf.nr_points = -1; continue; }  // :341-345
    const double min_ev = eig_mult * ev[2];  // :349
    if (ev[0] < min_ev) {
      ev[0] = min_ev;
      if (ev[1] < min_ev) ev[1] = min_ev;
      double Vi[9], VD[9];
      inv3(leaf.evecs, Vi);
      for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) VD[r * 3 + c] = leaf.evecs[r * 3 + c] * ev[c];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)  // :359  evecs * eigen_val * evecs.inverse()
          leaf.cov[r * 3 + c] = VD[r * 3 + 0] * Vi[0 * 3 + c] + VD[r * 3 + 1] * Vi[1 * 3 + c] + VD[r * 3 + 2] * Vi[2 * 3 + c];
    }
    for (int k = 0; k < 3; ++k) leaf.evals[k] = ev[k];
    inv3(leaf.cov, leaf.icov);  // :363
    double mxc = leaf.icov[0], mnc = leaf.icov[0];
    for (int k = 1; k < 9; ++k) { mxc = std::max(mxc, leaf.icov[k]); mnc = std::min(mnc, leaf.icov[k]); }
    if (mxc == std::numeric_limits<float>::infinity() || mnc == -std::numeric_limits<float>::infinity())
      leaf.nr_points = -1;  // :364-368
  }
  return m;
}

// ------------------------------------------------------------------------------------------------------------
// a-2  checkPlaneType (L/src/core/surfel_association.cpp:246-266) with sort_vec (L/include/utils/eigen_utils.hpp:72-87)
void sort_desc3(const double v[3], double sorted[3], int ind[3]) {
  ind[0] = 0; ind[1] = 1; ind[2] = 2;
  for (int i = 1; i < 3; ++i)
    for (int j = i; j > 0 && v[ind[j]] > v[ind[j - 1]]; --j) std::swap(ind[j], ind[j - 1]);
  for (int i = 0; i < 3; ++i) sorted[i] = v[ind[i]];
}
int check_plane_type(const double evals[3], const double evecs[9], double lambda) {
  double s[3]; int ind[3];
  sort_desc3(evals, s, ind);
  const double p = 2 * (s[1] - s[2]) / (s[2] + s[1] + s[0]);
  if (p < lambda) return -1;
  const int mi = ind[2];
  double nrm[3] = {std::fabs(evecs[0 * 3 + mi]), std::fabs(evecs[1 * 3 + mi]), std::fabs(evecs[2 * 3 + mi])};
  sort_desc3(nrm, s, ind);
  return ind[2];
}

// counter-based RNG shared with the CUDA kernel (DESIGN.md "deterministic RANSAC")
inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
inline uint32_t draw(uint64_t seed, uint32_t attempt, uint32_t j, uint32_t n) {
  return static_cast<uint32_t>(mix64(seed ^ (static_cast<uint64_t>(attempt) * 4u + j)) % n);
}
inline float plane_dist(const float c[4], const float* p) {  // fabsf(dot4(model, pt)), fixed order
  float t = c[0] * p[0];
  t = t + c[1] * p[1];
  t = t + c[2] * p[2];
  t = t + c[3];
  return std::fabs(t);
}
bool model_from3(const float* p0, const float* p1, const float* p2, float c[4]) {
  const float e1[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
  const float e2[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
  // SampleConsensusModelPlane::computeModelCoefficients collinearity test
  const float r0 = e1[0] / e2[0], r1 = e1[1] / e2[1], r2 = e1[2] / e2[2];
  if (r0 == r1 && r2 == r1) return false;
  float nx = e1[1] * e2[2] - e1[2] * e2[1];
  float ny = e1[2] * e2[0] - e1[0] * e2[2];
  float nz = e1[0] * e2[1] - e1[1] * e2[0];
  float l2 = nx * nx;
  l2 = l2 + ny * ny;
  l2 = l2 + nz * nz;
  const float len = std::sqrt(l2);
  if (!(len > 0.0f) || !std::isfinite(len)) return false;
  c[0] = nx / len; c[1] = ny / len; c[2] = nz / len;
  float d = c[0] * p0[0];
  d = d + c[1] * p0[1];
  d = d + c[2] * p0[2];
  c[3] = -d;
  return true;
}

// fitPlane: pcl::SACSegmentation(PLANE, RANSAC, thr, optimize=true)  (L/src/core/surfel_association.cpp:268-294)
bool fit_plane(const VoxelMap& m, const std::vector<int32_t>& idx, uint64_t key, float thr, int min_inliers,
               double p4[4], int32_t* n_inl) {
  const uint32_t n = static_cast<uint32_t>(idx.size());
  if (n < 3) return false;
  const int max_iterations = 50;                 // SACSegmentation default
  const double log_probability = std::log(1.0 - 0.99);
  const double one_over_n = 1.0 / static_cast<double>(n);
  const uint64_t seed = mix64(key * 0x2545F4914F6CDD1Dull + 12345ull);
  float best[4] = {0, 0, 0, 0};
  int best_count = 0;
  double k = 1.0;
  int iterations = 0, skipped = 0;
  uint32_t attempt = 0;
  const int max_skip = max_iterations * 10;
  while (iterations < k && skipped < max_skip) {
    uint32_t a = draw(seed, attempt, 0, n);
    uint32_t b = draw(seed, attempt, 1, n - 1);
    uint32_t c = draw(seed, attempt, 2, n - 2);
    ++attempt;
    if (b >= a) ++b;
    const uint32_t lo = std::min(a, b), hi = std::max(a, b);
    if (c >= lo) ++c;
    if (c >= hi) ++c;
    float coef[4];
    if (!model_from3(P(m, idx[a]), P(m, idx[b]), P(m, idx[c]), coef)) { ++skipped; continue; }
    int count = 0;
    for (uint32_t i = 0; i < n; ++i) count += plane_dist(coef, P(m, idx[i])) < thr ? 1 : 0;
    if (count > best_count) {
      best_count = count;
      std::memcpy(best, coef, sizeof(best));
      const double w = static_cast<double>(count) * one_over_n;
      double p_no = 1.0 - w * w * w;
      p_no = std::max(std::numeric_limits<double>::epsilon(), p_no);
      p_no = std::min(1.0 - std::numeric_limits<double>::epsilon(), p_no);
      k = log_probability / std::log(p_no);
    }
    ++iterations;
    if (iterations > max_iterations) break;
  }
  if (best_count == 0) return false;
  // optimizeModelCoefficients: PCA over the inliers (needs more than the 3-point sample)
  float fin[4] = {best[0], best[1], best[2], best[3]};
  if (best_count > 3) {
    int64_t first = -1;
    double s[3] = {0, 0, 0}, ss[6] = {0, 0, 0, 0, 0, 0};
    double x0[3] = {0, 0, 0};
    int cnt = 0;
    for (uint32_t i = 0; i < n; ++i) {
      const float* p = P(m, idx[i]);
      if (!(plane_dist(best, p) < thr)) continue;
      if (first < 0) { first = i; x0[0] = p[0]; x0[1] = p[1]; x0[2] = p[2]; }
      const double dx = static_cast<double>(p[0]) - x0[0], dy = static_cast<double>(p[1]) - x0[1], dz = static_cast<double>(p[2]) - x0[2];
      s[0] += dx; s[1] += dy; s[2] += dz;
      ss[0] += dx * dx; ss[1] += dx * dy; ss[2] += dx * dz; ss[3] += dy * dy; ss[4] += dy * dz; ss[5] += dz * dz;
      ++cnt;
    }
    const double inv_n = 1.0 / cnt;
    const double mr[3] = {s[0] * inv_n, s[1] * inv_n, s[2] * inv_n};
    // float-precision centroid / covariance (pcl::computeMeanAndCovarianceMatrix works in float)
    const float cf[3] = {static_cast<float>(x0[0] + mr[0]), static_cast<float>(x0[1] + mr[1]), static_cast<float>(x0[2] + mr[2])};
    const float cv[6] = {static_cast<float>(ss[0] * inv_n - mr[0] * mr[0]), static_cast<float>(ss[1] * inv_n - mr[0] * mr[1]),
                         static_cast<float>(ss[2] * inv_n - mr[0] * mr[2]), static_cast<float>(ss[3] * inv_n - mr[1] * mr[1]),
                         static_cast<float>(ss[4] * inv_n - mr[1] * mr[2]), static_cast<float>(ss[5] * inv_n - mr[2] * mr[2])};
    const double A[9] = {cv[0], cv[1], cv[2], cv[1], cv[3], cv[4], cv[2], cv[4], cv[5]};
    double ev[3], V[9];
    jacobi3(A, ev, V);
    const double nx = V[0], ny = V[3], nz = V[6];  // eigenvector of the smallest eigenvalue (pcl::eigen33)
    double d = nx * static_cast<double>(cf[0]);
    d = d + ny * static_cast<double>(cf[1]);
    d = d + nz * static_cast<double>(cf[2]);
    fin[0] = static_cast<float>(nx); fin[1] = static_cast<float>(ny); fin[2] = static_cast<float>(nz);
    fin[3] = static_cast<float>(-d);
  }
  // SACSegmentation::segment re-selects the inliers of the refined model
  int cnt2 = 0;
  for (uint32_t i = 0; i < n; ++i) cnt2 += plane_dist(fin, P(m, idx[i])) < thr ? 1 : 0;
  *n_inl = cnt2;
  if (cnt2 < min_inliers) return false;  // surfel_association.cpp:284
  for (int k2 = 0; k2 < 4; ++k2) p4[k2] = fin[k2];
  return true;
}

SurfelSet* surfel_extract(const VoxelMap& m, double lambda, int min_leaf_points, float thr, int min_inliers) {
  SurfelSet* s = new SurfelSet();
  for (const auto& kv : m.leaves) {  // ascending key order = std::map iteration (surfel_association.cpp:60)
    const Leaf& leaf = kv.second;
    if (leaf.nr_points < min_leaf_points) continue;
    if (check_plane_type(leaf.evals, leaf.evecs, lambda) < 0) continue;
    Plane pl;
    if (!fit_plane(m, leaf.points, kv.first, thr, min_inliers, pl.p4, &pl.n_inliers)) continue;
    for (int k = 0; k < 3; ++k) pl.Pi[k] = -pl.p4[3] * pl.p4[k];  // :79
    float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
    float mx[3] = {-mn[0], -mn[1], -mn[2]};
    for (int32_t i : leaf.points) {  // pcl::getMinMax3D(surfplane.cloud, min, max)  :82
      const float* p = P(m, i);
      for (int k = 0; k < 3; ++k) { mn[k] = std::min(mn[k], p[k]); mx[k] = std::max(mx[k], p[k]); }
    }
    for (int k = 0; k < 3; ++k) { pl.bmin[k] = mn[k]; pl.bmax[k] = mx[k]; }
    pl.key = static_cast<int64_t>(kv.first);
    s->planes.push_back(pl);
  }
  return s;
}

// ------------------------------------------------------------------------------------------------------------
// a-3  association
struct RawPoint { float x, y, z, pad; float intensity; float pad2; double timestamp; };  // licalib::PointXYZIT, 32 B
struct SurfelPoint { double timestamp; double point[3]; double point_in_map[3]; int64_t plane_id; };

inline double p2plane(const double pt[3], const double p4[4]) {  // point2PlaneDistance :296-303
  double d = pt[0] * p4[0];
  d = d + pt[1] * p4[1];
  d = d + pt[2] * p4[2];
  d = d + p4[3];
  return d > 0 ? d : -d;
}
inline bool hit(const Plane& pl, const float* q, double radius) {  // associateScanToSurfel inner test :317-326
  if (std::isnan(q[0])) return false;
  if (!(q[0] > pl.bmin[0] && q[0] < pl.bmax[0] && q[1] > pl.bmin[1] && q[1] < pl.bmax[1] && q[2] > pl.bmin[2] && q[2] < pl.bmax[2]))
    return false;
  const double pt[3] = {q[0], q[1], q[2]};
  return p2plane(pt, pl.p4) <= radius;
}

// mode 0: reference-faithful brute force over planes (omp over planes, :122); mode 1: O(N) voxel lookup
void associate_scan(const VoxelMap& m, const SurfelSet& s, const std::map<int64_t, int>& key2plane, const float* scan_map,
                    size_t map_stride, const RawPoint* raw, int W, int H, double radius, int k_per_ring, int mode,
                    std::vector<SurfelPoint>& all) {
  std::vector<int> flag(static_cast<size_t>(W) * H, -1);
  const int np = static_cast<int>(s.planes.size());
  if (mode == 0) {
#pragma omp parallel for schedule(dynamic, 8)
    for (int pid = 0; pid < np; ++pid) {
      const Plane& pl = s.planes[pid];
      std::vector<int> ring;
      for (int h = 0; h < H; ++h) {
        ring.clear();
        for (int w = 0; w < W; ++w)
          if (hit(pl, scan_map + (static_cast<size_t>(h) * W + w) * map_stride, radius)) ring.push_back(w);
        if (static_cast<int>(ring.size()) < k_per_ring * 2) continue;  // :128
        int step = static_cast<int>(ring.size()) / (k_per_ring + 1);
        step = std::max(step, 1);
        for (int sel = 0; sel < k_per_ring; ++sel) flag[static_cast<size_t>(h) * W + ring[step * (sel + 1) - 1]] = pid;
      }
    }
  } else {
    std::vector<int> cand(static_cast<size_t>(W) * H, -1);
    for (int h = 0; h < H; ++h)
      for (int w = 0; w < W; ++w) {
        const float* q = scan_map + (static_cast<size_t>(h) * W + w) * map_stride;
        if (!std::isfinite(q[0]) || !std::isfinite(q[1]) || !std::isfinite(q[2])) continue;
        const int i0 = static_cast<int>(std::floor(q[0] * m.inv_leaf) - static_cast<float>(m.min_b[0]));
        const int i1 = static_cast<int>(std::floor(q[1] * m.inv_leaf) - static_cast<float>(m.min_b[1]));
        const int i2 = static_cast<int>(std::floor(q[2] * m.inv_leaf) - static_cast<float>(m.min_b[2]));
        if (i0 < 0 || i1 < 0 || i2 < 0 || i0 >= m.div_b[0] || i1 >= m.div_b[1] || i2 >= m.div_b[2]) continue;
        const int64_t key = static_cast<int64_t>(i0) * m.divb_mul[0] + static_cast<int64_t>(i1) * m.divb_mul[1] + static_cast<int64_t>(i2) * m.divb_mul[2];
        auto it = key2plane.find(key);
        if (it == key2plane.end()) continue;
        if (hit(s.planes[it->second], q, radius)) cand[static_cast<size_t>(h) * W + w] = it->second;
      }
    std::map<int, std::vector<int>> per;
    for (int h = 0; h < H; ++h) {
      per.clear();
      for (int w = 0; w < W; ++w) { const int c = cand[static_cast<size_t>(h) * W + w]; if (c >= 0) per[c].push_back(w); }
      for (auto& kv : per) {
        const std::vector<int>& ring = kv.second;
        if (static_cast<int>(ring.size()) < k_per_ring * 2) continue;
        int step = static_cast<int>(ring.size()) / (k_per_ring + 1);
        step = std::max(step, 1);
        for (int sel = 0; sel < k_per_ring; ++sel) flag[static_cast<size_t>(h) * W + ring[step * (sel + 1) - 1]] = kv.first;
      }
    }
  }
  // chronological emission (:139-158): w outer, h inner, skip timestamp == 0
  for (int w = 0; w < W; ++w)
    for (int h = 0; h < H; ++h) {
      const size_t i = static_cast<size_t>(h) * W + w;
      if (flag[i] == -1 || 0 == raw[i].timestamp) continue;
      SurfelPoint sp;
      sp.timestamp = raw[i].timestamp;
      sp.point[0] = raw[i].x; sp.point[1] = raw[i].y; sp.point[2] = raw[i].z;
      const float* q = scan_map + i * map_stride;
      sp.point_in_map[0] = q[0]; sp.point_in_map[1] = q[1]; sp.point_in_map[2] = q[2];
      sp.plane_id = flag[i];
      all.push_back(sp);
    }
}

}  // namespace

extern "C" {

void* orc_voxel_build(const float* pts, int64_t stride_floats, int64_t n, float leaf, int min_points, double eig_mult) {
  return voxel_build(pts, static_cast<size_t>(stride_floats), n, leaf, min_points, eig_mult);
}
int orc_voxel_status(void* h) { return static_cast<VoxelMap*>(h)->status; }
void orc_voxel_free(void* h) { delete static_cast<VoxelMap*>(h); }
int64_t orc_voxel_num_leaves(void* h) { return static_cast<int64_t>(static_cast<VoxelMap*>(h)->leaves.size()); }
int64_t orc_voxel_num_points(void* h) { return static_cast<VoxelMap*>(h)->n_binned; }
void orc_voxel_grid(void* h, int32_t* min_b, int32_t* div_b) {
  VoxelMap* m = static_cast<VoxelMap*>(h);
  for (int k = 0; k < 3; ++k) { min_b[k] = m->min_b[k]; div_b[k] = m->div_b[k]; }
}
void orc_voxel_export(void* h, int64_t* keys, int32_t* nr, double* mean, double* cov, double* evals, double* evecs,
                      double* icov, int64_t* leaf_start, int32_t* point_index) {
  VoxelMap* m = static_cast<VoxelMap*>(h);
  int64_t l = 0, off = 0;
  for (const auto& kv : m->leaves) {
    const Leaf& f = kv.second;
    if (keys) keys[l] = static_cast<int64_t>(kv.first);
    if (nr) nr[l] = f.nr_points;
    if (mean) std::memcpy(mean + l * 3, f.mean, 24);
    if (cov) std::memcpy(cov + l * 9, f.cov, 72);
    if (evals) std::memcpy(evals + l * 3, f.evals, 24);
    if (evecs) std::memcpy(evecs + l * 9, f.evecs, 72);
    if (icov) std::memcpy(icov + l * 9, f.icov, 72);
    if (leaf_start) leaf_start[l] = off;
    if (point_index) std::memcpy(point_index + off, f.points.data(), f.points.size() * 4);
    off += static_cast<int64_t>(f.points.size());
    ++l;
  }
  if (leaf_start) leaf_start[l] = off;
}

void* orc_surfel_extract(void* h, double lambda, int min_leaf_points, float thr, int min_inliers) {
  return surfel_extract(*static_cast<VoxelMap*>(h), lambda, min_leaf_points, thr, min_inliers);
}
void orc_surfel_free(void* s) { delete static_cast<SurfelSet*>(s); }
int64_t orc_surfel_count(void* s) { return static_cast<int64_t>(static_cast<SurfelSet*>(s)->planes.size()); }
void orc_surfel_export(void* sp, double* p4, double* Pi, double* bmin, double* bmax, int64_t* key, int32_t* ninl) {
  SurfelSet* s = static_cast<SurfelSet*>(sp);
  for (size_t i = 0; i < s->planes.size(); ++i) {
    const Plane& p = s->planes[i];
    if (p4) std::memcpy(p4 + i * 4, p.p4, 32);
    if (Pi) std::memcpy(Pi + i * 3, p.Pi, 24);
    if (bmin) std::memcpy(bmin + i * 3, p.bmin, 24);
    if (bmax) std::memcpy(bmax + i * 3, p.bmax, 24);
    if (key) key[i] = p.key;
    if (ninl) ninl[i] = p.n_inliers;
  }
}

// getAssociation over n_scans + averageTimeDownSmaple(time_step).  Returns number of downsampled points;
// *n_all = spoints_all_.size().  out may be NULL (sizing call).
int64_t orc_associate(void* h, void* sp, const float* scans_map, int64_t map_stride_floats, const void* scans_raw,
                      int32_t n_scans, int32_t W, int32_t H, double radius, int32_t k_per_ring, int32_t time_step,
                      int32_t mode, void* out, int64_t cap, int64_t* n_all) {
  VoxelMap* m = static_cast<VoxelMap*>(h);
  SurfelSet* s = static_cast<SurfelSet*>(sp);
  std::map<int64_t, int> key2plane;
  for (size_t i = 0; i < s->planes.size(); ++i) key2plane[s->planes[i].key] = static_cast<int>(i);
  std::vector<SurfelPoint> all;
  const RawPoint* raw = static_cast<const RawPoint*>(scans_raw);
  for (int32_t sc = 0; sc < n_scans; ++sc) {
    const size_t off = static_cast<size_t>(sc) * W * H;
    associate_scan(*m, *s, key2plane, scans_map + off * map_stride_floats, static_cast<size_t>(map_stride_floats), raw + off, W, H,
                   radius, k_per_ring, mode, all);
  }
  if (n_all) *n_all = static_cast<int64_t>(all.size());
  int64_t cnt = 0;
  SurfelPoint* o = static_cast<SurfelPoint*>(out);
  for (size_t idx = 0; idx < all.size(); idx += static_cast<size_t>(time_step)) {  // :240-244
    if (o && cnt < cap) o[cnt] = all[idx];
    ++cnt;
  }
  return cnt;
}

// associateVisualPointsWithPlanes inner test (L/src/core/surfel_association.cpp:196-210) for pre-computed L0-frame
// landmark positions: strict bbox + point2PlaneDistance <= 2*radius; the LAST matching plane wins (:206).
void orc_associate_landmarks(void* sp, const double* pts, int64_t n, double radius, int32_t* plane_out) {
  SurfelSet* s = static_cast<SurfelSet*>(sp);
  for (int64_t i = 0; i < n; ++i) {
    const double* q = pts + 3 * i;
    int best = -1;
    for (size_t k = 0; k < s->planes.size(); ++k) {
      const Plane& pl = s->planes[k];
      if (q[0] > pl.bmin[0] && q[0] < pl.bmax[0] && q[1] > pl.bmin[1] && q[1] < pl.bmax[1] && q[2] > pl.bmin[2] && q[2] < pl.bmax[2]) {
        if (p2plane(q, pl.p4) <= radius * 2) best = static_cast<int>(k);
      }
    }
    plane_out[i] = best;
  }
}

}  // extern "C"
