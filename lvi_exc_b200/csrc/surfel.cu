// surfel.cu — planar-surfel extraction from the NDT leaves on sm_100a.
//
// Replaces SurfelAssociation::setSurfelMap / checkPlaneType / fitPlane (L/src/core/surfel_association.cpp:50-108,246-294)
// which loops serially over the std::map, copies every leaf cloud twice and runs pcl::SACSegmentation per leaf.
//   1. surfel_candidate_kernel : thread per leaf — nr_points >= 10 and planarity 2(l1-l2)/(l0+l1+l2) >= lambda (Lv et al. eq. 13)
//   2. surfel_fit_kernel       : one CTA per candidate leaf — RANSAC plane (50 iterations max, p = 0.99 adaptive stop,
//                                |n.x+d| < thr inlier test), PCA refinement over the inliers, inlier re-selection, bbox
//   3. compaction in leaf order -> plane_id = position in ascending voxel-index order (reference: push_back order)
// Leaf points are contiguous float4 in HBM (voxel.cu), so every RANSAC pass is a coalesced stream.
// PCL is not available: the sample sequence comes from a counter-based RNG keyed by the voxel index and the
// PCA sums are accumulated in fp64 and rounded to float, which makes the result independent of summation order; the
// CPU oracle implements the identical specification.  Compiled with -fmad=false.
#include <cub/cub.cuh>

#include "eig3.cuh"
#include <cstdlib>
#include <memory>

#include "map.cuh"

namespace lvi {

__device__ __forceinline__ void sort_desc3(const double v[3], double s[3], int ind[3]) {  // Eigen::sort_vec (L/include/utils/eigen_utils.hpp:72-87)
  ind[0] = 0; ind[1] = 1; ind[2] = 2;
  for (int i = 1; i < 3; ++i)
    for (int j = i; j > 0 && v[ind[j]] > v[ind[j - 1]]; --j) { const int t = ind[j]; ind[j] = ind[j - 1]; ind[j - 1] = t; }
  for (int i = 0; i < 3; ++i) s[i] = v[ind[i]];
}

__global__ void __launch_bounds__(256) surfel_candidate_kernel(const int32_t* __restrict__ npts, const double* __restrict__ evals, int n_leaves,
                                                               int min_leaf_points, double lambda, int32_t* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_leaves) return;
  int ok = 0;
  if (npts[i] >= min_leaf_points) {  // surfel_association.cpp:63
    const double ev[3] = {evals[3 * i], evals[3 * i + 1], evals[3 * i + 2]};
    double s[3]; int ind[3];
    sort_desc3(ev, s, ind);
    const double p = 2 * (s[1] - s[2]) / (s[2] + s[1] + s[0]);  // :252-253
    ok = !(p < lambda);
  }
  flag[i] = ok;
}

__global__ void __launch_bounds__(256) surfel_scatter_ids_kernel(const int32_t* __restrict__ flag, const int32_t* __restrict__ pos, int n,
                                                                 int32_t* __restrict__ ids) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flag[i]) ids[pos[i]] = i;
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ uint32_t draw(uint64_t seed, uint32_t attempt, uint32_t j, uint32_t n) {
  return static_cast<uint32_t>(mix64(seed ^ (static_cast<uint64_t>(attempt) * 4u + j)) % n);
}
__device__ __forceinline__ float plane_dist(const float c[4], const float4 p) {
  float t = c[0] * p.x;
  t = t + c[1] * p.y;
  t = t + c[2] * p.z;
  t = t + c[3];
  return fabsf(t);
}
__device__ __forceinline__ bool model_from3(const float4 p0, const float4 p1, const float4 p2, float c[4]) {
  const float e1x = p1.x - p0.x, e1y = p1.y - p0.y, e1z = p1.z - p0.z;
  const float e2x = p2.x - p0.x, e2y = p2.y - p0.y, e2z = p2.z - p0.z;
  const float r0 = e1x / e2x, r1 = e1y / e2y, r2 = e1z / e2z;  // SampleConsensusModelPlane collinearity test
  if (r0 == r1 && r2 == r1) return false;
  const float nx = e1y * e2z - e1z * e2y;
  const float ny = e1z * e2x - e1x * e2z;
  const float nz = e1x * e2y - e1y * e2x;
  float l2 = nx * nx;
  l2 = l2 + ny * ny;
  l2 = l2 + nz * nz;
  const float len = sqrtf(l2);
  if (!(len > 0.0f) || !isfinite(len)) return false;
  c[0] = nx / len; c[1] = ny / len; c[2] = nz / len;
  float d = c[0] * p0.x;
  d = d + c[1] * p0.y;
  d = d + c[2] * p0.z;
  c[3] = -d;
  return true;
}

struct FitOut { double p4[4]; double bmin[3], bmax[3]; int ninl; int ok; };

// One CTA (128 threads) per candidate leaf: a leaf of the room scene holds ~10^4 points and RANSAC makes up to 50 passes over them,
// so one warp per leaf (the first version) left the kernel latency-bound at 5 % of the warp slots.
constexpr int kFitThreads = 128;
__device__ __forceinline__ int block_sum_int(int v, int* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  return sh[0] + sh[1] + sh[2] + sh[3];
}
__device__ __forceinline__ unsigned block_min_uint(unsigned v, unsigned* sh) {
  v = __reduce_min_sync(0xffffffffu, v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  return min(min(sh[0], sh[1]), min(sh[2], sh[3]));
}
__device__ __forceinline__ double block_sum_double(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  return (sh[0] + sh[1]) + (sh[2] + sh[3]);
}
__device__ __forceinline__ float block_min_float(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  return fminf(fminf(sh[0], sh[1]), fminf(sh[2], sh[3]));
}

__global__ void __launch_bounds__(kFitThreads) surfel_fit_kernel(const float4* __restrict__ pts, const int32_t* __restrict__ leaf_start,
                                                                 const int32_t* __restrict__ leaf_key, const int32_t* __restrict__ cand, int n_cand,
                                                                 float thr, int min_inliers, uint32_t small_max, uint32_t large_min, FitOut* __restrict__ out) {
  __shared__ int shi[4];
  __shared__ unsigned shu[4];
  __shared__ double shd[4];
  __shared__ float shf[4];
  const int tid = threadIdx.x;
  __shared__ int s_leaf[32];
  __shared__ uint32_t s_size[32];
  const int step = static_cast<int>(gridDim.x);
  for (int ci0 = blockIdx.x; ci0 < n_cand; ci0 += 32 * step) {   // candidate sizes 32 at a time (see surfel_fit_cluster_kernel)
  __syncthreads();
  if (tid < 32) {
    const int c = ci0 + tid * step;
    s_size[tid] = 0;
    if (c < n_cand) { const int l = cand[c]; s_leaf[tid] = l; s_size[tid] = static_cast<uint32_t>(leaf_start[l + 1] - leaf_start[l]); }
  }
  __syncthreads();
  for (int q = 0; q < 32; ++q) {
    const int ci = ci0 + q * step;
    if (ci >= n_cand) break;
    const uint32_t n = s_size[q];
    if (n <= small_max || n > large_min) continue;   // surfel_fit_warp_kernel's / surfel_fit_cluster_kernel's
    const int leaf = s_leaf[q];
    const int beg = leaf_start[leaf];
    const float4* P = pts + beg;
    FitOut fo;
    fo.ok = 0; fo.ninl = 0;
    for (int k = 0; k < 4; ++k) fo.p4[k] = 0;
    for (int k = 0; k < 3; ++k) { fo.bmin[k] = 0; fo.bmax[k] = 0; }
    bool fail = n < 3;
    float best[4] = {0, 0, 0, 0};
    int best_count = 0;
    if (!fail) {  // every thread runs the same (uniform) RANSAC control flow; only the inlier counting is split
      const int max_iterations = 50;
      const double log_probability = log(1.0 - 0.99);
      const double one_over_n = 1.0 / static_cast<double>(n);
      const uint64_t seed = mix64(static_cast<uint64_t>(static_cast<uint32_t>(leaf_key[leaf])) * 0x2545F4914F6CDD1Dull + 12345ull);
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
        const uint32_t lo = min(a, b), hi = max(a, b);
        if (c >= lo) ++c;
        if (c >= hi) ++c;
        float coef[4];
        if (!model_from3(__ldg(P + a), __ldg(P + b), __ldg(P + c), coef)) { ++skipped; continue; }
        int count = 0;
        for (uint32_t i = tid; i < n; i += kFitThreads) count += plane_dist(coef, __ldg(P + i)) < thr ? 1 : 0;
        count = block_sum_int(count, shi);
        if (count > best_count) {
          best_count = count;
          best[0] = coef[0]; best[1] = coef[1]; best[2] = coef[2]; best[3] = coef[3];
          const double w = static_cast<double>(count) * one_over_n;
          double p_no = 1.0 - w * w * w;
          p_no = fmax(2.220446049250313e-16, p_no);
          p_no = fmin(1.0 - 2.220446049250313e-16, p_no);
          k = log_probability / log(p_no);
        }
        ++iterations;
        if (iterations > max_iterations) break;
      }
      if (best_count == 0) fail = true;
    }
    float fin[4] = {best[0], best[1], best[2], best[3]};
    if (!fail && best_count > 3) {
      // first inlier in leaf order defines the local origin of the PCA sums
      uint32_t first = 0xffffffffu;
      for (uint32_t i = tid; i < n && first == 0xffffffffu; i += kFitThreads) if (plane_dist(best, __ldg(P + i)) < thr) first = i;
      first = block_min_uint(first, shu);
      const float4 p0 = __ldg(P + first);
      const double x0 = p0.x, y0 = p0.y, z0 = p0.z;
      double sums[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      int cnt = 0;
      for (uint32_t i = tid; i < n; i += kFitThreads) {
        const float4 p = __ldg(P + i);
        if (!(plane_dist(best, p) < thr)) continue;
        const double dx = static_cast<double>(p.x) - x0, dy = static_cast<double>(p.y) - y0, dz = static_cast<double>(p.z) - z0;
        sums[0] += dx; sums[1] += dy; sums[2] += dz;
        sums[3] += dx * dx; sums[4] += dx * dy; sums[5] += dx * dz; sums[6] += dy * dy; sums[7] += dy * dz; sums[8] += dz * dz;
        ++cnt;
      }
#pragma unroll
      for (int q = 0; q < 9; ++q) sums[q] = block_sum_double(sums[q], shd);
      cnt = block_sum_int(cnt, shi);
      const double inv_n = 1.0 / cnt;
      const double m0 = sums[0] * inv_n, m1 = sums[1] * inv_n, m2 = sums[2] * inv_n;
      const float cf0 = static_cast<float>(x0 + m0), cf1 = static_cast<float>(y0 + m1), cf2 = static_cast<float>(z0 + m2);
      const float cv0 = static_cast<float>(sums[3] * inv_n - m0 * m0), cv1 = static_cast<float>(sums[4] * inv_n - m0 * m1),
                  cv2 = static_cast<float>(sums[5] * inv_n - m0 * m2), cv3 = static_cast<float>(sums[6] * inv_n - m1 * m1),
                  cv4 = static_cast<float>(sums[7] * inv_n - m1 * m2), cv5 = static_cast<float>(sums[8] * inv_n - m2 * m2);
      const double A[9] = {cv0, cv1, cv2, cv1, cv3, cv4, cv2, cv4, cv5};
      double ev[3], V[9];
      jacobi3_lower(A, ev, V);
      const double nx = V[0], ny = V[3], nz = V[6];
      double d = nx * static_cast<double>(cf0);
      d = d + ny * static_cast<double>(cf1);
      d = d + nz * static_cast<double>(cf2);
      fin[0] = static_cast<float>(nx); fin[1] = static_cast<float>(ny); fin[2] = static_cast<float>(nz);
      fin[3] = static_cast<float>(-d);
    }
    if (!fail) {
      int cnt2 = 0;
      float mn0 = 3.402823466e38f, mn1 = mn0, mn2 = mn0, mx0 = -mn0, mx1 = -mn0, mx2 = -mn0;
      for (uint32_t i = tid; i < n; i += kFitThreads) {
        const float4 p = __ldg(P + i);
        cnt2 += plane_dist(fin, p) < thr ? 1 : 0;
        mn0 = fminf(mn0, p.x); mn1 = fminf(mn1, p.y); mn2 = fminf(mn2, p.z);
        mx0 = fmaxf(mx0, p.x); mx1 = fmaxf(mx1, p.y); mx2 = fmaxf(mx2, p.z);
      }
      cnt2 = block_sum_int(cnt2, shi);
      mn0 = block_min_float(mn0, shf); mn1 = block_min_float(mn1, shf); mn2 = block_min_float(mn2, shf);
      mx0 = -block_min_float(-mx0, shf); mx1 = -block_min_float(-mx1, shf); mx2 = -block_min_float(-mx2, shf);
      fo.ninl = cnt2;
      fo.ok = cnt2 >= min_inliers;  // surfel_association.cpp:284
      for (int k = 0; k < 4; ++k) fo.p4[k] = fin[k];
      fo.bmin[0] = mn0; fo.bmin[1] = mn1; fo.bmin[2] = mn2; fo.bmax[0] = mx0; fo.bmax[1] = mx1; fo.bmax[2] = mx2;  // getMinMax3D :82
    }
    if (tid == 0) out[ci] = fo;
    __syncthreads();
  }
  }
}

// Large leaves (> kClusterFitMin points; the room scenes of C2 / C4 put 10^4 .. 10^5 points into a 0.5 m voxel): one THREAD-BLOCK CLUSTER of
// kFitCluster CTAs per leaf.  Every CTA streams its 1/8 of the leaf's points per pass; counts, the first-inlier index, the nine fp64 PCA sums and
// the box are reduced inside the CTA and then across the cluster through distributed shared memory (each CTA publishes its partials in its own
// shared memory, one hardware cluster barrier, every CTA reads the eight sets in rank order -- identical results in all CTAs, so the RANSAC
// control flow stays uniform without a broadcast).  Two alternating exchange slots make one cluster barrier per reduction enough.  With one CTA
// per leaf, 252 leaves per rank (C4 at 8 GPUs) left 85 % of the warp slots empty: 3.5 ms per pass.
constexpr int kFitCluster = 8;
constexpr uint32_t kClusterFitMin = 4096;
__device__ __forceinline__ unsigned cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_id_x() { unsigned r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned n_clusters_x() { unsigned r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double ld_dsmem_f64(const double* local, unsigned rank) {   // the same shared-memory variable in CTA `rank` of the cluster
  unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(local)), r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(r) : "memory");
  return v;
}
// Exchange of up to 16 doubles per CTA.  op: 0 = sum, 1 = min.  vals[] holds this thread's values on entry (all threads of the CTA call it) and the
// cluster-wide result in every thread of every CTA on return.
struct ClusterExchange {
  double part[2][16];        // [slot][value]: this CTA's block-level partials
  double warp[kFitThreads / 32][16];
  int use;
};
template <int NV>
__device__ __forceinline__ void cluster_reduce(ClusterExchange& ex, double (&vals)[NV], int op) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double t = __shfl_xor_sync(0xffffffffu, vals[k], o);
      vals[k] = op == 0 ? vals[k] + t : fmin(vals[k], t);
    }
  }
  __syncthreads();   // the previous reduction's reads of ex.warp are done
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) ex.warp[warp][k] = vals[k];
  }
  __syncthreads();
  const int slot = ex.use & 1;
  if (threadIdx.x < NV) {
    double r = ex.warp[0][threadIdx.x];
    for (int w = 1; w < kFitThreads / 32; ++w) r = op == 0 ? r + ex.warp[w][threadIdx.x] : fmin(r, ex.warp[w][threadIdx.x]);
    ex.part[slot][threadIdx.x] = r;
  }
  cluster_barrier();   // every CTA's partials are published (release / acquire at cluster scope)
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double r = ld_dsmem_f64(&ex.part[slot][k], 0);
    for (unsigned c = 1; c < static_cast<unsigned>(kFitCluster); ++c) { const double t = ld_dsmem_f64(&ex.part[slot][k], c); r = op == 0 ? r + t : fmin(r, t); }
    vals[k] = r;
  }
  __syncthreads();
  if (threadIdx.x == 0) ex.use = ex.use + 1;   // the other slot next time: a CTA that reaches the NEXT cluster barrier has finished these reads
}

__global__ void __cluster_dims__(kFitCluster, 1, 1) __launch_bounds__(kFitThreads)
surfel_fit_cluster_kernel(const float4* __restrict__ pts, const int32_t* __restrict__ leaf_start, const int32_t* __restrict__ leaf_key,
                          const int32_t* __restrict__ cand, int n_cand, float thr, int min_inliers, FitOut* __restrict__ out) {
  __shared__ ClusterExchange ex;
  const int tid = threadIdx.x;
  const unsigned rank = cluster_rank();
  if (tid == 0) ex.use = 0;
  __syncthreads();
  const uint32_t stride = kFitCluster * kFitThreads, first_i = rank * kFitThreads + tid;
  __shared__ int s_leaf[kFitThreads];
  __shared__ uint32_t s_size[kFitThreads];
  const int step = static_cast<int>(n_clusters_x());
  // the sizes of this cluster's next 128 candidates are looked up together (two dependent loads per candidate: one at a time, walking past the
  // 150 k small leaves of a C3 map cost 0.5 ms)
  for (int ci0 = cluster_id_x(); ci0 < n_cand; ci0 += kFitThreads * step) {
  __syncthreads();
  {
    const int c = ci0 + tid * step;
    s_size[tid] = 0;
    if (c < n_cand) { const int l = cand[c]; s_leaf[tid] = l; s_size[tid] = static_cast<uint32_t>(leaf_start[l + 1] - leaf_start[l]); }
  }
  __syncthreads();
  for (int q = 0; q < kFitThreads; ++q) {
    const int ci = ci0 + q * step;
    if (ci >= n_cand) break;
    const uint32_t n = s_size[q];
    if (n <= kClusterFitMin) continue;   // the one-CTA / one-warp kernels' (cluster-uniform)
    const int leaf = s_leaf[q];
    const int beg = leaf_start[leaf];
    const float4* P = pts + beg;
    FitOut fo;
    fo.ok = 0; fo.ninl = 0;
    for (int k = 0; k < 4; ++k) fo.p4[k] = 0;
    for (int k = 0; k < 3; ++k) { fo.bmin[k] = 0; fo.bmax[k] = 0; }
    bool fail = false;   // n > kClusterFitMin >= 3
    float best[4] = {0, 0, 0, 0};
    int best_count = 0;
    {
      const int max_iterations = 50;
      const double log_probability = log(1.0 - 0.99);
      const double one_over_n = 1.0 / static_cast<double>(n);
      const uint64_t seed = mix64(static_cast<uint64_t>(static_cast<uint32_t>(leaf_key[leaf])) * 0x2545F4914F6CDD1Dull + 12345ull);
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
        const uint32_t lo = min(a, b), hi = max(a, b);
        if (c >= lo) ++c;
        if (c >= hi) ++c;
        float coef[4];
        if (!model_from3(__ldg(P + a), __ldg(P + b), __ldg(P + c), coef)) { ++skipped; continue; }
        double cv[1] = {0.0};
        for (uint32_t i = first_i; i < n; i += stride) cv[0] += plane_dist(coef, __ldg(P + i)) < thr ? 1.0 : 0.0;   // counts < 2^31: exact in fp64
        cluster_reduce(ex, cv, 0);
        const int count = static_cast<int>(cv[0]);
        if (count > best_count) {
          best_count = count;
          best[0] = coef[0]; best[1] = coef[1]; best[2] = coef[2]; best[3] = coef[3];
          const double w = static_cast<double>(count) * one_over_n;
          double p_no = 1.0 - w * w * w;
          p_no = fmax(2.220446049250313e-16, p_no);
          p_no = fmin(1.0 - 2.220446049250313e-16, p_no);
          k = log_probability / log(p_no);
        }
        ++iterations;
        if (iterations > max_iterations) break;
      }
      if (best_count == 0) fail = true;
    }
    float fin[4] = {best[0], best[1], best[2], best[3]};
    if (!fail && best_count > 3) {
      double fv[1] = {4294967295.0};
      for (uint32_t i = first_i; i < n && fv[0] == 4294967295.0; i += stride) if (plane_dist(best, __ldg(P + i)) < thr) fv[0] = static_cast<double>(i);
      cluster_reduce(ex, fv, 1);
      const uint32_t first = static_cast<uint32_t>(fv[0]);
      const float4 p0 = __ldg(P + first);
      const double x0 = p0.x, y0 = p0.y, z0 = p0.z;
      double sums[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // [9] = inlier count
      for (uint32_t i = first_i; i < n; i += stride) {
        const float4 p = __ldg(P + i);
        if (!(plane_dist(best, p) < thr)) continue;
        const double dx = static_cast<double>(p.x) - x0, dy = static_cast<double>(p.y) - y0, dz = static_cast<double>(p.z) - z0;
        sums[0] += dx; sums[1] += dy; sums[2] += dz;
        sums[3] += dx * dx; sums[4] += dx * dy; sums[5] += dx * dz; sums[6] += dy * dy; sums[7] += dy * dz; sums[8] += dz * dz;
        sums[9] += 1.0;
      }
      cluster_reduce(ex, sums, 0);
      const double inv_n = 1.0 / sums[9];
      const double m0 = sums[0] * inv_n, m1 = sums[1] * inv_n, m2 = sums[2] * inv_n;
      const float cf0 = static_cast<float>(x0 + m0), cf1 = static_cast<float>(y0 + m1), cf2 = static_cast<float>(z0 + m2);
      const float cv0 = static_cast<float>(sums[3] * inv_n - m0 * m0), cv1 = static_cast<float>(sums[4] * inv_n - m0 * m1),
                  cv2 = static_cast<float>(sums[5] * inv_n - m0 * m2), cv3 = static_cast<float>(sums[6] * inv_n - m1 * m1),
                  cv4 = static_cast<float>(sums[7] * inv_n - m1 * m2), cv5 = static_cast<float>(sums[8] * inv_n - m2 * m2);
      const double A[9] = {cv0, cv1, cv2, cv1, cv3, cv4, cv2, cv4, cv5};
      double ev[3], V[9];
      jacobi3_lower(A, ev, V);
      const double nx = V[0], ny = V[3], nz = V[6];
      double d = nx * static_cast<double>(cf0);
      d = d + ny * static_cast<double>(cf1);
      d = d + nz * static_cast<double>(cf2);
      fin[0] = static_cast<float>(nx); fin[1] = static_cast<float>(ny); fin[2] = static_cast<float>(nz);
      fin[3] = static_cast<float>(-d);
    }
    if (!fail) {
      double mm[7] = {3.402823466e38, 3.402823466e38, 3.402823466e38, 3.402823466e38, 3.402823466e38, 3.402823466e38, 0.0};   // min xyz, min of -xyz
      double cnt2[1] = {0.0};
      for (uint32_t i = first_i; i < n; i += stride) {
        const float4 p = __ldg(P + i);
        cnt2[0] += plane_dist(fin, p) < thr ? 1.0 : 0.0;
        mm[0] = fmin(mm[0], static_cast<double>(p.x)); mm[1] = fmin(mm[1], static_cast<double>(p.y)); mm[2] = fmin(mm[2], static_cast<double>(p.z));
        mm[3] = fmin(mm[3], -static_cast<double>(p.x)); mm[4] = fmin(mm[4], -static_cast<double>(p.y)); mm[5] = fmin(mm[5], -static_cast<double>(p.z));
      }
      cluster_reduce(ex, cnt2, 0);
      cluster_reduce(ex, mm, 1);
      fo.ninl = static_cast<int>(cnt2[0]);
      fo.ok = fo.ninl >= min_inliers;
      for (int k = 0; k < 4; ++k) fo.p4[k] = fin[k];
      fo.bmin[0] = mm[0]; fo.bmin[1] = mm[1]; fo.bmin[2] = mm[2]; fo.bmax[0] = -mm[3]; fo.bmax[1] = -mm[4]; fo.bmax[2] = -mm[5];
    }
    if (rank == 0 && tid == 0) out[ci] = fo;
  }
  }
  cluster_barrier();   // no CTA leaves while a neighbour may still read its shared memory
}

// Small leaves (<= kWarpFitMax points: every leaf of a map with 10^5 .. 10^6 surfels): ONE WARP per leaf with the leaf's points staged once
// in the warp's slice of shared memory (coalesced, all loads in flight together), so a RANSAC pass is a short rolled loop over conflict-free
// shared-memory reads and one warp reduction -- no block barrier, no second trip to L1/L2, and a code footprint that stays in the
// instruction cache (the first version kept the points in registers with every pass unrolled 24x: 14 of 35 k stall samples were
// instruction fetches).  Same specification as surfel_fit_kernel (counts are integers; the fp64 PCA sums are rounded to float before they
// are used), which keeps the larger leaves.
constexpr int kWarpFitMax = 768;
__device__ __forceinline__ double warp_sum_double(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__global__ void __launch_bounds__(kFitThreads) surfel_fit_warp_kernel(const float4* __restrict__ pts, const int32_t* __restrict__ leaf_start,
                                                                      const int32_t* __restrict__ leaf_key, const int32_t* __restrict__ cand, int n_cand,
                                                                      float thr, int min_inliers, FitOut* __restrict__ out) {
  __shared__ float s_xyz[kFitThreads / 32][3][kWarpFitMax];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_warps = (gridDim.x * kFitThreads) >> 5;
  float* sx = s_xyz[warp][0]; float* sy = s_xyz[warp][1]; float* sz = s_xyz[warp][2];
  for (int ci = (blockIdx.x * kFitThreads + threadIdx.x) >> 5; ci < n_cand; ci += n_warps) {
    const int leaf = cand[ci];
    const int beg = leaf_start[leaf];
    const uint32_t n = static_cast<uint32_t>(leaf_start[leaf + 1] - beg);
    if (n > static_cast<uint32_t>(kWarpFitMax)) continue;   // surfel_fit_kernel's
    const float4* P = pts + beg;
    __syncwarp();
    for (uint32_t i = lane; i < n; i += 32) { const float4 p = __ldg(P + i); sx[i] = p.x; sy[i] = p.y; sz[i] = p.z; }
    __syncwarp();
    auto count_inliers = [&](float c0, float c1, float c2, float c3) {
      int count = 0;
      for (uint32_t i = lane; i < n; i += 32) {
        float t = c0 * sx[i];
        t = t + c1 * sy[i];
        t = t + c2 * sz[i];
        t = t + c3;
        count += fabsf(t) < thr ? 1 : 0;
      }
      return __reduce_add_sync(0xffffffffu, count);
    };
    FitOut fo;
    fo.ok = 0; fo.ninl = 0;
    for (int k = 0; k < 4; ++k) fo.p4[k] = 0;
    for (int k = 0; k < 3; ++k) { fo.bmin[k] = 0; fo.bmax[k] = 0; }
    bool fail = n < 3;
    float best[4] = {0, 0, 0, 0};
    int best_count = 0;
    if (!fail) {
      const int max_iterations = 50;
      const double log_probability = log(1.0 - 0.99);
      const double one_over_n = 1.0 / static_cast<double>(n);
      const uint64_t seed = mix64(static_cast<uint64_t>(static_cast<uint32_t>(leaf_key[leaf])) * 0x2545F4914F6CDD1Dull + 12345ull);
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
        const uint32_t lo = min(a, b), hi = max(a, b);
        if (c >= lo) ++c;
        if (c >= hi) ++c;
        float coef[4];
        if (!model_from3(make_float4(sx[a], sy[a], sz[a], 0.f), make_float4(sx[b], sy[b], sz[b], 0.f), make_float4(sx[c], sy[c], sz[c], 0.f), coef)) { ++skipped; continue; }
        const int count = count_inliers(coef[0], coef[1], coef[2], coef[3]);
        if (count > best_count) {
          best_count = count;
          best[0] = coef[0]; best[1] = coef[1]; best[2] = coef[2]; best[3] = coef[3];
          const double w = static_cast<double>(count) * one_over_n;
          double p_no = 1.0 - w * w * w;
          p_no = fmax(2.220446049250313e-16, p_no);
          p_no = fmin(1.0 - 2.220446049250313e-16, p_no);
          k = log_probability / log(p_no);
        }
        ++iterations;
        if (iterations > max_iterations) break;
      }
      if (best_count == 0) fail = true;
    }
    float fin[4] = {best[0], best[1], best[2], best[3]};
    if (!fail && best_count > 3) {
      uint32_t first = 0xffffffffu;
      for (uint32_t i = lane; i < n && first == 0xffffffffu; i += 32)
        if (plane_dist(best, make_float4(sx[i], sy[i], sz[i], 0.f)) < thr) first = i;
      first = __reduce_min_sync(0xffffffffu, first);
      const double x0 = sx[first], y0 = sy[first], z0 = sz[first];
      double sums[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      int cnt = 0;
      for (uint32_t i = lane; i < n; i += 32) {
        const float4 p = make_float4(sx[i], sy[i], sz[i], 0.f);
        if (!(plane_dist(best, p) < thr)) continue;
        const double dx = static_cast<double>(p.x) - x0, dy = static_cast<double>(p.y) - y0, dz = static_cast<double>(p.z) - z0;
        sums[0] += dx; sums[1] += dy; sums[2] += dz;
        sums[3] += dx * dx; sums[4] += dx * dy; sums[5] += dx * dz; sums[6] += dy * dy; sums[7] += dy * dz; sums[8] += dz * dz;
        ++cnt;
      }
#pragma unroll
      for (int q = 0; q < 9; ++q) sums[q] = warp_sum_double(sums[q]);
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      const double inv_n = 1.0 / cnt;
      const double m0 = sums[0] * inv_n, m1 = sums[1] * inv_n, m2 = sums[2] * inv_n;
      const float cf0 = static_cast<float>(x0 + m0), cf1 = static_cast<float>(y0 + m1), cf2 = static_cast<float>(z0 + m2);
      const float cv0 = static_cast<float>(sums[3] * inv_n - m0 * m0), cv1 = static_cast<float>(sums[4] * inv_n - m0 * m1),
                  cv2 = static_cast<float>(sums[5] * inv_n - m0 * m2), cv3 = static_cast<float>(sums[6] * inv_n - m1 * m1),
                  cv4 = static_cast<float>(sums[7] * inv_n - m1 * m2), cv5 = static_cast<float>(sums[8] * inv_n - m2 * m2);
      const double A[9] = {cv0, cv1, cv2, cv1, cv3, cv4, cv2, cv4, cv5};
      double ev[3], V[9];
      jacobi3_lower(A, ev, V);
      const double nx = V[0], ny = V[3], nz = V[6];
      double d = nx * static_cast<double>(cf0);
      d = d + ny * static_cast<double>(cf1);
      d = d + nz * static_cast<double>(cf2);
      fin[0] = static_cast<float>(nx); fin[1] = static_cast<float>(ny); fin[2] = static_cast<float>(nz);
      fin[3] = static_cast<float>(-d);
    }
    if (!fail) {
      int cnt2 = 0;
      float mn0 = 3.402823466e38f, mn1 = mn0, mn2 = mn0, mx0 = -mn0, mx1 = -mn0, mx2 = -mn0;
      for (uint32_t i = lane; i < n; i += 32) {
        const float4 p = make_float4(sx[i], sy[i], sz[i], 0.f);
        cnt2 += plane_dist(fin, p) < thr ? 1 : 0;
        mn0 = fminf(mn0, p.x); mn1 = fminf(mn1, p.y); mn2 = fminf(mn2, p.z);
        mx0 = fmaxf(mx0, p.x); mx1 = fmaxf(mx1, p.y); mx2 = fmaxf(mx2, p.z);
      }
      cnt2 = __reduce_add_sync(0xffffffffu, cnt2);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mn0 = fminf(mn0, __shfl_xor_sync(0xffffffffu, mn0, o)); mn1 = fminf(mn1, __shfl_xor_sync(0xffffffffu, mn1, o)); mn2 = fminf(mn2, __shfl_xor_sync(0xffffffffu, mn2, o));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, o)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, o)); mx2 = fmaxf(mx2, __shfl_xor_sync(0xffffffffu, mx2, o));
      }
      fo.ninl = cnt2;
      fo.ok = cnt2 >= min_inliers;
      for (int k = 0; k < 4; ++k) fo.p4[k] = fin[k];
      fo.bmin[0] = mn0; fo.bmin[1] = mn1; fo.bmin[2] = mn2; fo.bmax[0] = mx0; fo.bmax[1] = mx1; fo.bmax[2] = mx2;
    }
    if (lane == 0) out[ci] = fo;
  }
}

__global__ void __launch_bounds__(256) surfel_flag_ok_kernel(const FitOut* __restrict__ fit, int n, int32_t* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = fit[i].ok;
}

__global__ void __launch_bounds__(256) surfel_pack_kernel(const FitOut* __restrict__ fit, const int32_t* __restrict__ cand, const int32_t* __restrict__ pos,
                                                          const int32_t* __restrict__ leaf_key, int32_t* __restrict__ key_out,
                                                          int n_cand, double* __restrict__ p4, double* __restrict__ Pi, double* __restrict__ bmin,
                                                          double* __restrict__ bmax, int32_t* __restrict__ leaf, int32_t* __restrict__ ninl,
                                                          int32_t* __restrict__ leaf2plane) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cand) return;
  const FitOut f = fit[i];
  if (!f.ok) return;
  const int p = pos[i];
  for (int k = 0; k < 4; ++k) p4[4 * p + k] = f.p4[k];
  for (int k = 0; k < 3; ++k) {
    Pi[3 * p + k] = -f.p4[3] * f.p4[k];  // surfel_association.cpp:79
    bmin[3 * p + k] = f.bmin[k]; bmax[3 * p + k] = f.bmax[k];
  }
  leaf[p] = cand[i]; ninl[p] = f.ninl; key_out[p] = leaf_key[cand[i]];
  leaf2plane[cand[i]] = p;
}

static int exclusive_scan_count(lvi_ctx* ctx, const int32_t* flag, int32_t* pos, int n) {
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, flag, pos, n, ctx->stream);
  DBuf<char> tmp(tb + 16);
  LVI_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, flag, pos, n, ctx->stream));
  ctx->launches += 2;
  int last_pos = 0, last_flag = 0;
  LVI_CUDA(cudaMemcpyAsync(&last_pos, pos + n - 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
  LVI_CUDA(cudaMemcpyAsync(&last_flag, flag + n - 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
  LVI_CUDA(cudaStreamSynchronize(ctx->stream));
  return last_pos + last_flag;
}

}  // namespace lvi

using namespace lvi;

extern "C" {

int lvi_surfel_extract(lvi_ctx* ctx, const lvi_voxel_map* m, double lambda, int min_leaf_points, float ransac_threshold, int min_inliers,
                       lvi_surfel_set** out) {
  return guarded([&] {
    LVI_REQUIRE(ctx && m && out, LVI_ERR_INVALID, "lvi_surfel_extract: null argument");
    activate(ctx);
    auto s = std::unique_ptr<lvi_surfel_set>(new lvi_surfel_set());
    s->ctx = ctx;
    const int L = static_cast<int>(m->n_leaves);
    s->leaf2plane.alloc(std::max(L, 1));
    LVI_CUDA(cudaMemsetAsync(s->leaf2plane.p, 0xff, sizeof(int32_t) * std::max(L, 1), ctx->stream));
    if (L == 0) { *out = s.release(); return; }
    DBuf<int32_t> flag(L), pos(L);
    LVI_LAUNCH(ctx, surfel_candidate_kernel, (L + 255) / 256, 256, 0, m->leaf_npts.p, m->leaf_evals.p, L, min_leaf_points, lambda, flag.p);
    const int n_cand = exclusive_scan_count(ctx, flag.p, pos.p, L);
    if (n_cand == 0) { *out = s.release(); return; }
    DBuf<int32_t> cand(n_cand);
    LVI_LAUNCH(ctx, surfel_scatter_ids_kernel, (L + 255) / 256, 256, 0, flag.p, pos.p, L, cand.p);
    DBuf<FitOut> fit(n_cand);
    // leaves of up to 768 points: one warp each, points in registers; larger ones: one CTA each (both kernels walk the candidate list and
    // take their own leaves).  LVI_SURFEL_WARP_FIT=0 sends every leaf through the CTA kernel (diagnostics).
    static const bool warp_fit = !(std::getenv("LVI_SURFEL_WARP_FIT") && std::atoi(std::getenv("LVI_SURFEL_WARP_FIT")) == 0);
    const uint32_t small_max = warp_fit ? static_cast<uint32_t>(kWarpFitMax) : 0u;
    if (warp_fit)
      LVI_LAUNCH(ctx, surfel_fit_warp_kernel, std::min((n_cand + 3) / 4, ctx->sm_count * 24), kFitThreads, 0, m->pts_sorted.p, m->leaf_start.p,
                 m->leaf_key.p, cand.p, n_cand, ransac_threshold, min_inliers, fit.p);
    // leaves above kClusterFitMin points: a cluster of 8 CTAs each (LVI_SURFEL_CLUSTER_FIT=0: the one-CTA kernel takes them, diagnostics)
    static const bool cluster_fit = !(std::getenv("LVI_SURFEL_CLUSTER_FIT") && std::atoi(std::getenv("LVI_SURFEL_CLUSTER_FIT")) == 0);
    const uint32_t large_min = cluster_fit ? kClusterFitMin : 0xffffffffu;
    if (cluster_fit)
      LVI_LAUNCH(ctx, surfel_fit_cluster_kernel, kFitCluster * std::min(n_cand, ctx->sm_count * 2), kFitThreads, 0, m->pts_sorted.p, m->leaf_start.p,
                 m->leaf_key.p, cand.p, n_cand, ransac_threshold, min_inliers, fit.p);
    LVI_LAUNCH(ctx, surfel_fit_kernel, std::min(n_cand, ctx->sm_count * 16), kFitThreads, 0, m->pts_sorted.p, m->leaf_start.p,
               m->leaf_key.p, cand.p, n_cand, ransac_threshold, min_inliers, small_max, large_min, fit.p);
    DBuf<int32_t> okf(n_cand), okpos(n_cand);
    LVI_LAUNCH(ctx, surfel_flag_ok_kernel, (n_cand + 255) / 256, 256, 0, fit.p, n_cand, okf.p);
    const int P = exclusive_scan_count(ctx, okf.p, okpos.p, n_cand);
    s->n_planes = P;
    const int Pa = std::max(P, 1);
    s->p4.alloc(4 * Pa); s->Pi.alloc(3 * Pa); s->bmin.alloc(3 * Pa); s->bmax.alloc(3 * Pa); s->leaf.alloc(Pa); s->ninl.alloc(Pa); s->key.alloc(Pa);
    if (P) LVI_LAUNCH(ctx, surfel_pack_kernel, (n_cand + 255) / 256, 256, 0, fit.p, cand.p, okpos.p, m->leaf_key.p, s->key.p, n_cand, s->p4.p, s->Pi.p, s->bmin.p, s->bmax.p,
                      s->leaf.p, s->ninl.p, s->leaf2plane.p);
    LVI_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = s.release();
  });
}

int lvi_surfel_destroy(lvi_surfel_set* s) {
  delete s;
  return LVI_OK;
}
int64_t lvi_surfel_count(const lvi_surfel_set* s) { return s ? s->n_planes : 0; }

int lvi_surfel_export(lvi_ctx* ctx, const lvi_surfel_set* s, double* p4, double* Pi, double* box_min, double* box_max, int64_t* leaf_key,
                      int32_t* n_inliers) {
  return guarded([&] {
    LVI_REQUIRE(ctx && s, LVI_ERR_INVALID, "lvi_surfel_export: null argument");
    activate(ctx);
    const size_t P = static_cast<size_t>(s->n_planes);
    cudaStream_t st = ctx->stream;
    if (p4) s->p4.download(p4, 4 * P, st);
    if (Pi) s->Pi.download(Pi, 3 * P, st);
    if (box_min) s->bmin.download(box_min, 3 * P, st);
    if (box_max) s->bmax.download(box_max, 3 * P, st);
    if (n_inliers) s->ninl.download(n_inliers, P, st);
    std::vector<int32_t> key(P);
    if (leaf_key) s->key.download(key.data(), P, st);
    LVI_CUDA(cudaStreamSynchronize(st));
    if (leaf_key) for (size_t i = 0; i < P; ++i) leaf_key[i] = key[i];
  });
}

}  // extern "C"
