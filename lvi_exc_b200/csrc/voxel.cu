// voxel.cu — NDT voxel-grid covariance / eigen build on sm_100a.
//
// Replaces pclomp::VoxelGridCovariance::applyFilter (N/voxel_grid_covariance_omp_impl.hpp:49-374), which is serial,
// inserts every point into a std::map and copies it into Leaf::pointList_.
//
// The input is a scan batch (map.cuh): packed float4 points in scan order with the float min/max of every scan already known, so there
// is no min/max pass.  A spinning LiDAR sweeps one surface for many consecutive firings, so consecutive points of a ring mostly share
// their voxel: the build works on RUNS (maximal stretches of consecutive finite points with one voxel index), not on points.
//   1. grid_params : min_b_/div_b_/divb_mul_ in the reference's FLOAT arithmetic (:87-103, Q2), overflow guard (:75-84), from the
//                    min/max of the selected scans
//   2. runs        : one streaming pass (16 B/point, coalesced): voxel index per point (bit-exact with the reference), run heads and
//                    lengths from a 1024-point tile's boundary bitmap in shared memory, appended as (index << 32 | first point, length)
//   3. sort        : stable order = std::map order of the leaves and cloud order inside a leaf (Leaf::pointList_): radix sort of the
//                    RUN records (CUB; 20 - 30x fewer records than points on real scans, one per point at worst)
//   4. leaves      : run-length of the sorted voxel indices -> CSR leaves; exclusive sum of the run lengths -> point offsets
//   5. gather      : one warp per run copies its points (coalesced on both sides) into leaf order, tagging each with its original cloud
//                    index, and leaves the run's fp64 partial sums {sum x, sum x x^T}
//   6. leaf stats  : per leaf the partial sums are added in run order (deterministic), then mean/cov (Q1 identity seed, :333-334),
//                    Jacobi eigen-decomposition, eigenvalue inflation (:349-360) and inverse covariance (:363-368), one thread per leaf
// Each point crosses HBM three times (read for the index, read + written by the gather) = 48 B against the 12 B of its coordinates
// (DESIGN.md §4); the previous pipeline (compact copy, keys, 3-pass pair sort, gather, stats) moved ~190 B per point.
// Compiled with -fmad=false: index arithmetic must match the CPU bit for bit.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstring>
#include <memory>

#include "eig3.cuh"
#include "map.cuh"

namespace lvi {

// min/max of the selected scans of a batch -> mm[6] (ordered ints)
__global__ void voxel_reduce_minmax_kernel(const int* __restrict__ scan_mm, const int* __restrict__ scan_base, int n_scans, int* __restrict__ mm) {
  __shared__ int sh[6][32];
  int v[6];
  for (int k = 0; k < 3; ++k) { v[k] = f2ord(3.402823466e38f); v[3 + k] = f2ord(-3.402823466e38f); }
  for (int s = threadIdx.x; s < n_scans; s += blockDim.x) {
    if (scan_base && scan_base[s] < 0) continue;
    for (int k = 0; k < 3; ++k) { v[k] = min(v[k], scan_mm[6 * s + k]); v[3 + k] = max(v[3 + k], scan_mm[6 * s + 3 + k]); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    for (int k = 0; k < 3; ++k) { v[k] = min(v[k], __shfl_xor_sync(0xffffffffu, v[k], o)); v[3 + k] = max(v[3 + k], __shfl_xor_sync(0xffffffffu, v[3 + k], o)); }
  if ((threadIdx.x & 31) == 0) for (int k = 0; k < 6; ++k) sh[k][threadIdx.x >> 5] = v[k];
  __syncthreads();
  if (threadIdx.x < 6) {
    const int k = threadIdx.x;
    int r = sh[k][0];
    for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w) r = k < 3 ? min(r, sh[k][w]) : max(r, sh[k][w]);
    mm[k] = r;
  }
}

// ---- 2. grid parameters (float arithmetic of the reference) -------------------------------------------------------
__global__ void voxel_grid_params_kernel(const int* __restrict__ mm, float leaf, GridParams* g) {
  GridParams p;
  const float inv = 1.0f / leaf;  // pcl::VoxelGrid::setLeafSize
  p.inv_leaf = inv;
  for (int k = 0; k < 3; ++k) { p.mn[k] = ord2f(mm[k]); p.mx[k] = ord2f(mm[3 + k]); }
  p.status = 0; p.key_bits = 1; p.ncell = 0;
  for (int k = 0; k < 3; ++k) { p.min_b[k] = 0; p.div_b[k] = 0; p.mul[k] = 0; }
  if (p.mn[0] > p.mx[0]) { p.status = 1; *g = p; return; }
  const long long dx = static_cast<long long>((p.mx[0] - p.mn[0]) * inv) + 1;  // impl.hpp:75-77
  const long long dy = static_cast<long long>((p.mx[1] - p.mn[1]) * inv) + 1;
  const long long dz = static_cast<long long>((p.mx[2] - p.mn[2]) * inv) + 1;
  // the reference multiplies in int64 (:80); the double product guards the cases where that itself would wrap
  if (static_cast<double>(dx) * static_cast<double>(dy) * static_cast<double>(dz) > 2147483647.0 || dx * dy * dz > 2147483647LL) { p.status = 2; *g = p; return; }
  for (int k = 0; k < 3; ++k) {
    p.min_b[k] = static_cast<int>(floorf(p.mn[k] * inv));  // :87-92
    const int max_b = static_cast<int>(floorf(p.mx[k] * inv));
    p.div_b[k] = max_b - p.min_b[k] + 1;
  }
  p.mul[0] = 1; p.mul[1] = p.div_b[0]; p.mul[2] = p.div_b[0] * p.div_b[1];  // :103
  p.ncell = static_cast<long long>(p.div_b[0]) * p.div_b[1] * p.div_b[2];
  int bits = 1;
  while ((1LL << bits) <= p.ncell) ++bits;  // the sentinel key == ncell must be representable
  p.key_bits = bits;
  *g = p;
}

// ---- 2. runs --------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t voxel_key(const GridParams& g, float x, float y, float z) {
  // impl.hpp:220-225: static_cast<int>(floor(x * inverse_leaf_size_[0]) - static_cast<float>(min_b_[0])) ... ; idx = ijk . divb_mul_
  const int i0 = static_cast<int>(floorf(x * g.inv_leaf) - static_cast<float>(g.min_b[0]));
  const int i1 = static_cast<int>(floorf(y * g.inv_leaf) - static_cast<float>(g.min_b[1]));
  const int i2 = static_cast<int>(floorf(z * g.inv_leaf) - static_cast<float>(g.min_b[2]));
  return static_cast<uint32_t>(i0 * g.mul[0] + i1 * g.mul[1] + i2 * g.mul[2]);
}

constexpr int kRunTile = 1024;            // points per tile: runs are cut at tile boundaries (one extra record per 1024 points)
constexpr uint32_t kNoKey = 0xffffffffu;  // non-finite point / point of a scan that is not part of the map

// One CTA per tile.  scan_base (optional): per scan of the batch, the index its first point has in the cloud the map is built from, or
// -1 when the scan is left out (LiDAROdometry only feeds key scans to the map, L/src/core/lidar_odometry.cpp:89-104).
__global__ void __launch_bounds__(256) voxel_runs_kernel(const float4* __restrict__ pts, int64_t n, int64_t pts_per_scan, const int* __restrict__ scan_base,
                                                         const GridParams* __restrict__ gp, unsigned long long* __restrict__ run_key,
                                                         uint32_t* __restrict__ run_len, unsigned int* __restrict__ n_runs) {
  __shared__ GridParams g;
  __shared__ uint32_t skey[kRunTile + 1];
  __shared__ uint32_t sbnd[kRunTile / 32 + 1];   // bit b of word w: position 32 w + b starts a run or is not a point of the map
  if (threadIdx.x == 0) g = *gp;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n_tiles = (n + kRunTile - 1) / kRunTile;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();
    const int64_t base = tile * kRunTile;
    const int cnt = static_cast<int>(min(static_cast<int64_t>(kRunTile), n - base));
#pragma unroll
    for (int j = 0; j < kRunTile / 256; ++j) {
      const int p = tid + 256 * j;
      uint32_t key = kNoKey;
      if (p < cnt) {
        const float4 v = __ldcs(pts + base + p);   // streamed: read once here, once more by the gather
        const bool in_map = !scan_base || scan_base[(base + p) / pts_per_scan] >= 0;
        if (in_map && isfinite(v.x) && isfinite(v.y) && isfinite(v.z)) key = voxel_key(g, v.x, v.y, v.z);
      }
      skey[p] = key;
    }
    if (tid == 0) { skey[kRunTile] = kNoKey; sbnd[kRunTile / 32] = 0xffffffffu; }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kRunTile / 256; ++j) {   // warp `warp` builds words warp, warp + 8, ...
      const int w = warp + 8 * j, p = 32 * w + lane;
      const uint32_t k = skey[p];
      const bool bnd = k == kNoKey || p == 0 || skey[p - 1] != k;
      const unsigned m = __ballot_sync(0xffffffffu, bnd);
      if (lane == 0) sbnd[w] = m;
    }
    __syncthreads();
    // every thread owns 4 consecutive positions; a position with a valid index and a boundary bit is a run head
    int heads = 0;
    uint32_t hk[4]; int hp[4], hl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int p = 4 * tid + j;
      const uint32_t k = skey[p];
      if (k == kNoKey || !((sbnd[p >> 5] >> (p & 31)) & 1u)) continue;
      // length: distance to the next boundary bit (the tile end is one)
      int q = p + 1, w = q >> 5;
      uint32_t word = sbnd[w] & (0xffffffffu << (q & 31));
      while (word == 0u) word = sbnd[++w];
      const int end = min(32 * w + __ffs(word) - 1, cnt);
      hk[heads] = k; hp[heads] = p; hl[heads] = end - p; ++heads;
    }
    // one atomic per warp reserves the record slots (their order is irrelevant: the records are sorted by (index, first point) next)
    int off = heads;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, off, o); if (lane >= o) off += t; }
    const int total = __shfl_sync(0xffffffffu, off, 31);
    unsigned int slot0 = 0;
    if (lane == 31 && total > 0) slot0 = atomicAdd(n_runs, static_cast<unsigned int>(total));
    slot0 = __shfl_sync(0xffffffffu, slot0, 31) + static_cast<unsigned int>(off - heads);
    for (int j = 0; j < heads; ++j) {
      run_key[slot0 + j] = (static_cast<unsigned long long>(hk[j]) << 32) | static_cast<unsigned long long>(static_cast<uint32_t>(base + hp[j]));
      run_len[slot0 + j] = static_cast<uint32_t>(hl[j]);
    }
  }
}

// ---- 4. leaves ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) voxel_run_flag_kernel(const unsigned long long* __restrict__ run_key, int n_runs, int32_t* __restrict__ flag) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_runs) flag[r] = (r == 0 || (run_key[r] >> 32) != (run_key[r - 1] >> 32)) ? 1 : 0;
}
// leaf_of_run is the INCLUSIVE sum of the head flags (1-based leaf number); run_off the exclusive sum of the lengths
__global__ void __launch_bounds__(256) voxel_leaf_bounds_kernel(const unsigned long long* __restrict__ run_key, const uint32_t* __restrict__ run_len,
                                                                const int32_t* __restrict__ flag, const int32_t* __restrict__ leaf_of_run,
                                                                const int32_t* __restrict__ run_off, int n_runs, int32_t* __restrict__ leaf_key,
                                                                int32_t* __restrict__ leaf_start, int32_t* __restrict__ leaf_run0, int* __restrict__ totals) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_runs) return;
  if (flag[r]) {
    const int l = leaf_of_run[r] - 1;
    leaf_key[l] = static_cast<int32_t>(run_key[r] >> 32);
    leaf_start[l] = run_off[r];
    leaf_run0[l] = r;
  }
  if (r == n_runs - 1) {
    const int L = leaf_of_run[r], nb = run_off[r] + static_cast<int>(run_len[r]);
    leaf_start[L] = nb; leaf_run0[L] = n_runs;
    totals[0] = L; totals[1] = nb;
  }
}

// ---- 5. gather + partial sums: one CTA per tile of 256 OUTPUT points ------------------------------------------------------------
// Output position o (leaf order) belongs to the sorted run r with run_off[r] <= o < run_off[r] + len[r] and comes from source point
// src[r] + (o - run_off[r]).  The runs that overlap a tile are loaded into shared memory (tile_run0 gives the first one) and every thread
// finds its run with a 9-step search there: writes are perfectly coalesced, reads are contiguous stretches of one run each.
// The same pass leaves the fp64 partial sums {sum x, sum x x^T} of every (leaf, tile) pair at slot leaf + tile -- both indices grow
// monotonically along the output, so the pairs form a staircase and the slot is unique; a leaf's partials are then the contiguous slots
// leaf + first_tile ... leaf + last_tile, added in that order by voxel_leaf_sums_kernel (deterministic).
constexpr int kGatherTile = 1024;   // output points per CTA (4 per thread)
constexpr int kGatherThreads = 256;
constexpr int kRunSums = 9;          // sum x,y,z ; sum xx,xy,xz,yy,yz,zz

__global__ void __launch_bounds__(256) voxel_tile_run0_kernel(const int32_t* __restrict__ run_off, const uint32_t* __restrict__ run_len, int n_runs,
                                                              int32_t* __restrict__ tile_run0) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_runs) return;
  const int off = run_off[r], end = off + static_cast<int>(run_len[r]);
  for (int t = (off + kGatherTile - 1) / kGatherTile; t * kGatherTile < end; ++t) tile_run0[t] = r;   // the run that holds the tile's first point
}

__device__ __forceinline__ void warp_reduce9(double (&v)[kRunSums]) {
#pragma unroll
  for (int k = 0; k < kRunSums; ++k) {
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], sh);
  }
}

__global__ void __launch_bounds__(kGatherThreads) voxel_gather_kernel(const float4* __restrict__ pts, int64_t pts_per_scan, const int* __restrict__ scan_base,
                                                                      const unsigned long long* __restrict__ run_key, const int32_t* __restrict__ run_off,
                                                                      const int32_t* __restrict__ leaf_of_run, const int32_t* __restrict__ tile_run0,
                                                                      const int32_t* __restrict__ leaf_start, int n_runs, int n_out,
                                                                      float4* __restrict__ out, double* __restrict__ partial, int idx_from_w) {
  __shared__ int s_off[kGatherTile + 1];
  __shared__ uint32_t s_src[kGatherTile];
  __shared__ float s_x[kGatherTile], s_y[kGatherTile], s_z[kGatherTile];
  __shared__ double s_red[kGatherThreads / 32][kRunSums];
  __shared__ int s_nr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x;
  const int o0 = tile * kGatherTile, o1 = min(o0 + kGatherTile, n_out);
  const int r0 = tile_run0[tile];
  // runs overlapping [o0, o1): at most one per output point
  {
    int mine = 0;
#pragma unroll
    for (int j = 0; j < kGatherTile / kGatherThreads; ++j) {
      const int q = tid + kGatherThreads * j, r = r0 + q;
      int off = 0x7fffffff;
      if (r < n_runs) off = __ldg(run_off + r);
      if (off < o1) { s_off[q] = off; s_src[q] = static_cast<uint32_t>(__ldg(run_key + r)); ++mine; }
    }
    // run offsets are increasing, so the in-range runs are a prefix: their number is a block sum
    int tot = mine;
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, sh);
    __shared__ int s_cnt[kGatherThreads / 32];
    if (lane == 0) s_cnt[warp] = tot;
    __syncthreads();
    if (tid == 0) { int a = 0; for (int w = 0; w < kGatherThreads / 32; ++w) a += s_cnt[w]; s_nr = a; s_off[a] = 0x7fffffff; }
  }
  __syncthreads();
  const int nr = s_nr;
  const int l0 = __ldg(leaf_of_run + r0) - 1, l1 = __ldg(leaf_of_run + r0 + nr - 1) - 1;
  double v[kRunSums];
#pragma unroll
  for (int k = 0; k < kRunSums; ++k) v[k] = 0.0;
#pragma unroll
  for (int j = 0; j < kGatherTile / kGatherThreads; ++j) {
    const int q = tid + kGatherThreads * j, o = o0 + q;
    if (o >= o1) continue;
    int lo = 0, hi = nr - 1;   // last run with s_off <= o
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (s_off[mid] <= o) lo = mid; else hi = mid - 1; }
    const uint32_t sp = s_src[lo] + static_cast<uint32_t>(o - s_off[lo]);
    const float4 p = __ldcs(pts + sp);
    int cidx = static_cast<int>(sp);   // index in the cloud the map is built from (Leaf::pointList_ consumers index by it)
    if (idx_from_w) cidx = __float_as_int(p.w);   // sharded build: the point carries its index in the global map cloud
    else if (scan_base) { const uint32_t sc = sp / static_cast<uint32_t>(pts_per_scan); cidx = scan_base[sc] + static_cast<int>(sp - sc * static_cast<uint32_t>(pts_per_scan)); }
    __stcs(out + o, make_float4(p.x, p.y, p.z, __int_as_float(cidx)));
    s_x[q] = p.x; s_y[q] = p.y; s_z[q] = p.z;
    const double x = p.x, y = p.y, z = p.z;
    v[0] += x; v[1] += y; v[2] += z; v[3] += x * x; v[4] += x * y; v[5] += x * z; v[6] += y * y; v[7] += y * z; v[8] += z * z;
  }
  if (l0 == l1) {   // the whole tile lies in one leaf: block reduction of the per-thread sums
    warp_reduce9(v);
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < kRunSums; ++k) s_red[warp][k] = v[k];
    }
    __syncthreads();
    if (tid < kRunSums) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < kGatherThreads / 32; ++w) t += s_red[w][tid];
      partial[static_cast<size_t>(l0 + tile) * kRunSums + tid] = t;
    }
  } else if (l1 - l0 >= 32) {   // many small leaves in the tile (millions of leaves of a few points): one THREAD per leaf, in point order
    __syncthreads();
    for (int l = l0 + tid; l <= l1; l += kGatherThreads) {
      const int a = max(__ldg(leaf_start + l), o0) - o0, e = min(__ldg(leaf_start + l + 1), o1) - o0;
      double t[kRunSums];
#pragma unroll
      for (int k = 0; k < kRunSums; ++k) t[k] = 0.0;
      for (int i = a; i < e; ++i) {
        const double x = s_x[i], y = s_y[i], z = s_z[i];
        t[0] += x; t[1] += y; t[2] += z; t[3] += x * x; t[4] += x * y; t[5] += x * z; t[6] += y * y; t[7] += y * z; t[8] += z * z;
      }
#pragma unroll
      for (int k = 0; k < kRunSums; ++k) partial[static_cast<size_t>(l + tile) * kRunSums + k] = t[k];
    }
  } else {          // a few leaves in the tile: one warp per leaf adds up the tile's share of it from shared memory
    __syncthreads();
    for (int l = l0 + warp; l <= l1; l += kGatherThreads / 32) {
      const int a = max(__ldg(leaf_start + l), o0) - o0, e = min(__ldg(leaf_start + l + 1), o1) - o0;
      double t[kRunSums];
#pragma unroll
      for (int k = 0; k < kRunSums; ++k) t[k] = 0.0;
      for (int i = a + lane; i < e; i += 32) {
        const double x = s_x[i], y = s_y[i], z = s_z[i];
        t[0] += x; t[1] += y; t[2] += z; t[3] += x * x; t[4] += x * y; t[5] += x * z; t[6] += y * y; t[7] += y * z; t[8] += z * z;
      }
      warp_reduce9(t);
      if (lane < kRunSums) {
        const double tv = lane == 0 ? t[0] : lane == 1 ? t[1] : lane == 2 ? t[2] : lane == 3 ? t[3] : lane == 4 ? t[4] : lane == 5 ? t[5] : lane == 6 ? t[6] : lane == 7 ? t[7] : t[8];
        partial[static_cast<size_t>(l + tile) * kRunSums + lane] = tv;
      }
    }
  }
}

// ---- 6. per-leaf statistics ---------------------------------------------------------------------------------------------
// (a) the (leaf, tile) partial sums of a leaf are added in tile order by a group of 8 lanes (lane j takes tiles j, j+8, ...; the 8 sums
//     are then combined in a fixed tree): deterministic
__global__ void __launch_bounds__(256) voxel_leaf_sums_kernel(const double* __restrict__ partial, const int32_t* __restrict__ leaf_start, int n_leaves,
                                                              double* __restrict__ leaf_sums) {
  const int sub = threadIdx.x & 7;
  const int leaf = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const bool ok = leaf < n_leaves;
  double s[kRunSums];
#pragma unroll
  for (int k = 0; k < kRunSums; ++k) s[k] = 0.0;
  if (ok) {
    const int t0 = leaf_start[leaf] / kGatherTile, t1 = (leaf_start[leaf + 1] - 1) / kGatherTile;
    for (int t = t0 + sub; t <= t1; t += 8) {
      const double* p = partial + static_cast<size_t>(leaf + t) * kRunSums;
#pragma unroll
      for (int k = 0; k < kRunSums; ++k) s[k] += p[k];
    }
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1)
#pragma unroll
    for (int k = 0; k < kRunSums; ++k) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
  if (ok && sub == 0)
#pragma unroll
    for (int k = 0; k < kRunSums; ++k) leaf_sums[static_cast<size_t>(leaf) * kRunSums + k] = s[k];
}

// (b) one thread per leaf: mean, covariance, eigen-decomposition, inflation, inverse (all lanes busy: the 3x3 Jacobi sweeps are the bulk
//     of the arithmetic when the map has millions of leaves)
__global__ void __launch_bounds__(128) voxel_leaf_stats_kernel(const double* __restrict__ leaf_sums, const int32_t* __restrict__ leaf_start,
                                                               int n_leaves, int min_points, double eig_mult, int32_t* __restrict__ npts_out,
                                                               double* __restrict__ mean_out, double* __restrict__ cov_out,
                                                               double* __restrict__ evals_out, double* __restrict__ evecs_out,
                                                               double* __restrict__ icov_out) {
  const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
  if (leaf >= n_leaves) return;
  const double* S = leaf_sums + static_cast<size_t>(leaf) * kRunSums;
  // Leaf() seeds cov_ with Identity (N/voxel_grid_covariance_omp.h:97-106, Q1)
  const double sx = S[0], sy = S[1], sz = S[2], sxx = S[3] + 1.0, sxy = S[4], sxz = S[5], syy = S[6] + 1.0, syz = S[7], szz = S[8] + 1.0;
  int n = leaf_start[leaf + 1] - leaf_start[leaf];
  const double nn = n;
  const double pt_sum[3] = {sx, sy, sz};
  double mean[3] = {sx / nn, sy / nn, sz / nn};  // impl.hpp:299
  double cov[9] = {sxx, sxy, sxz, sxy, syy, syz, sxz, syz, szz};
  double evals[3] = {0, 0, 0}, evecs[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, icov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (n >= min_points) {
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c)  // :333
        cov[r * 3 + c] = (cov[r * 3 + c] - 2 * (pt_sum[r] * mean[c])) / nn + mean[r] * mean[c];
    for (int k = 0; k < 9; ++k) cov[k] *= (nn - 1.0) / nn;  // :334
    double ev[3];
    jacobi3_lower(cov, ev, evecs);  // :337-339
    if (ev[0] < 0 || ev[1] < 0 || ev[2] <= 0) {
      n = -1;  // :341-345
    } else {
      const double min_ev = eig_mult * ev[2];  // :349
      if (ev[0] < min_ev) {
        ev[0] = min_ev;
        if (ev[1] < min_ev) ev[1] = min_ev;
        double Vi[9], VD[9];
        inv3_cofactor(evecs, Vi);
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) VD[r * 3 + c] = evecs[r * 3 + c] * ev[c];
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c)  // :359
            cov[r * 3 + c] = VD[r * 3 + 0] * Vi[0 * 3 + c] + VD[r * 3 + 1] * Vi[1 * 3 + c] + VD[r * 3 + 2] * Vi[2 * 3 + c];
      }
      for (int k = 0; k < 3; ++k) evals[k] = ev[k];
      inv3_cofactor(cov, icov);  // :363
      double mxc = icov[0], mnc = icov[0];
      for (int k = 1; k < 9; ++k) { mxc = fmax(mxc, icov[k]); mnc = fmin(mnc, icov[k]); }
      if (mxc == static_cast<double>(__int_as_float(0x7f800000)) || mnc == -static_cast<double>(__int_as_float(0x7f800000))) n = -1;  // :364-368
    }
  }
  npts_out[leaf] = n;
  for (int k = 0; k < 3; ++k) { mean_out[leaf * 3 + k] = mean[k]; evals_out[leaf * 3 + k] = evals[k]; }
  for (int k = 0; k < 9; ++k) { cov_out[leaf * 9 + k] = cov[k]; evecs_out[leaf * 9 + k] = evecs[k]; icov_out[leaf * 9 + k] = icov[k]; }
}

__global__ void __launch_bounds__(256) voxel_cell2leaf_kernel(const int32_t* __restrict__ keys, int n_leaves, int32_t* __restrict__ table) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_leaves) table[keys[i]] = i;
}

// ------------------------------------------------------------------------------------------------------------------
// scan_keep (host, optional): which scans of the batch make up the map cloud
lvi_voxel_map* voxel_build_from_batch(lvi_ctx* ctx, const lvi_scan_batch* b, const uint8_t* scan_keep, float leaf, int min_points, double eig_mult) {
  return voxel_build_core(ctx, b, scan_keep, leaf, min_points, eig_mult, VoxelBuildOptions{});
}

// opt.forced_grid_d: grid of a map this cloud is a PART of (sharded build: min_b_ / div_b_ must be those of the whole cloud);
// opt.idx_from_w: the w component of every point is its index in the whole map cloud
lvi_voxel_map* voxel_build_core(lvi_ctx* ctx, const lvi_scan_batch* b, const uint8_t* scan_keep, float leaf, int min_points, double eig_mult,
                                const VoxelBuildOptions& opt) {
  const int64_t n = b->n;
  LVI_REQUIRE(n > 0 && n < 2147483647LL, LVI_ERR_INVALID, "lvi_voxel_build: n_points must be in (0, 2^31)");
  LVI_REQUIRE(leaf > 0, LVI_ERR_INVALID, "lvi_voxel_build: leaf_size must be positive");
  auto m = std::unique_ptr<lvi_voxel_map>(new lvi_voxel_map());
  m->ctx = ctx; m->leaf_size = leaf; m->min_points = min_points; m->eig_mult = eig_mult;
  cudaStream_t st = ctx->stream;
  DBuf<int> scan_base;
  int64_t n_cloud = n;
  if (scan_keep) {
    std::vector<int> hb(b->n_scans);
    int64_t acc = 0;
    for (int s = 0; s < b->n_scans; ++s) {
      const int64_t cnt = std::min<int64_t>(b->pts_per_scan, n - static_cast<int64_t>(s) * b->pts_per_scan);
      if (scan_keep[s]) { hb[s] = static_cast<int>(acc); acc += cnt; } else hb[s] = -1;
    }
    n_cloud = acc;
    LVI_REQUIRE(n_cloud > 0, LVI_ERR_INVALID, "lvi_voxel_build: no scan selected");
    scan_base.alloc(b->n_scans);
    scan_base.upload(hb.data(), hb.size(), st);
    LVI_CUDA(cudaStreamSynchronize(st));   // hb goes out of scope
  }
  m->n_points = n_cloud;
  DBuf<int> mm(8);
  m->grid_d.alloc(1);
  if (opt.forced_grid_d) {
    LVI_CUDA(cudaMemcpyAsync(m->grid_d.p, opt.forced_grid_d, sizeof(GridParams), cudaMemcpyDeviceToDevice, st));
  } else {
    LVI_LAUNCH(ctx, voxel_reduce_minmax_kernel, 1, 256, 0, b->mm.p, scan_base.p, b->n_scans, mm.p);
    LVI_LAUNCH(ctx, voxel_grid_params_kernel, 1, 1, 0, mm.p, leaf, m->grid_d.p);
  }
  // run records: at most one per finite point (+ one per tile cut); 12 B each, twice (sort double buffer)
  const size_t cap = static_cast<size_t>(n) + static_cast<size_t>((n + kRunTile - 1) / kRunTile) + 1;
  DBuf<unsigned long long> rk(cap), rk2(cap);
  DBuf<uint32_t> rl(cap), rl2(cap);
  DBuf<unsigned int> nruns_d(4);
  nruns_d.zero(st);
  const int64_t n_tiles = (n + kRunTile - 1) / kRunTile;
  LVI_LAUNCH(ctx, voxel_runs_kernel, static_cast<int>(std::min<int64_t>(n_tiles, static_cast<int64_t>(ctx->sm_count) * 16)), 256, 0, b->pts.p, n, b->pts_per_scan,
             scan_base.p, m->grid_d.p, rk.p, rl.p, nruns_d.p);
  unsigned int h_runs = 0;
  m->grid_d.download(&m->grid, 1, st);
  LVI_CUDA(cudaMemcpyAsync(&h_runs, nruns_d.p, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
  LVI_CUDA(cudaStreamSynchronize(st));   // host wait 1 of 2: grid status + number of runs (CUB needs the item count on the host)
  LVI_REQUIRE(m->grid.status != 1, LVI_ERR_INVALID, "lvi_voxel_build: no finite point in the cloud");
  LVI_REQUIRE(m->grid.status != 2, LVI_ERR_OVERFLOW, "Leaf size is too small for the input dataset. Integer indices would overflow.");
  const int R = static_cast<int>(h_runs);
  LVI_REQUIRE(R > 0, LVI_ERR_INVALID, "lvi_voxel_build: no finite point in the cloud");
  DBuf<int32_t> flag(R), leaf_of_run(R), run_off(R);
  size_t tb_sort = 0, tb_scan = 0, tb_scan2 = 0;
  const int end_bit = 32 + m->grid.key_bits;
  cub::DeviceRadixSort::SortPairs(nullptr, tb_sort, rk.p, rk2.p, rl.p, rl2.p, R, 0, end_bit, st);
  cub::DeviceScan::InclusiveSum(nullptr, tb_scan, flag.p, leaf_of_run.p, R, st);
  cub::DeviceScan::ExclusiveSum(nullptr, tb_scan2, reinterpret_cast<const int32_t*>(rl2.p), run_off.p, R, st);
  DBuf<char> tmp(std::max(tb_sort, std::max(tb_scan, tb_scan2)) + 16);
  size_t tb = tmp.n;
  LVI_TIMED(ctx, "cub_radix_sort_runs", cub::DeviceRadixSort::SortPairs(tmp.p, tb, rk.p, rk2.p, rl.p, rl2.p, R, 0, end_bit, st));
  ctx->launches += (end_bit + 7) / 8 + 1;
  LVI_LAUNCH(ctx, voxel_run_flag_kernel, (R + 255) / 256, 256, 0, rk2.p, R, flag.p);
  tb = tmp.n;
  LVI_TIMED(ctx, "cub_scan_runs", cub::DeviceScan::InclusiveSum(tmp.p, tb, flag.p, leaf_of_run.p, R, st));
  tb = tmp.n;
  LVI_TIMED(ctx, "cub_scan_runs", cub::DeviceScan::ExclusiveSum(tmp.p, tb, reinterpret_cast<const int32_t*>(rl2.p), run_off.p, R, st));
  ctx->launches += 4;
  // leaves <= runs: the CSR arrays are sized by the run count and trimmed by n_leaves
  DBuf<int32_t> leaf_run0(static_cast<size_t>(R) + 1);
  m->leaf_key.alloc(static_cast<size_t>(R)); m->leaf_start.alloc(static_cast<size_t>(R) + 1);
  DBuf<int> totals(2);
  LVI_LAUNCH(ctx, voxel_leaf_bounds_kernel, (R + 255) / 256, 256, 0, rk2.p, rl2.p, flag.p, leaf_of_run.p, run_off.p, R, m->leaf_key.p, m->leaf_start.p,
             leaf_run0.p, totals.p);
  int h_tot[2] = {0, 0};
  totals.download(h_tot, 2, st);
  LVI_CUDA(cudaStreamSynchronize(st));   // host wait 2 of 2: number of leaves / binned points (sizes of everything that follows)
  const int L = h_tot[0];
  m->n_leaves = L; m->n_binned = h_tot[1];
  const int n_out = h_tot[1];
  const int n_gt = (n_out + kGatherTile - 1) / kGatherTile;
  DBuf<int32_t> tile_run0(static_cast<size_t>(n_gt) + 1);
  DBuf<double> partial((static_cast<size_t>(L) + n_gt + 1) * kRunSums);
  m->pts_sorted.alloc(static_cast<size_t>(std::max(n_out, 1)));
  LVI_LAUNCH(ctx, voxel_tile_run0_kernel, (R + 255) / 256, 256, 0, run_off.p, rl2.p, R, tile_run0.p);
  LVI_LAUNCH(ctx, voxel_gather_kernel, n_gt, kGatherThreads, 0, b->pts.p, b->pts_per_scan, scan_base.p, rk2.p, run_off.p, leaf_of_run.p, tile_run0.p,
             m->leaf_start.p, R, n_out, m->pts_sorted.p, partial.p, opt.idx_from_w ? 1 : 0);
  m->leaf_npts.alloc(L); m->leaf_mean.alloc(3 * L); m->leaf_cov.alloc(9 * L); m->leaf_evals.alloc(3 * L); m->leaf_evecs.alloc(9 * L); m->leaf_icov.alloc(9 * L);
  if (L) {
    DBuf<double> leaf_sums(static_cast<size_t>(L) * kRunSums);
    LVI_LAUNCH(ctx, voxel_leaf_sums_kernel, static_cast<int>((static_cast<int64_t>(L) * 8 + 255) / 256), 256, 0, partial.p, m->leaf_start.p, L, leaf_sums.p);
    LVI_LAUNCH(ctx, voxel_leaf_stats_kernel, (L + 127) / 128, 128, 0, leaf_sums.p, m->leaf_start.p, L, min_points, eig_mult, m->leaf_npts.p, m->leaf_mean.p,
               m->leaf_cov.p, m->leaf_evals.p, m->leaf_evecs.p, m->leaf_icov.p);
    if (m->grid.ncell <= (1LL << 26)) {
      m->cell2leaf.alloc(m->grid.ncell);
      LVI_CUDA(cudaMemsetAsync(m->cell2leaf.p, 0xff, sizeof(int32_t) * m->grid.ncell, st));
      LVI_LAUNCH(ctx, voxel_cell2leaf_kernel, (L + 255) / 256, 256, 0, m->leaf_key.p, L, m->cell2leaf.p);
    }
    LVI_CUDA(cudaStreamSynchronize(st));  // leaf_sums is freed on return
  }
  return m.release();
}

void voxel_local_minmax(lvi_ctx* ctx, const lvi_scan_batch* b, const int* scan_base_d, int* mm_d) {
  LVI_LAUNCH(ctx, voxel_reduce_minmax_kernel, 1, 256, 0, b->mm.p, scan_base_d, b->n_scans, mm_d);
}
void voxel_grid_from_minmax(lvi_ctx* ctx, const int* mm_d, float leaf, GridParams* grid_d) {
  LVI_LAUNCH(ctx, voxel_grid_params_kernel, 1, 1, 0, mm_d, leaf, grid_d);
}

static lvi_voxel_map* build_from_device(lvi_ctx* ctx, const void* xyz_d, size_t stride, int64_t n, float leaf, int min_points, double eig_mult) {
  // a PCL cloud from the ABI: packed once (one streaming pass that also yields the min/max), then the batch path
  std::unique_ptr<lvi_scan_batch> b(batch_import_xyzi(ctx, xyz_d, stride, n, 65536));
  return voxel_build_from_batch(ctx, b.get(), nullptr, leaf, min_points, eig_mult);
}

}  // namespace lvi

using namespace lvi;

extern "C" {

int lvi_voxel_build_d(lvi_ctx* ctx, const void* xyz_d, size_t stride_bytes, int64_t n_points, float leaf_size, int min_points, double eig_mult,
                      lvi_voxel_map** out) {
  return guarded([&] {
    LVI_REQUIRE(ctx && xyz_d && out, LVI_ERR_INVALID, "lvi_voxel_build_d: null argument");
    activate(ctx);
    *out = build_from_device(ctx, xyz_d, stride_bytes, n_points, leaf_size, min_points, eig_mult);
  });
}

int lvi_voxel_build_batch(lvi_ctx* ctx, const lvi_scan_batch* batch, const uint8_t* scan_keep, float leaf_size, int min_points, double eig_mult,
                          lvi_voxel_map** out) {
  return guarded([&] {
    LVI_REQUIRE(ctx && batch && out, LVI_ERR_INVALID, "lvi_voxel_build_batch: null argument");
    activate(ctx);
    *out = voxel_build_from_batch(ctx, batch, scan_keep, leaf_size, min_points, eig_mult);
  });
}

int lvi_voxel_build(lvi_ctx* ctx, const void* xyz, size_t stride_bytes, int64_t n_points, float leaf_size, int min_points, double eig_mult,
                    lvi_voxel_map** out) {
  return guarded([&] {
    LVI_REQUIRE(ctx && xyz && out, LVI_ERR_INVALID, "lvi_voxel_build: null argument");
    LVI_REQUIRE(n_points > 0, LVI_ERR_INVALID, "lvi_voxel_build: empty cloud");
    activate(ctx);
    DBuf<char> in(static_cast<size_t>(n_points) * stride_bytes);
    LVI_CUDA(cudaMemcpyAsync(in.p, xyz, in.n, cudaMemcpyHostToDevice, ctx->stream));
    *out = build_from_device(ctx, in.p, stride_bytes, n_points, leaf_size, min_points, eig_mult);
  });
}

int lvi_voxel_destroy(lvi_voxel_map* m) {
  delete m;
  return LVI_OK;
}
int64_t lvi_voxel_num_leaves(const lvi_voxel_map* m) { return m ? m->n_leaves : 0; }
int64_t lvi_voxel_num_points(const lvi_voxel_map* m) { return m ? m->n_binned : 0; }
int lvi_voxel_grid(const lvi_voxel_map* m, int32_t min_b[3], int32_t div_b[3]) {
  if (!m) return LVI_ERR_INVALID;
  for (int k = 0; k < 3; ++k) { min_b[k] = m->grid.min_b[k]; div_b[k] = m->grid.div_b[k]; }
  return LVI_OK;
}

int lvi_voxel_export(lvi_ctx* ctx, const lvi_voxel_map* m, int64_t* keys, int32_t* nr_points, double* mean, double* cov, double* evals,
                     double* evecs, double* icov, int64_t* leaf_start, int32_t* point_index) {
  return guarded([&] {
    LVI_REQUIRE(ctx && m, LVI_ERR_INVALID, "lvi_voxel_export: null argument");
    LVI_REQUIRE(!m->lookup_only, LVI_ERR_INVALID, "lvi_voxel_export: the look-up map of a sharded build holds no leaf statistics");
    activate(ctx);
    cudaStream_t st = ctx->stream;
    const size_t L = static_cast<size_t>(m->n_leaves);
    std::vector<int32_t> k32, ls32;
    std::vector<float4> pts;
    if (keys) { k32.resize(L); m->leaf_key.download(k32.data(), L, st); }
    if (nr_points) m->leaf_npts.download(nr_points, L, st);
    if (mean) m->leaf_mean.download(mean, 3 * L, st);
    if (cov) m->leaf_cov.download(cov, 9 * L, st);
    if (evals) m->leaf_evals.download(evals, 3 * L, st);
    if (evecs) m->leaf_evecs.download(evecs, 9 * L, st);
    if (icov) m->leaf_icov.download(icov, 9 * L, st);
    if (leaf_start) { ls32.resize(L + 1); m->leaf_start.download(ls32.data(), L + 1, st); }
    if (point_index) { pts.resize(m->n_binned); m->pts_sorted.download(pts.data(), m->n_binned, st); }
    LVI_CUDA(cudaStreamSynchronize(st));
    if (keys) for (size_t i = 0; i < L; ++i) keys[i] = k32[i];
    if (leaf_start) for (size_t i = 0; i <= L; ++i) leaf_start[i] = ls32[i];
    if (point_index) for (int64_t i = 0; i < m->n_binned; ++i) { int v; memcpy(&v, &pts[i].w, 4); point_index[i] = v; }
  });
}

}  // extern "C"
