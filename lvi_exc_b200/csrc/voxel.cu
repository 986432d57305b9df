// voxel.cu — NDT voxel-grid covariance / eigen build on sm_100a.
//
// Replaces pclomp::VoxelGridCovariance::applyFilter (N/voxel_grid_covariance_omp_impl.hpp:49-374), which is serial,
// inserts every point into a std::map and copies it into Leaf::pointList_.  Here:
//   1. compact_minmax : one streaming pass over the AoS cloud (coalesced 16 B loads) -> float4 copy + float min/max
//   2. grid_params    : min_b_/div_b_/divb_mul_ in the reference's FLOAT arithmetic (:87-103, Q2), overflow guard (:75-84)
//   3. keys           : per point linear voxel index (bit-exact with the reference), non-finite -> sentinel
//   4. stable radix sort by key (CUB, only the significant key bits) + run-length encode -> CSR leaves in std::map order
//   5. reorder        : points gathered into leaf order (cloud order inside a leaf = pointList_)
//   6. leaf_stats     : one warp per leaf, fp64 sums reduced with warp shuffles, then mean/cov (Q1 identity seed, :333-334),
//                       Jacobi eigen-decomposition, eigenvalue inflation (:349-360) and inverse covariance (:363-368)
// HBM-bound; algorithmic bytes: 12 B read per point + 200 B written per leaf.
// Compiled with -fmad=false: index arithmetic must match the CPU bit for bit.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstring>
#include <memory>

#include "eig3.cuh"
#include "map.cuh"

namespace lvi {

__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// ---- 1. compact + min/max ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) voxel_compact_minmax_kernel(const char* __restrict__ in, size_t stride, int64_t n,
                                                                   float4* __restrict__ out, int* __restrict__ mm /*[6] ordered*/) {
  float mn0 = 3.402823466e38f, mn1 = mn0, mn2 = mn0, mx0 = -mn0, mx1 = -mn0, mx2 = -mn0;
  const bool vec = (stride % 16 == 0) && ((reinterpret_cast<size_t>(in) & 15) == 0);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float x, y, z;
    if (vec) { const float4 v = __ldg(reinterpret_cast<const float4*>(in + i * stride)); x = v.x; y = v.y; z = v.z; }
    else { const float* p = reinterpret_cast<const float*>(in + i * stride); x = p[0]; y = p[1]; z = p[2]; }
    out[i] = make_float4(x, y, z, __int_as_float(static_cast<int>(i)));
    if (isfinite(x) && isfinite(y) && isfinite(z)) {  // pcl::getMinMax3D on a non-dense cloud (impl.hpp:72)
      mn0 = fminf(mn0, x); mn1 = fminf(mn1, y); mn2 = fminf(mn2, z);
      mx0 = fmaxf(mx0, x); mx1 = fmaxf(mx1, y); mx2 = fmaxf(mx2, z);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn0 = fminf(mn0, __shfl_xor_sync(0xffffffffu, mn0, o)); mn1 = fminf(mn1, __shfl_xor_sync(0xffffffffu, mn1, o));
    mn2 = fminf(mn2, __shfl_xor_sync(0xffffffffu, mn2, o)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, o));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, o)); mx2 = fmaxf(mx2, __shfl_xor_sync(0xffffffffu, mx2, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(mm + 0, f2ord(mn0)); atomicMin(mm + 1, f2ord(mn1)); atomicMin(mm + 2, f2ord(mn2));
    atomicMax(mm + 3, f2ord(mx0)); atomicMax(mm + 4, f2ord(mx1)); atomicMax(mm + 5, f2ord(mx2));
  }
}

__global__ void voxel_minmax_init_kernel(int* mm) {
  if (threadIdx.x < 3) mm[threadIdx.x] = f2ord(3.402823466e38f);
  else if (threadIdx.x < 6) mm[threadIdx.x] = f2ord(-3.402823466e38f);
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

// ---- 3. keys --------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool voxel_index(const GridParams& g, float x, float y, float z, int& i0, int& i1, int& i2) {
  // impl.hpp:220-222: static_cast<int>(floor(x * inverse_leaf_size_[0]) - static_cast<float>(min_b_[0]))
  i0 = static_cast<int>(floorf(x * g.inv_leaf) - static_cast<float>(g.min_b[0]));
  i1 = static_cast<int>(floorf(y * g.inv_leaf) - static_cast<float>(g.min_b[1]));
  i2 = static_cast<int>(floorf(z * g.inv_leaf) - static_cast<float>(g.min_b[2]));
  return true;
}

__global__ void __launch_bounds__(256) voxel_key_kernel(const float4* __restrict__ pts, int64_t n, const GridParams* __restrict__ gp,
                                                        uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  __shared__ GridParams g;
  if (threadIdx.x == 0) g = *gp;
  __syncthreads();
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 v = pts[i];
    uint32_t key = static_cast<uint32_t>(g.ncell);  // sentinel: sorts last
    if (isfinite(v.x) && isfinite(v.y) && isfinite(v.z)) {
      int i0, i1, i2;
      voxel_index(g, v.x, v.y, v.z, i0, i1, i2);
      key = static_cast<uint32_t>(i0 * g.mul[0] + i1 * g.mul[1] + i2 * g.mul[2]);  // :225
    }
    keys[i] = key;
    vals[i] = static_cast<uint32_t>(i);
  }
}

// ---- 5. reorder ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) voxel_reorder_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ order, int64_t n,
                                                            float4* __restrict__ out) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[i] = __ldg(pts + order[i]);
}

__global__ void voxel_leaf_start_fix_kernel(const uint32_t* __restrict__ ukeys, const int* __restrict__ nruns, uint32_t sentinel,
                                            int* __restrict__ n_leaves) {
  const int r = *nruns;
  *n_leaves = (r > 0 && ukeys[r - 1] == sentinel) ? r - 1 : r;
}

__global__ void __launch_bounds__(256) voxel_cell2leaf_kernel(const int32_t* __restrict__ keys, int n_leaves, int32_t* __restrict__ table) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_leaves) table[keys[i]] = i;
}

// ---- 6. per-leaf statistics: one warp per leaf -------------------------------------------------------------------
__global__ void __launch_bounds__(256) voxel_leaf_stats_kernel(const float4* __restrict__ pts, const int32_t* __restrict__ leaf_start,
                                                               int n_leaves, int min_points, double eig_mult, int32_t* __restrict__ npts_out,
                                                               double* __restrict__ mean_out, double* __restrict__ cov_out,
                                                               double* __restrict__ evals_out, double* __restrict__ evecs_out,
                                                               double* __restrict__ icov_out) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int leaf = blockIdx.x * warps_per_block + (threadIdx.x >> 5); leaf < n_leaves; leaf += gridDim.x * warps_per_block) {
    const int beg = leaf_start[leaf], end = leaf_start[leaf + 1];
    // Leaf() seeds cov_ with Identity (N/voxel_grid_covariance_omp.h:97-106, Q1)
    double sx = 0, sy = 0, sz = 0, sxx = lane == 0 ? 1.0 : 0.0, sxy = 0, sxz = 0, syy = sxx, syz = 0, szz = sxx;
    for (int i = beg + lane; i < end; i += 32) {
      const float4 v = __ldg(pts + i);
      const double x = v.x, y = v.y, z = v.z;
      sx += x; sy += y; sz += z;
      sxx += x * x; sxy += x * y; sxz += x * z; syy += y * y; syz += y * z; szz += z * z;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); sz += __shfl_xor_sync(0xffffffffu, sz, o);
      sxx += __shfl_xor_sync(0xffffffffu, sxx, o); sxy += __shfl_xor_sync(0xffffffffu, sxy, o); sxz += __shfl_xor_sync(0xffffffffu, sxz, o);
      syy += __shfl_xor_sync(0xffffffffu, syy, o); syz += __shfl_xor_sync(0xffffffffu, syz, o); szz += __shfl_xor_sync(0xffffffffu, szz, o);
    }
    if (lane != 0) continue;
    int n = end - beg;
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
}

// ------------------------------------------------------------------------------------------------------------------
static lvi_voxel_map* build_from_device(lvi_ctx* ctx, const void* xyz_d, size_t stride, int64_t n, float leaf, int min_points, double eig_mult) {
  LVI_REQUIRE(n > 0 && n < 2147483647LL, LVI_ERR_INVALID, "lvi_voxel_build: n_points must be in (0, 2^31)");
  LVI_REQUIRE(stride >= 12 && stride % 4 == 0, LVI_ERR_INVALID, "lvi_voxel_build: stride must be a multiple of 4, >= 12");
  LVI_REQUIRE(leaf > 0, LVI_ERR_INVALID, "lvi_voxel_build: leaf_size must be positive");
  auto m = std::unique_ptr<lvi_voxel_map>(new lvi_voxel_map());
  m->ctx = ctx; m->n_points = n; m->leaf_size = leaf; m->min_points = min_points; m->eig_mult = eig_mult;
  cudaStream_t st = ctx->stream;
  DBuf<float4> pts(n);
  DBuf<int> mm(8);
  m->grid_d.alloc(1);
  LVI_LAUNCH(ctx, voxel_minmax_init_kernel, 1, 32, 0, mm.p);
  const int grid = grid_for(n, 256, ctx->sm_count, 8);
  LVI_LAUNCH(ctx, voxel_compact_minmax_kernel, grid, 256, 0, static_cast<const char*>(xyz_d), stride, n, pts.p, mm.p);
  LVI_LAUNCH(ctx, voxel_grid_params_kernel, 1, 1, 0, mm.p, leaf, m->grid_d.p);
  m->grid_d.download(&m->grid, 1, st);
  LVI_CUDA(cudaStreamSynchronize(st));
  LVI_REQUIRE(m->grid.status != 1, LVI_ERR_INVALID, "lvi_voxel_build: no finite point in the cloud");
  LVI_REQUIRE(m->grid.status != 2, LVI_ERR_OVERFLOW, "Leaf size is too small for the input dataset. Integer indices would overflow.");
  DBuf<uint32_t> keys(n), vals(n), keys2(n), vals2(n), ukeys(n);
  DBuf<int> counts(n + 1), nruns(2);
  LVI_LAUNCH(ctx, voxel_key_kernel, grid, 256, 0, pts.p, n, m->grid_d.p, keys.p, vals.p);
  size_t tmp_bytes = 0, tb2 = 0, tb3 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.p, keys2.p, vals.p, vals2.p, static_cast<int>(n), 0, m->grid.key_bits, st);
  cub::DeviceRunLengthEncode::Encode(nullptr, tb2, keys2.p, ukeys.p, counts.p, nruns.p, static_cast<int>(n), st);
  cub::DeviceScan::ExclusiveSum(nullptr, tb3, counts.p, counts.p, static_cast<int>(n), st);
  DBuf<char> tmp(std::max(tmp_bytes, std::max(tb2, tb3)) + 16);
  size_t tb = tmp.n;
  LVI_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys.p, keys2.p, vals.p, vals2.p, static_cast<int>(n), 0, m->grid.key_bits, st));
  ctx->launches += (m->grid.key_bits + 7) / 8 + 1;
  tb = tmp.n;
  LVI_CUDA(cub::DeviceRunLengthEncode::Encode(tmp.p, tb, keys2.p, ukeys.p, counts.p, nruns.p, static_cast<int>(n), st));
  ctx->launches += 2;
  LVI_LAUNCH(ctx, voxel_leaf_start_fix_kernel, 1, 1, 0, ukeys.p, nruns.p, static_cast<uint32_t>(m->grid.ncell), nruns.p + 1);
  int h_runs[2] = {0, 0};
  nruns.download(h_runs, 2, st);
  LVI_CUDA(cudaStreamSynchronize(st));
  const int L = h_runs[1];
  m->n_leaves = L;
  // exclusive scan of the run lengths -> CSR offsets (L+1 entries; entry L = number of binned points)
  m->leaf_start.alloc(static_cast<size_t>(L) + 1);
  tb = tmp.n;
  LVI_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, counts.p, m->leaf_start.p, L + 1, st));
  ctx->launches += 2;
  int h_binned = 0;
  LVI_CUDA(cudaMemcpyAsync(&h_binned, m->leaf_start.p + L, sizeof(int), cudaMemcpyDeviceToHost, st));
  LVI_CUDA(cudaStreamSynchronize(st));
  m->n_binned = h_binned;
  m->leaf_key.alloc(L);
  LVI_CUDA(cudaMemcpyAsync(m->leaf_key.p, ukeys.p, sizeof(int32_t) * L, cudaMemcpyDeviceToDevice, st));
  m->pts_sorted.alloc(std::max<int64_t>(m->n_binned, 1));
  if (m->n_binned) LVI_LAUNCH(ctx, voxel_reorder_kernel, grid_for(m->n_binned, 256, ctx->sm_count, 8), 256, 0, pts.p, vals2.p, m->n_binned, m->pts_sorted.p);
  m->leaf_npts.alloc(L); m->leaf_mean.alloc(3 * L); m->leaf_cov.alloc(9 * L); m->leaf_evals.alloc(3 * L); m->leaf_evecs.alloc(9 * L); m->leaf_icov.alloc(9 * L);
  if (L) {
    LVI_LAUNCH(ctx, voxel_leaf_stats_kernel, grid_for(static_cast<int64_t>(L) * 32, 256, ctx->sm_count, 8), 256, 0, m->pts_sorted.p, m->leaf_start.p, L,
               min_points, eig_mult, m->leaf_npts.p, m->leaf_mean.p, m->leaf_cov.p, m->leaf_evals.p, m->leaf_evecs.p, m->leaf_icov.p);
    if (m->grid.ncell <= (1LL << 26)) {
      m->cell2leaf.alloc(m->grid.ncell);
      LVI_CUDA(cudaMemsetAsync(m->cell2leaf.p, 0xff, sizeof(int32_t) * m->grid.ncell, st));
      LVI_LAUNCH(ctx, voxel_cell2leaf_kernel, (L + 255) / 256, 256, 0, m->leaf_key.p, L, m->cell2leaf.p);
    }
  }
  LVI_CUDA(cudaStreamSynchronize(st));  // temporaries are freed on return
  return m.release();
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
