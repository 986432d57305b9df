// shard.cu — the LiDAR map path sharded over the GPUs of a node (SURVEY §8e collectives 1 and 2).
//
// Every rank holds the scans of ITS time chunk (de-skewed into the map frame on that rank).  The reference builds one NDT voxel map over
// the whole cloud (LiDAROdometry::updateKeyScan -> ndt_omp_->setInputTarget, L/src/core/lidar_odometry.cpp:89-104) and associates every scan
// against it (L/src/core/surfel_association.cpp:111-159,240-244).  Sharded:
//   1. grid: the per-rank min / max of the map cloud meet in ncclAllReduce(min) / (max) on the order-preserving int images of the floats,
//      so min_b_ / div_b_ -- and with them every voxel index -- are those of the whole cloud, bit for bit (Q2);
//   2. leaves are OWNED by ranks: a 4096-bin histogram of the voxel indices (ncclAllReduce(sum)) gives balanced index ranges; every rank
//      sorts its points by owner (stable) and an all-to-all (grouped ncclSend / ncclRecv) moves each point, tagged with its index in the
//      whole map cloud, to the owner of its voxel.  Ranks are time chunks and the partition is stable, so an owner receives the points of
//      a leaf in cloud order: its leaves, point lists, RANSAC planes and boxes are those of the single-GPU build (no merging of partial
//      sums, no dependence on N);
//   3. the planes of all ranks are gathered on every rank (ncclBroadcast per owner; owners hold ascending index ranges, so the
//      concatenation is in the reference's plane_id order) together with a look-up map holding just the surfel leaves;
//   4. association: every rank runs the single-GPU kernels on its own scans; the every-10th decimation of the reference runs over ALL
//      associated points in time order, so the ranks exchange their hit counts (ncclAllGather) and keep the hits whose GLOBAL rank is a
//      multiple of the step; the selected points are then gathered on every rank, in time order.
// With world == 1 both entry points reduce to the single-GPU calls.
#include <cub/cub.cuh>

#include <cstdlib>
#include <cstring>

#include "map.cuh"
#include "nccl_dyn.hpp"

namespace lvi {

constexpr int kHistBins = 4096;
constexpr int kPlaneRecDoubles = 14;   // p4[4] Pi[3] bmin[3] bmax[3] | (key, n_inliers) as two int32

static void nccl_ok(ncclResult_t r, const char* what) {
  LVI_REQUIRE(r == ncclSuccess, LVI_ERR_NCCL, std::string(what) + ": " + nccl().GetErrorString(r));
}
static ncclComm_t comm_of(lvi_ctx* ctx) { return static_cast<ncclComm_t>(ctx->nccl); }

__device__ __forceinline__ bool shard_key(const GridParams& g, const float4& v, uint32_t& key) {
  if (!(isfinite(v.x) && isfinite(v.y) && isfinite(v.z))) return false;
  // N/voxel_grid_covariance_omp_impl.hpp:220-225, same float arithmetic as voxel.cu
  const int i0 = static_cast<int>(floorf(v.x * g.inv_leaf) - static_cast<float>(g.min_b[0]));
  const int i1 = static_cast<int>(floorf(v.y * g.inv_leaf) - static_cast<float>(g.min_b[1]));
  const int i2 = static_cast<int>(floorf(v.z * g.inv_leaf) - static_cast<float>(g.min_b[2]));
  key = static_cast<uint32_t>(i0 * g.mul[0] + i1 * g.mul[1] + i2 * g.mul[2]);
  return true;
}

__global__ void __launch_bounds__(256) shard_hist_kernel(const float4* __restrict__ pts, int64_t n, int64_t pts_per_scan, const int* __restrict__ scan_base,
                                                         const GridParams* __restrict__ gp, unsigned long long* __restrict__ hist) {
  __shared__ unsigned int sh[kHistBins];
  __shared__ GridParams g;
  if (threadIdx.x == 0) g = *gp;
  for (int b = threadIdx.x; b < kHistBins; b += 256) sh[b] = 0;
  __syncthreads();
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
    if (scan_base && scan_base[i / pts_per_scan] < 0) continue;
    uint32_t key;
    if (!shard_key(g, pts[i], key)) continue;
    atomicAdd(&sh[static_cast<unsigned long long>(key) * kHistBins / static_cast<unsigned long long>(g.ncell)], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < kHistBins; b += 256) if (sh[b]) atomicAdd(hist + b, static_cast<unsigned long long>(sh[b]));
}

// owner of every point (world = not part of the map) + per-owner counts
__global__ void __launch_bounds__(256) shard_owner_kernel(const float4* __restrict__ pts, int64_t n, int64_t pts_per_scan, const int* __restrict__ scan_base,
                                                          const GridParams* __restrict__ gp, const uint32_t* __restrict__ split, int world,
                                                          unsigned char* __restrict__ owner, uint32_t* __restrict__ src, unsigned int* __restrict__ counts) {
  __shared__ GridParams g;
  __shared__ unsigned int sc[32];
  if (threadIdx.x == 0) g = *gp;
  if (threadIdx.x < 32) sc[threadIdx.x] = 0;
  __syncthreads();
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
    int o = world;
    uint32_t key;
    if ((!scan_base || scan_base[i / pts_per_scan] >= 0) && shard_key(g, pts[i], key)) {
      o = 0;
      while (o + 1 < world && key >= split[o + 1]) ++o;
      atomicAdd(&sc[o], 1u);
    }
    owner[i] = static_cast<unsigned char>(o);
    src[i] = static_cast<uint32_t>(i);
  }
  __syncthreads();
  if (threadIdx.x < world && sc[threadIdx.x]) atomicAdd(counts + threadIdx.x, sc[threadIdx.x]);
}

// send buffer in (owner, cloud order): x, y, z, index of the point in the WHOLE map cloud
__global__ void __launch_bounds__(256) shard_pack_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ order, int64_t n_send, int64_t pts_per_scan,
                                                         const int* __restrict__ scan_base, long long cloud_offset, float4* __restrict__ out) {
  for (int64_t j = blockIdx.x * 256ll + threadIdx.x; j < n_send; j += gridDim.x * 256ll) {
    const uint32_t i = order[j];
    const float4 v = pts[i];
    long long c = i;
    if (scan_base) { const long long s = i / pts_per_scan; c = scan_base[s] + (i - s * pts_per_scan); }
    out[j] = make_float4(v.x, v.y, v.z, __int_as_float(static_cast<int>(cloud_offset + c)));
  }
}

__global__ void shard_plane_pack_kernel(const double* __restrict__ p4, const double* __restrict__ Pi, const double* __restrict__ bmin, const double* __restrict__ bmax,
                                        const int32_t* __restrict__ key, const int32_t* __restrict__ ninl, int P, double* __restrict__ rec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  double* r = rec + static_cast<size_t>(i) * kPlaneRecDoubles;
  for (int k = 0; k < 4; ++k) r[k] = p4[4 * i + k];
  for (int k = 0; k < 3; ++k) { r[4 + k] = Pi[3 * i + k]; r[7 + k] = bmin[3 * i + k]; r[10 + k] = bmax[3 * i + k]; }
  int2 kk = make_int2(key[i], ninl[i]);
  r[13] = *reinterpret_cast<double*>(&kk);
}
__global__ void shard_plane_unpack_kernel(const double* __restrict__ rec, int P, double* __restrict__ p4, double* __restrict__ Pi, double* __restrict__ bmin,
                                          double* __restrict__ bmax, int32_t* __restrict__ key, int32_t* __restrict__ ninl, int32_t* __restrict__ leaf,
                                          int32_t* __restrict__ leaf2plane, int32_t* __restrict__ leaf_key) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const double* r = rec + static_cast<size_t>(i) * kPlaneRecDoubles;
  for (int k = 0; k < 4; ++k) p4[4 * i + k] = r[k];
  for (int k = 0; k < 3; ++k) { Pi[3 * i + k] = r[4 + k]; bmin[3 * i + k] = r[7 + k]; bmax[3 * i + k] = r[10 + k]; }
  const int2 kk = *reinterpret_cast<const int2*>(r + 13);
  key[i] = kk.x; ninl[i] = kk.y; leaf[i] = i; leaf2plane[i] = i; leaf_key[i] = kk.x;
}
__global__ void shard_cell2leaf_kernel(const int32_t* __restrict__ keys, int n, int32_t* __restrict__ table) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) table[keys[i]] = i;
}

// every `step`-th of the associated points counted over all ranks: local hit i has global rank offset + i
__global__ void shard_decimate_kernel(const lvi_surfel_point* __restrict__ all, long long n_local, long long offset, int step, lvi_surfel_point* __restrict__ out) {
  const long long first = (step - offset % step) % step;   // first local i with (offset + i) % step == 0
  for (long long j = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; first + j * step < n_local; j += gridDim.x * static_cast<long long>(blockDim.x))
    out[j] = all[first + j * step];
}

template <class T>
static std::vector<T> all_gather_host(lvi_ctx* ctx, const T& mine) {   // one small value per rank, returned on the host
  const int W = ctx->world;
  std::vector<T> all(W);
  if (W == 1) { all[0] = mine; return all; }
  DBuf<T> s(1), r(W);
  LVI_CUDA(cudaMemcpyAsync(s.p, &mine, sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  nccl_ok(nccl().AllGather(s.p, r.p, sizeof(T), ncclChar, comm_of(ctx), ctx->stream), "ncclAllGather");
  LVI_CUDA(cudaMemcpyAsync(all.data(), r.p, sizeof(T) * W, cudaMemcpyDeviceToHost, ctx->stream));
  LVI_CUDA(cudaStreamSynchronize(ctx->stream));
  return all;
}

// variable-length gather on every rank: rank q contributes cnt[q] records of `rec_bytes`; buf holds all of them at off[q] (own part in place)
static void all_gather_v(lvi_ctx* ctx, char* buf, const std::vector<long long>& cnt, const std::vector<long long>& off, size_t rec_bytes) {
  if (ctx->world == 1) return;
  nccl_ok(nccl().GroupStart(), "ncclGroupStart");
  for (int q = 0; q < ctx->world; ++q) {
    if (cnt[q] == 0) continue;
    char* at = buf + static_cast<size_t>(off[q]) * rec_bytes;
    nccl_ok(nccl().Broadcast(at, at, static_cast<size_t>(cnt[q]) * rec_bytes, ncclChar, q, comm_of(ctx), ctx->stream), "ncclBroadcast");
  }
  nccl_ok(nccl().GroupEnd(), "ncclGroupEnd");
}

}  // namespace lvi

using namespace lvi;

extern "C" {

int lvi_map_build_sharded(lvi_ctx* ctx, const lvi_scan_batch* local, const uint8_t* scan_keep, float leaf_size, int min_points, double eig_mult, double lambda,
                          int min_leaf_points, float ransac_threshold, int min_inliers, lvi_voxel_map** out_map, lvi_surfel_set** out_surfels, int64_t* stats) {
  return guarded([&] {
    LVI_REQUIRE(ctx && local && out_map && out_surfels, LVI_ERR_INVALID, "lvi_map_build_sharded: null argument");
    LVI_REQUIRE(leaf_size > 0, LVI_ERR_INVALID, "lvi_map_build_sharded: leaf_size must be positive");
    activate(ctx);
    const int W = ctx->world, R = ctx->rank;
    if (W == 1) {
      lvi_voxel_map* m = voxel_build_from_batch(ctx, local, scan_keep, leaf_size, min_points, eig_mult);
      lvi_surfel_set* s = nullptr;
      const int rc = lvi_surfel_extract(ctx, m, lambda, min_leaf_points, ransac_threshold, min_inliers, &s);
      if (rc != LVI_OK) { lvi_voxel_destroy(m); throw Error(rc, lvi_last_error()); }
      if (stats) { stats[0] = 0; stats[1] = m->n_binned; stats[2] = m->n_leaves; stats[3] = s->n_planes; }
      *out_map = m; *out_surfels = s;
      return;
    }
    LVI_REQUIRE(W <= 32, LVI_ERR_INVALID, "lvi_map_build_sharded: more than 32 ranks");
    cudaStream_t st = ctx->stream;
    const int64_t n = local->n;
    LVI_REQUIRE(n > 0 && n < 2147483647LL, LVI_ERR_INVALID, "lvi_map_build_sharded: local points must be in (0, 2^31)");
    // the part of the local batch that belongs to the map cloud (key scans) and its offset in the whole cloud
    DBuf<int> scan_base;
    long long n_cloud = n;
    if (scan_keep) {
      std::vector<int> hb(local->n_scans);
      long long acc = 0;
      for (int s = 0; s < local->n_scans; ++s) {
        const long long cnt = std::min<long long>(local->pts_per_scan, n - static_cast<long long>(s) * local->pts_per_scan);
        if (scan_keep[s]) { hb[s] = static_cast<int>(acc); acc += cnt; } else hb[s] = -1;
      }
      n_cloud = acc;
      scan_base.alloc(local->n_scans);
      scan_base.upload(hb.data(), hb.size(), st);
      LVI_CUDA(cudaStreamSynchronize(st));
    }
    const std::vector<long long> clouds = all_gather_host(ctx, n_cloud);
    long long cloud_offset = 0, cloud_total = 0;
    for (int q = 0; q < W; ++q) { if (q < R) cloud_offset += clouds[q]; cloud_total += clouds[q]; }
    LVI_REQUIRE(cloud_total > 0 && cloud_total < 2147483647LL, LVI_ERR_INVALID, "lvi_map_build_sharded: the whole map cloud must have (0, 2^31) points");
    // 1. grid of the WHOLE cloud
    DBuf<int> mm(8);
    DBuf<GridParams> grid_d(1);
    voxel_local_minmax(ctx, local, scan_base.p, mm.p);
    nccl_ok(nccl().AllReduce(mm.p, mm.p, 3, ncclInt, ncclMin, comm_of(ctx), st), "ncclAllReduce(min)");
    nccl_ok(nccl().AllReduce(mm.p + 3, mm.p + 3, 3, ncclInt, ncclMax, comm_of(ctx), st), "ncclAllReduce(max)");
    voxel_grid_from_minmax(ctx, mm.p, leaf_size, grid_d.p);
    GridParams grid;
    grid_d.download(&grid, 1, st);
    LVI_CUDA(cudaStreamSynchronize(st));
    LVI_REQUIRE(grid.status != 1, LVI_ERR_INVALID, "lvi_map_build_sharded: no finite point in the cloud");
    LVI_REQUIRE(grid.status != 2, LVI_ERR_OVERFLOW, "Leaf size is too small for the input dataset. Integer indices would overflow.");
    // 2. balanced ownership ranges of the voxel index
    DBuf<unsigned long long> hist(kHistBins);
    hist.zero(st);
    const int grid_pts = static_cast<int>(std::min<int64_t>((n + 255) / 256, static_cast<int64_t>(ctx->sm_count) * 16));
    LVI_LAUNCH(ctx, shard_hist_kernel, grid_pts, 256, 0, local->pts.p, n, local->pts_per_scan, scan_base.p, grid_d.p, hist.p);
    nccl_ok(nccl().AllReduce(hist.p, hist.p, kHistBins, ncclUint64, ncclSum, comm_of(ctx), st), "ncclAllReduce(hist)");
    std::vector<unsigned long long> h(kHistBins);
    hist.download(h.data(), kHistBins, st);
    LVI_CUDA(cudaStreamSynchronize(st));
    unsigned long long total = 0;
    for (auto v : h) total += v;
    std::vector<uint32_t> split(W + 1, 0);
    {
      unsigned long long acc = 0;
      int q = 1;
      for (int b = 0; b < kHistBins && q < W; ++b) {
        acc += h[b];
        while (q < W && acc * W >= total * static_cast<unsigned long long>(q)) {
          // first voxel index of bin b + 1: the smallest key with key * bins / ncell >= b + 1
          const unsigned long long k = (static_cast<unsigned long long>(b + 1) * static_cast<unsigned long long>(grid.ncell) + kHistBins - 1) / kHistBins;
          split[q++] = static_cast<uint32_t>(std::min<unsigned long long>(k, static_cast<unsigned long long>(grid.ncell)));
        }
      }
      for (; q < W; ++q) split[q] = static_cast<uint32_t>(grid.ncell);
      split[W] = static_cast<uint32_t>(grid.ncell);
    }
    DBuf<uint32_t> split_d(W + 1);
    split_d.upload(split.data(), split.size(), st);
    // 3. stable partition by owner, exchange
    DBuf<unsigned char> owner(n), owner2(n);
    DBuf<uint32_t> src(n), order(n);
    DBuf<unsigned int> counts_d(32);
    counts_d.zero(st);
    LVI_LAUNCH(ctx, shard_owner_kernel, grid_pts, 256, 0, local->pts.p, n, local->pts_per_scan, scan_base.p, grid_d.p, split_d.p, W, owner.p, src.p, counts_d.p);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, owner.p, owner2.p, src.p, order.p, static_cast<int>(n), 0, 6, st);
    DBuf<char> tmp(tb + 16);
    LVI_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, owner.p, owner2.p, src.p, order.p, static_cast<int>(n), 0, 6, st));
    struct Counts { unsigned int c[32]; } mine{};
    LVI_CUDA(cudaMemcpyAsync(mine.c, counts_d.p, sizeof(unsigned int) * 32, cudaMemcpyDeviceToHost, st));
    LVI_CUDA(cudaStreamSynchronize(st));
    const std::vector<Counts> all = all_gather_host(ctx, mine);
    std::vector<long long> s_off(W + 1, 0), r_off(W + 1, 0);
    for (int q = 0; q < W; ++q) { s_off[q + 1] = s_off[q] + mine.c[q]; r_off[q + 1] = r_off[q] + all[q].c[R]; }
    const long long n_send = s_off[W], n_recv = r_off[W];
    LVI_REQUIRE(n_recv < 2147483647LL, LVI_ERR_INVALID, "lvi_map_build_sharded: a rank would own 2^31 points or more");
    DBuf<float4> sendb(std::max<long long>(n_send, 1));
    std::unique_ptr<lvi_scan_batch> rb(new lvi_scan_batch());
    rb->ctx = ctx; rb->n_scans = 1; rb->pts_per_scan = std::max<long long>(n_recv, 1); rb->n = n_recv;
    rb->pts.alloc(std::max<long long>(n_recv, 1));
    if (n_send) LVI_LAUNCH(ctx, shard_pack_kernel, static_cast<int>(std::min<int64_t>((n_send + 255) / 256, static_cast<int64_t>(ctx->sm_count) * 16)), 256, 0, local->pts.p,
                           order.p, n_send, local->pts_per_scan, scan_base.p, cloud_offset, sendb.p);
    nccl_ok(nccl().GroupStart(), "ncclGroupStart");
    for (int q = 0; q < W; ++q) {
      if (mine.c[q]) nccl_ok(nccl().Send(sendb.p + s_off[q], static_cast<size_t>(mine.c[q]) * sizeof(float4), ncclChar, q, comm_of(ctx), st), "ncclSend");
      if (all[q].c[R]) nccl_ok(nccl().Recv(rb->pts.p + r_off[q], static_cast<size_t>(all[q].c[R]) * sizeof(float4), ncclChar, q, comm_of(ctx), st), "ncclRecv");
    }
    nccl_ok(nccl().GroupEnd(), "ncclGroupEnd");
    // 4. this rank's leaves and planes
    std::unique_ptr<lvi_voxel_map> ml;
    lvi_surfel_set* sl = nullptr;
    long long P_local = 0;
    if (n_recv > 0) {
      VoxelBuildOptions opt;
      opt.forced_grid_d = grid_d.p; opt.idx_from_w = true;
      ml.reset(voxel_build_core(ctx, rb.get(), nullptr, leaf_size, min_points, eig_mult, opt));
      const int rc = lvi_surfel_extract(ctx, ml.get(), lambda, min_leaf_points, ransac_threshold, min_inliers, &sl);
      if (rc != LVI_OK) throw Error(rc, lvi_last_error());
      P_local = sl->n_planes;
    }
    std::unique_ptr<lvi_surfel_set> sl_owner(sl);
    // 5. all planes on every rank, in ascending voxel-index order (= the reference's plane_id order)
    const std::vector<long long> Ps = all_gather_host(ctx, P_local);
    std::vector<long long> p_off(W + 1, 0);
    for (int q = 0; q < W; ++q) p_off[q + 1] = p_off[q] + Ps[q];
    const long long P = p_off[W];
    DBuf<double> rec(std::max<long long>(P, 1) * kPlaneRecDoubles);
    if (P_local) LVI_LAUNCH(ctx, shard_plane_pack_kernel, static_cast<int>((P_local + 255) / 256), 256, 0, sl->p4.p, sl->Pi.p, sl->bmin.p, sl->bmax.p, sl->key.p, sl->ninl.p,
                            static_cast<int>(P_local), rec.p + p_off[R] * kPlaneRecDoubles);
    all_gather_v(ctx, reinterpret_cast<char*>(rec.p), Ps, p_off, sizeof(double) * kPlaneRecDoubles);
    auto S = std::unique_ptr<lvi_surfel_set>(new lvi_surfel_set());
    auto M = std::unique_ptr<lvi_voxel_map>(new lvi_voxel_map());
    S->ctx = ctx; S->n_planes = P;
    M->ctx = ctx; M->grid = grid; M->leaf_size = leaf_size; M->min_points = min_points; M->eig_mult = eig_mult; M->n_points = cloud_total; M->n_binned = static_cast<int64_t>(total);
    M->n_leaves = P; M->lookup_only = true;
    M->grid_d.alloc(1);
    LVI_CUDA(cudaMemcpyAsync(M->grid_d.p, grid_d.p, sizeof(GridParams), cudaMemcpyDeviceToDevice, st));
    const long long Pa = std::max<long long>(P, 1);
    S->p4.alloc(4 * Pa); S->Pi.alloc(3 * Pa); S->bmin.alloc(3 * Pa); S->bmax.alloc(3 * Pa); S->leaf.alloc(Pa); S->ninl.alloc(Pa); S->key.alloc(Pa); S->leaf2plane.alloc(Pa);
    M->leaf_key.alloc(Pa);
    if (P) {
      LVI_LAUNCH(ctx, shard_plane_unpack_kernel, static_cast<int>((P + 255) / 256), 256, 0, rec.p, static_cast<int>(P), S->p4.p, S->Pi.p, S->bmin.p, S->bmax.p, S->key.p,
                 S->ninl.p, S->leaf.p, S->leaf2plane.p, M->leaf_key.p);
      if (grid.ncell <= (1LL << 26)) {
        M->cell2leaf.alloc(grid.ncell);
        LVI_CUDA(cudaMemsetAsync(M->cell2leaf.p, 0xff, sizeof(int32_t) * grid.ncell, st));
        LVI_LAUNCH(ctx, shard_cell2leaf_kernel, static_cast<int>((P + 255) / 256), 256, 0, M->leaf_key.p, static_cast<int>(P), M->cell2leaf.p);
      }
    }
    LVI_CUDA(cudaStreamSynchronize(st));   // the staging buffers go out of scope
    if (stats) { stats[0] = n_send - mine.c[R]; stats[1] = n_recv; stats[2] = ml ? ml->n_leaves : 0; stats[3] = P_local; }
    *out_map = M.release(); *out_surfels = S.release();
  });
}

int lvi_associate_sharded(lvi_ctx* ctx, const lvi_voxel_map* m, const lvi_surfel_set* s, const lvi_scan_batch* local, const lvi_point_xyzit* scans_raw_d,
                          int32_t W, int32_t H, double radius, int32_t k_per_ring, int32_t time_step, lvi_surfel_point* out_d, int64_t cap, int64_t* n_out,
                          int64_t* n_all) {
  return guarded([&] {
    LVI_REQUIRE(ctx && m && s && local && scans_raw_d && n_out, LVI_ERR_INVALID, "lvi_associate_sharded: null argument");
    LVI_REQUIRE(time_step >= 1, LVI_ERR_INVALID, "lvi_associate_sharded: bad time_step");
    activate(ctx);
    if (ctx->world == 1) {
      const int rc = lvi_associate_batch(ctx, m, s, local, scans_raw_d, W, H, radius, k_per_ring, time_step, out_d, cap, n_out, n_all);
      if (rc != LVI_OK) throw Error(rc, lvi_last_error());
      return;
    }
    cudaStream_t st = ctx->stream;
    // every associated point of the local scans, in the reference's emission order
    int64_t cap_all = local->n / 8 + 4096, got = 0, tot = 0;
    DBuf<lvi_surfel_point> hits(cap_all);
    int rc = lvi_associate_batch(ctx, m, s, local, scans_raw_d, W, H, radius, k_per_ring, 1, hits.p, cap_all, &got, &tot);
    if (rc == LVI_OK && tot > cap_all) {
      cap_all = tot;
      hits.alloc(cap_all);
      rc = lvi_associate_batch(ctx, m, s, local, scans_raw_d, W, H, radius, k_per_ring, 1, hits.p, cap_all, &got, &tot);
    }
    if (rc != LVI_OK) throw Error(rc, lvi_last_error());
    const std::vector<long long> tots = all_gather_host(ctx, static_cast<long long>(tot));
    long long offset = 0, total = 0;
    for (int q = 0; q < ctx->world; ++q) { if (q < ctx->rank) offset += tots[q]; total += tots[q]; }
    // the points whose GLOBAL rank is a multiple of the step (averageTimeDownSmaple over all associated points, :240-244)
    std::vector<long long> sel(ctx->world), sel_off(ctx->world + 1, 0);
    {
      long long o = 0;
      for (int q = 0; q < ctx->world; ++q) {
        const long long first = (time_step - o % time_step) % time_step;
        sel[q] = tots[q] > first ? (tots[q] - first + time_step - 1) / time_step : 0;
        sel_off[q + 1] = sel_off[q] + sel[q];
        o += tots[q];
      }
    }
    const long long n_sel = sel_off[ctx->world];
    if (n_all) *n_all = total;
    *n_out = n_sel;
    if (!out_d || cap <= 0) return;
    LVI_REQUIRE(cap >= n_sel, LVI_ERR_INVALID, "lvi_associate_sharded: output capacity too small (call with out_d = NULL to size)");
    if (sel[ctx->rank])
      LVI_LAUNCH(ctx, shard_decimate_kernel, static_cast<int>(std::min<long long>((sel[ctx->rank] + 255) / 256, ctx->sm_count * 8)), 256, 0, hits.p, static_cast<long long>(tot),
                 offset, time_step, out_d + sel_off[ctx->rank]);
    all_gather_v(ctx, reinterpret_cast<char*>(out_d), sel, sel_off, sizeof(lvi_surfel_point));
    LVI_CUDA(cudaStreamSynchronize(st));
  });
}

}  // extern "C"
