// map.cuh — device-resident LiDAR map structures shared by voxel.cu / surfel.cu / assoc.cu.
//
// HBM layout:
//   pts_sorted  float4[n_binned]  x,y,z + original cloud index (as int bits), grouped by leaf, cloud order inside a
//                                 leaf (= Leaf::pointList_, N/voxel_grid_covariance_omp.h:188)
//   leaf_*      SoA per occupied voxel in ascending linear-index order (= std::map iteration order, :198)
//   cell2leaf   optional dense int32 table over the voxel grid (div_b[0]*div_b[1]*div_b[2] cells) for O(1) lookups
#pragma once
#include "common.cuh"

typedef struct lvi_scan_batch lvi_scan_batch;
typedef struct lvi_voxel_map lvi_voxel_map;

namespace lvi {

struct GridParams {  // filled on the device by voxel_grid_params_kernel
  float mn[3], mx[3];
  int min_b[3], div_b[3], mul[3];
  float inv_leaf;
  int status;      // 0 ok, 1 no finite point, 2 index overflow
  int key_bits;    // bits needed to represent a linear index
  long long ncell;
};

#ifdef __CUDACC__
// order-preserving int image of a float (atomicMin / atomicMax on floats)
__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// Running min/max of the finite points a thread produces for one scan (pcl::getMinMax3D on a non-dense cloud,
// N/voxel_grid_covariance_omp_impl.hpp:72).  The producers walk the cloud in CTA-contiguous chunks, so a thread stays inside one scan for
// a whole chunk: it accumulates in registers and touches the scan's slots once per chunk -- and only the ones it would move (the plain
// read may be stale, which at worst costs a redundant atomic: the slots move monotonically).
struct ScanMinMax {
  float mn0, mn1, mn2, mx0, mx1, mx2;
  int64_t scan;
  __device__ __forceinline__ void reset(int64_t s) {
    const float big = 3.402823466e38f;
    mn0 = mn1 = mn2 = big; mx0 = mx1 = mx2 = -big; scan = s;
  }
  __device__ __forceinline__ void flush(int* __restrict__ mm) const {
    if (scan < 0 || !(mn0 <= mx0)) return;
    int* m = mm + 6 * scan;
    // the six slots are read together (one round trip to L2, not six dependent ones) before any of them is compared
    const int2 c01 = __ldcg(reinterpret_cast<const int2*>(m)), c23 = __ldcg(reinterpret_cast<const int2*>(m) + 1), c45 = __ldcg(reinterpret_cast<const int2*>(m) + 2);
    int o;
    o = f2ord(mn0); if (o < c01.x) atomicMin(m + 0, o);
    o = f2ord(mn1); if (o < c01.y) atomicMin(m + 1, o);
    o = f2ord(mn2); if (o < c23.x) atomicMin(m + 2, o);
    o = f2ord(mx0); if (o > c23.y) atomicMax(m + 3, o);
    o = f2ord(mx1); if (o > c45.x) atomicMax(m + 4, o);
    o = f2ord(mx2); if (o > c45.y) atomicMax(m + 5, o);
  }
  __device__ __forceinline__ void add(int* __restrict__ mm, int64_t s, float x, float y, float z) {
    if (s != scan) { flush(mm); reset(s); }
    if (isfinite(x) && isfinite(y) && isfinite(z)) {
      mn0 = fminf(mn0, x); mn1 = fminf(mn1, y); mn2 = fminf(mn2, z);
      mx0 = fmaxf(mx0, x); mx1 = fmaxf(mx1, y); mx2 = fmaxf(mx2, z);
    }
  }
};
constexpr int kProducerChunk = 256 * 8;   // points one 256-thread CTA produces between two flushes
#endif

// batch producers shared across translation units (defined in undistort.cu)
lvi_scan_batch* batch_alloc(lvi_ctx* ctx, int32_t n_scans, int64_t pts_per_scan, int64_t n);
lvi_scan_batch* batch_import_xyzi(lvi_ctx* ctx, const void* xyz_d, size_t stride, int64_t n, int64_t pts_per_scan);
// voxel.cu
struct VoxelBuildOptions { const GridParams* forced_grid_d = nullptr; bool idx_from_w = false; };
lvi_voxel_map* voxel_build_core(lvi_ctx* ctx, const lvi_scan_batch* b, const uint8_t* scan_keep, float leaf, int min_points, double eig_mult, const VoxelBuildOptions& opt);
// grid of the cloud selected by scan_base from the per-scan min / max slots of a batch (device): mm[6] ordered ints, then the grid
void voxel_local_minmax(lvi_ctx* ctx, const lvi_scan_batch* b, const int* scan_base_d, int* mm_d);
void voxel_grid_from_minmax(lvi_ctx* ctx, const int* mm_d, float leaf, GridParams* grid_d);
lvi_voxel_map* voxel_build_from_batch(lvi_ctx* ctx, const lvi_scan_batch* b, const uint8_t* scan_keep, float leaf, int min_points, double eig_mult);

}  // namespace lvi

// A batch of organised scans expressed in ONE frame and resident in HBM: what ScanUndistortion keeps as scan_data_in_map_ / map_cloud_
// (L/include/core/scan_undistortion.h:182-188).  The library's own producers (de-skew, pose transform) write it PACKED -- 16 B per
// point, x y z intensity -- together with the float min / max of every scan, so the consumers (voxel build, association) stream 16 B
// per point instead of the 32 B PCL record and the voxel build needs no min/max pass; the 32 B pcl::PointXYZI layout only exists at
// the ABI (lvi_scan_batch_export_xyzi / the *_d entry points that take a PCL cloud).
struct lvi_scan_batch {
  lvi_ctx* ctx = nullptr;
  int32_t n_scans = 0;            // number of min/max slots ("scans"; a flat cloud imported from the ABI is cut into chunks)
  int64_t pts_per_scan = 0;
  int64_t n = 0;                  // points (the last slot may be short for an imported flat cloud)
  lvi::DBuf<float4> pts;          // [n] x, y, z, intensity ; NaN x = no return
  lvi::DBuf<int> mm;              // [n_scans * 6] order-preserving int images of min x,y,z / max x,y,z over the finite points of each scan
};

struct lvi_voxel_map {
  lvi_ctx* ctx = nullptr;
  lvi::GridParams grid{};           // host copy
  lvi::DBuf<lvi::GridParams> grid_d;
  int64_t n_points = 0;             // input points
  int64_t n_binned = 0;             // finite points
  int64_t n_leaves = 0;
  float leaf_size = 0;
  int min_points = 6;
  double eig_mult = 0.01;
  lvi::DBuf<float4> pts_sorted;
  lvi::DBuf<int32_t> leaf_key;      // [L]
  lvi::DBuf<int32_t> leaf_start;    // [L+1]
  lvi::DBuf<int32_t> leaf_npts;     // [L]  nr_points (-1 = rejected)
  lvi::DBuf<double> leaf_mean;      // [L*3]
  lvi::DBuf<double> leaf_cov;       // [L*9]
  lvi::DBuf<double> leaf_evals;     // [L*3]
  lvi::DBuf<double> leaf_evecs;     // [L*9]
  lvi::DBuf<double> leaf_icov;      // [L*9]
  lvi::DBuf<int32_t> cell2leaf;     // dense lookup or empty
  bool lookup_only = false;         // sharded build: keys of the surfel leaves of ALL ranks (leaf l = plane l), no statistics / point lists
};

struct lvi_surfel_set {
  lvi_ctx* ctx = nullptr;
  int64_t n_planes = 0;
  lvi::DBuf<double> p4;        // [P*4]
  lvi::DBuf<double> Pi;        // [P*3]
  lvi::DBuf<double> bmin;      // [P*3]
  lvi::DBuf<double> bmax;      // [P*3]
  lvi::DBuf<int32_t> leaf;     // [P] leaf index
  lvi::DBuf<int32_t> key;      // [P] voxel linear index of the leaf
  lvi::DBuf<int32_t> ninl;     // [P]
  lvi::DBuf<int32_t> leaf2plane;  // [L] plane id or -1
  mutable lvi::DBuf<int32_t> cell2plane;   // dense cell -> plane table (cell2leaf composed with leaf2plane), built by the first association
};
