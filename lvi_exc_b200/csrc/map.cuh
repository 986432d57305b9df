// map.cuh — device-resident LiDAR map structures shared by voxel.cu / surfel.cu / assoc.cu.
//
// HBM layout:
//   pts_sorted  float4[n_binned]  x,y,z + original cloud index (as int bits), grouped by leaf, cloud order inside a
//                                 leaf (= Leaf::pointList_, N/voxel_grid_covariance_omp.h:188)
//   leaf_*      SoA per occupied voxel in ascending linear-index order (= std::map iteration order, :198)
//   cell2leaf   optional dense int32 table over the voxel grid (div_b[0]*div_b[1]*div_b[2] cells) for O(1) lookups
#pragma once
#include "common.cuh"

namespace lvi {

struct GridParams {  // filled on the device by voxel_grid_params_kernel
  float mn[3], mx[3];
  int min_b[3], div_b[3], mul[3];
  float inv_leaf;
  int status;      // 0 ok, 1 no finite point, 2 index overflow
  int key_bits;    // bits needed to represent a linear index
  long long ncell;
};

}  // namespace lvi

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
};
