// assoc.cu — scan-point -> surfel association on sm_100a.
//
// Replaces SurfelAssociation::getAssociation / associateScanToSurfel / averageTimeDownSmaple
// (L/src/core/surfel_association.cpp:111-159,240-244,305-331).  The reference sweeps all W x H points once PER PLANE
// (O(P*W*H), omp over planes).  Every surfel box is the min/max of the points of ONE voxel, so boxes are disjoint and a
// point strictly inside a box lies in that box's voxel (float floor(x*inv) is monotone): the test collapses to an O(1)
// voxel lookup per point with identical results (tests/test_oracle_map.py proves the equivalence on the oracle).
//   1. assoc_hit_kernel    : per point — voxel index (same float arithmetic as the map build) -> plane -> strict bbox test
//                            + |n.x+d| <= radius in fp64.  Streams 12 of every 32 input bytes; HBM-bound.
//   2. assoc_select_kernel : one warp per (scan, ring) — rank of every hit inside its plane's hit list (match_any + a per-warp hash
//                            table of running counts) -> picks hits[step*(s+1)-1], step = max(hits/(k+1),1), if hits >= 2k (:127-136)
//   3. exclusive scan over the reference's emission order (scan, w outer, h inner, timestamp != 0, :139-158), then every
//      `time_step`-th emitted point is gathered into a SurfelPoint record (:240-244).
// Compiled with -fmad=false (bit-exact index selection).
#include <cub/cub.cuh>

#include <cstdlib>

#include "map.cuh"

namespace lvi {

__device__ __forceinline__ double p2plane(double x, double y, double z, const double* __restrict__ p4) {  // point2PlaneDistance :296-303
  double d = x * p4[0];
  d = d + y * p4[1];
  d = d + z * p4[2];
  d = d + p4[3];
  return d > 0 ? d : -d;
}

__device__ __forceinline__ int lookup_leaf(const GridParams& g, const int32_t* __restrict__ cell2leaf, const int32_t* __restrict__ leaf_key,
                                           int n_leaves, float x, float y, float z) {
  if (!(isfinite(x) && isfinite(y) && isfinite(z))) return -1;
  const int i0 = static_cast<int>(floorf(x * g.inv_leaf) - static_cast<float>(g.min_b[0]));
  const int i1 = static_cast<int>(floorf(y * g.inv_leaf) - static_cast<float>(g.min_b[1]));
  const int i2 = static_cast<int>(floorf(z * g.inv_leaf) - static_cast<float>(g.min_b[2]));
  if (i0 < 0 || i1 < 0 || i2 < 0 || i0 >= g.div_b[0] || i1 >= g.div_b[1] || i2 >= g.div_b[2]) return -1;
  const int key = i0 * g.mul[0] + i1 * g.mul[1] + i2 * g.mul[2];
  if (cell2leaf) return cell2leaf[key];
  int lo = 0, hi = n_leaves - 1;  // sparse grids: binary search in the sorted leaf keys
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    const int k = leaf_key[mid];
    if (k == key) return mid;
    if (k < key) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

// Per point: voxel index -> leaf -> plane -> strict bbox test + |n.x + d| <= radius.  stride == 16 with a 16-byte aligned base is the
// library's own packed scan batch: the kernel then moves 16 B in and 4 B out per point.  Every step is a dependent lookup (point ->
// cell2leaf -> leaf2plane -> plane record), so each thread keeps FOUR points in flight and walks the chain stage by stage; the plane
// record is one packed 48-byte row (PlaneRec) instead of 128 B spread over three fp64 arrays.  The box test compares floats (both sides
// are floats widened to double in the reference, L/src/core/surfel_association.cpp:305-331, so the float comparison is the same
// comparison); the plane distance is evaluated in fp64 in the reference's order (:296-303).
struct __align__(16) PlaneRec { float mn[3], a; float mx[3], b; float c, d, pad0, pad1; };

__global__ void __launch_bounds__(256) assoc_plane_rec_kernel(const double* __restrict__ p4, const double* __restrict__ bmin, const double* __restrict__ bmax,
                                                              int n_planes, PlaneRec* __restrict__ rec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_planes) return;
  PlaneRec r;
  for (int k = 0; k < 3; ++k) { r.mn[k] = static_cast<float>(bmin[3 * i + k]); r.mx[k] = static_cast<float>(bmax[3 * i + k]); }   // exact: they ARE floats
  r.a = static_cast<float>(p4[4 * i]); r.b = static_cast<float>(p4[4 * i + 1]); r.c = static_cast<float>(p4[4 * i + 2]); r.d = static_cast<float>(p4[4 * i + 3]);
  r.pad0 = r.pad1 = 0.f;
  rec[i] = r;
}

// cell -> plane in one step: the dense cell2leaf table of the map composed with the surfel set's leaf2plane (built once per surfel set,
// association_tables()), so the per-point chain is point -> cell2plane -> plane record
__global__ void __launch_bounds__(256) assoc_cell2plane_kernel(const int32_t* __restrict__ plane_key, int n_planes, int32_t* __restrict__ table) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_planes) table[plane_key[i]] = i;
}

template <bool DIRECT, int kHitIlp, bool EAGER>
__global__ void __launch_bounds__(256) assoc_hit_kernel(const char* __restrict__ scan_map, size_t stride, int64_t n, const GridParams* __restrict__ gp,
                                                        const int32_t* __restrict__ cell2leaf, const int32_t* __restrict__ leaf_key, int n_leaves,
                                                        const int32_t* __restrict__ leaf2plane, const PlaneRec* __restrict__ planes, double radius,
                                                        int32_t* __restrict__ cand) {
  __shared__ GridParams g;
  if (threadIdx.x == 0) g = *gp;
  __syncthreads();
  const bool vec = (stride % 16 == 0) && ((reinterpret_cast<size_t>(scan_map) & 15) == 0);
  const int64_t span = static_cast<int64_t>(blockDim.x) * kHitIlp;
  const float nanv = __int_as_float(0x7fc00000);
  auto load = [&](int64_t c0, float (&x)[kHitIlp], float (&y)[kHitIlp], float (&z)[kHitIlp]) {
#pragma unroll
    for (int u = 0; u < kHitIlp; ++u) {
      const int64_t i = c0 + threadIdx.x + static_cast<int64_t>(u) * blockDim.x;
      x[u] = y[u] = z[u] = nanv;
      if (i < n) {
        if (vec) { const float4 v = __ldcs(reinterpret_cast<const float4*>(scan_map + i * stride)); x[u] = v.x; y[u] = v.y; z[u] = v.z; }   // streamed: L2 stays with the tables
        else { const float* p = reinterpret_cast<const float*>(scan_map + i * stride); x[u] = p[0]; y[u] = p[1]; z[u] = p[2]; }
      }
    }
  };
  // the points of the NEXT chunk are requested before the dependent look-ups of the current one start, so the HBM stream never waits for
  // the table chain (point -> cell -> plane record)
  float xn[kHitIlp], yn[kHitIlp], zn[kHitIlp];
  const int64_t step = static_cast<int64_t>(gridDim.x) * span;
  if (blockIdx.x * span < n) load(blockIdx.x * span, xn, yn, zn);
  for (int64_t c0 = blockIdx.x * span; c0 < n; c0 += step) {
    float x[kHitIlp], y[kHitIlp], z[kHitIlp];
    int pl[kHitIlp];
#pragma unroll
    for (int u = 0; u < kHitIlp; ++u) { x[u] = xn[u]; y[u] = yn[u]; z[u] = zn[u]; }
    if (c0 + step < n) load(c0 + step, xn, yn, zn);
    if (DIRECT) {   // cell2leaf IS the cell -> plane table
#pragma unroll
      for (int u = 0; u < kHitIlp; ++u) pl[u] = lookup_leaf(g, cell2leaf, nullptr, 0, x[u], y[u], z[u]);
    } else {
      int leaf[kHitIlp];
#pragma unroll
      for (int u = 0; u < kHitIlp; ++u) leaf[u] = lookup_leaf(g, cell2leaf, leaf_key, n_leaves, x[u], y[u], z[u]);
#pragma unroll
      for (int u = 0; u < kHitIlp; ++u) pl[u] = leaf[u] >= 0 ? __ldg(leaf2plane + leaf[u]) : -1;
    }
    float4 r0[kHitIlp], r1[kHitIlp], r2[kHitIlp];
    if (EAGER) {
#pragma unroll
      for (int u = 0; u < kHitIlp; ++u) {   // all plane records in flight before the first test
        const float4* r = reinterpret_cast<const float4*>(planes + max(pl[u], 0));
        r0[u] = __ldg(r); r1[u] = __ldg(r + 1); r2[u] = __ldg(r + 2);
      }
    }
#pragma unroll
    for (int u = 0; u < kHitIlp; ++u) {
      const int64_t i = c0 + threadIdx.x + static_cast<int64_t>(u) * blockDim.x;
      if (i >= n) continue;
      int res = -1;
      if (pl[u] >= 0) {
        if (!EAGER) {
          const float4* r = reinterpret_cast<const float4*>(planes + pl[u]);
          r0[u] = __ldg(r); r1[u] = __ldg(r + 1); r2[u] = __ldg(r + 2);
        }
        if (x[u] > r0[u].x && x[u] < r1[u].x && y[u] > r0[u].y && y[u] < r1[u].y && z[u] > r0[u].z && z[u] < r1[u].z) {
          double d = static_cast<double>(x[u]) * static_cast<double>(r0[u].w);   // point2PlaneDistance :296-303, in its order
          d = d + static_cast<double>(y[u]) * static_cast<double>(r1[u].w);
          d = d + static_cast<double>(z[u]) * static_cast<double>(r2[u].x);
          d = d + static_cast<double>(r2[u].y);
          if ((d > 0 ? d : -d) <= radius) res = pl[u];
        }
      }
      __stcs(cand + i, res);
    }
  }
}

// One WARP per (scan, ring). cand: [n_scans][H][W]; sel (emission order): [n_scans][W][H].
// The reference collects, per plane, the hit columns of the ring in ascending order and picks hits[step*(s+1)-1] (:127-136).  Hits
// arrive in column order already, so a hit only needs its RANK among the hits of its plane and the plane's total: pass 1 ranks the
// hits 32 columns at a time (__match_any_sync groups equal planes, a small per-warp hash table in shared memory carries the running
// count per plane) and parks the rank in `sel`; pass 2 keeps the hits whose rank is one of the k selected ones.  No sort: the first
// version bitonic-sorted 2048 (plane, column) keys per ring and was the slowest kernel of the map path (0.94 ms at 17 M points).
constexpr int kSelHash = 256;   // distinct planes per ring are far fewer (a ring crosses a 0.5 m voxel a handful of times)
constexpr int kSelWarps = 8;

__device__ __forceinline__ int sel_hash_slot(int* hkey, int plane) {  // find-or-insert with linear probing; one lane at a time
  unsigned h = (static_cast<unsigned>(plane) * 2654435761u) >> 24;    // 8 bits
  for (int probe = 0; probe < kSelHash; ++probe) {
    const int sl = (h + probe) & (kSelHash - 1);
    const int k = hkey[sl];
    if (k == plane) return sl;
    if (k == -1) { hkey[sl] = plane; return sl; }
  }
  return -1;  // table full: cannot happen for W <= 4096 with 256 slots unless > 256 distinct planes hit one ring
}

// concurrent find-or-insert: the leaders of one 32-column step insert their (distinct) planes at the same time
__device__ __forceinline__ int sel_hash_slot_atomic(int* hkey, int plane) {
  unsigned h = (static_cast<unsigned>(plane) * 2654435761u) >> 24;
  for (int probe = 0; probe < kSelHash; ++probe) {
    const int sl = (h + probe) & (kSelHash - 1);
    const int k = atomicCAS(hkey + sl, -1, plane);
    if (k == -1 || k == plane) return sl;
  }
  return -1;
}
__device__ __forceinline__ int sel_hash_find(const int* hkey, int plane) {
  unsigned h = (static_cast<unsigned>(plane) * 2654435761u) >> 24;
  for (int probe = 0; probe < kSelHash; ++probe) {
    const int sl = (h + probe) & (kSelHash - 1);
    if (hkey[sl] == plane) return sl;
  }
  return 0;   // not reached: every plane of the ring was inserted by pass 1
}

__global__ void __launch_bounds__(kSelWarps * 32) assoc_select_kernel(const int32_t* __restrict__ cand, const lvi_point_xyzit* __restrict__ raw, int n_rings,
                                                                       int W, int H, int k_per_ring, int32_t* __restrict__ sel, int* __restrict__ overflow,
                                                                       int32_t* __restrict__ ring_overflow) {
  __shared__ int hkey[kSelWarps][kSelHash], hcnt[kSelWarps][kSelHash];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned FULL = 0xffffffffu;
  const int ring = blockIdx.x * kSelWarps + warp;
  if (ring >= n_rings) return;
  const int scan = ring / H, h = ring % H;
  bool full = false;   // this lane could not place its plane in the table
  int* hk = hkey[warp]; int* hc = hcnt[warp];
  for (int i = lane; i < kSelHash; i += 32) { hk[i] = -1; hc[i] = 0; }
  __syncwarp();
  const int32_t* row = cand + (static_cast<int64_t>(scan) * H + h) * W;
  int32_t* selrow = sel + static_cast<int64_t>(scan) * W * H + h;  // + w*H
  // pass 1: rank of every hit among the hits of its plane
  for (int w0 = 0; w0 < W; w0 += 32) {
    const int w = w0 + lane;
    const int c = w < W ? row[w] : -1;
    const unsigned hits = __ballot_sync(FULL, c >= 0);
    int rank = -1;
    if (c >= 0) {
      const unsigned grp = __match_any_sync(hits, c);
      const int leader = __ffs(grp) - 1;
      int base = 0;
      // leaders update the table one after another (usually 1-3 distinct planes per 32 columns)
      unsigned leaders = __ballot_sync(hits, lane == leader);
      while (leaders) {
        const int l = __ffs(leaders) - 1;
        leaders &= leaders - 1;
        if (lane == l) {
          const int sl = sel_hash_slot(hk, c);
          if (sl < 0) { full = true; base = 0; }
          else { base = hc[sl]; hc[sl] = base + __popc(grp); }
        }
        __syncwarp(hits);
      }
      base = __shfl_sync(grp, base, leader);
      rank = base + __popc(grp & ((1u << lane) - 1));
    }
    if (w < W) selrow[static_cast<int64_t>(w) * H] = rank;
  }
  __syncwarp();
  // More than kSelHash distinct surfels on one ring (a ground ring at 20 - 30 m radius crosses 250 - 380 half-metre voxels): the ring is
  // handed to assoc_select_dense_kernel, which has no table.  The reference has no such limit (surfel_association.cpp:111-159).
  if (__any_sync(FULL, full)) {
    if (lane == 0) { ring_overflow[ring] = 1; atomicExch(overflow, 1); }
    return;
  }
  // pass 2: keep hits[step*(s+1)-1], s < k, of planes with >= 2k hits (:128-136); timestamp == 0 points are never emitted (:141-143)
  for (int w0 = 0; w0 < W; w0 += 32) {
    const int w = w0 + lane;
    if (w >= W) continue;
    const int c = row[w];
    int out = -1;
    if (c >= 0) {
      const int rank = selrow[static_cast<int64_t>(w) * H];
      unsigned hsh = (static_cast<unsigned>(c) * 2654435761u) >> 24;
      int total = 0;
      for (int probe = 0; probe < kSelHash; ++probe) { const int sl = (hsh + probe) & (kSelHash - 1); if (hk[sl] == c) { total = hc[sl]; break; } if (hk[sl] == -1) break; }
      if (total >= k_per_ring * 2) {
        int step = total / (k_per_ring + 1);
        step = step > 1 ? step : 1;
        const int r1 = rank + 1;
        if (r1 % step == 0 && r1 / step >= 1 && r1 / step <= k_per_ring) {
          const double ts = raw[(static_cast<int64_t>(scan) * H + h) * W + w].timestamp;
          if (ts != 0.0) out = c;
        }
      }
    }
    selrow[static_cast<int64_t>(w) * H] = out;
  }
}

// Fallback for the rings assoc_select_kernel flagged: one CTA per such ring, the ring's candidates in shared memory, rank and total of
// every hit by direct counting over the ring (O(W^2) compares: at most 8 M for W = 4096, and only for the few rings that need it).
__global__ void __launch_bounds__(256) assoc_select_dense_kernel(const int32_t* __restrict__ cand, const lvi_point_xyzit* __restrict__ raw, int n_rings,
                                                                 int W, int H, int k_per_ring, int32_t* __restrict__ sel,
                                                                 const int32_t* __restrict__ ring_overflow) {
  __shared__ int32_t row_s[4096];
  for (int ring = blockIdx.x; ring < n_rings; ring += gridDim.x) {
    if (!ring_overflow[ring]) continue;   // block-uniform
    const int scan = ring / H, h = ring % H;
    const int32_t* row = cand + (static_cast<int64_t>(scan) * H + h) * W;
    int32_t* selrow = sel + static_cast<int64_t>(scan) * W * H + h;
    __syncthreads();
    for (int w = threadIdx.x; w < W; w += blockDim.x) row_s[w] = row[w];
    __syncthreads();
    for (int w = threadIdx.x; w < W; w += blockDim.x) {
      const int c = row_s[w];
      int out = -1;
      if (c >= 0) {
        int rank = 0, total = 0;
        for (int v = 0; v < W; ++v) { const int same = row_s[v] == c; total += same; rank += same & (v < w); }
        if (total >= k_per_ring * 2) {
          int step = total / (k_per_ring + 1);
          step = step > 1 ? step : 1;
          const int r1 = rank + 1;
          if (r1 % step == 0 && r1 / step >= 1 && r1 / step <= k_per_ring) {
            const double ts = raw[(static_cast<int64_t>(scan) * H + h) * W + w].timestamp;
            if (ts != 0.0) out = c;
          }
        }
      }
      selrow[static_cast<int64_t>(w) * H] = out;
    }
  }
}

struct SelFlag {
  const int32_t* sel;
  __host__ __device__ int operator()(int64_t i) const { return sel[i] >= 0 ? 1 : 0; }
};

__global__ void __launch_bounds__(256) assoc_flag_kernel(const int32_t* __restrict__ sel, int64_t n, int32_t* __restrict__ flag) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    flag[i] = sel[i] >= 0 ? 1 : 0;
}

__global__ void __launch_bounds__(256) assoc_emit_kernel(const int32_t* __restrict__ sel, const int32_t* __restrict__ rank, int64_t n, int W, int H,
                                                         int time_step, const char* __restrict__ scan_map, size_t stride,
                                                         const lvi_point_xyzit* __restrict__ raw, lvi_surfel_point* __restrict__ out, int64_t cap) {
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < n; e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int pl = sel[e];
    if (pl < 0) continue;
    const int r = rank[e];
    if (r % time_step != 0) continue;  // averageTimeDownSmaple :240-244
    const int64_t o = r / time_step;
    if (o >= cap) continue;
    const int64_t scan = e / (static_cast<int64_t>(W) * H);
    const int64_t rem = e - scan * W * H;
    const int w = static_cast<int>(rem / H), h = static_cast<int>(rem % H);
    const int64_t idx = (scan * H + h) * W + w;
    const lvi_point_xyzit p = raw[idx];
    const float* q = reinterpret_cast<const float*>(scan_map + idx * stride);
    lvi_surfel_point sp;
    sp.timestamp = p.timestamp;
    sp.point[0] = p.x; sp.point[1] = p.y; sp.point[2] = p.z;
    sp.point_in_map[0] = q[0]; sp.point_in_map[1] = q[1]; sp.point_in_map[2] = q[2];
    sp.plane_id = pl;
    out[o] = sp;
  }
}

// ---- per-scan selection: one CTA per scan ---------------------------------------------------------------------------------------
// assoc_hit_kernel leaves the candidate plane of every point (cand[scan][h][w]).  For scans of up to kScanSelMaxPts points one CTA then
// does the rest of a scan without another pass over HBM beyond reading those 4 B per point:
//   B  one warp per ring: pass 1 counts the hits of every plane (match_any + warp hash table), pass 2 walks the ring again with running
//      counts and marks hits[step*(s+1)-1] in a bitmap laid out in emission order [w][h] (surfel_association.cpp:127-158)
//   C  ordered compaction of the bitmap -> (position, plane) list of the scan + its length
// A tiny exclusive scan over the per-scan lengths then gives every selected point its global emission rank, and assoc_emit_list_kernel
// writes every time_step-th of them as a SurfelPoint (:240-244).  (The first version ranked into a [scan][w][h] int array with strided
// writes, flagged, scanned and re-read it: 4 passes of 4 B per point, 0.65 ms at C2.)
constexpr int kScanSelThreads = 512;
constexpr int kScanSelWarps = kScanSelThreads / 32;
constexpr int kScanSelMaxPts = 1 << 17;   // bitmap of 16 KB in shared memory
constexpr int kSelAhead = 8;               // 32-column steps of a ring whose loads are issued together

struct FusedSel { int32_t e; int32_t plane; };   // e = w * H + h: position inside the scan in emission order

__global__ void __launch_bounds__(kScanSelThreads) assoc_scan_select_kernel(const int32_t* __restrict__ cand, const lvi_point_xyzit* __restrict__ raw,
                                                                            int n_scans, int W, int H, int k_per_ring, FusedSel* __restrict__ lists,
                                                                            int32_t* __restrict__ counts) {
  extern __shared__ __align__(16) unsigned char sel_smem[];
  __shared__ int s_part[kScanSelWarps + 1];
  const int HW = H * W;
  const int n_words = (HW + 31) >> 5;
  unsigned* bits = reinterpret_cast<unsigned*>(sel_smem);                   // [n_words] emission order
  int* hkey = reinterpret_cast<int*>(bits + n_words);                       // [warps][kSelHash]
  int* htot = hkey + kScanSelWarps * kSelHash;
  int* hrun = htot + kScanSelWarps * kSelHash;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned FULL = 0xffffffffu;
  for (int scan = blockIdx.x; scan < n_scans; scan += gridDim.x) {
    __syncthreads();
    const int64_t base = static_cast<int64_t>(scan) * HW;
    for (int w = tid; w < n_words; w += kScanSelThreads) bits[w] = 0u;
    __syncthreads();
    // ---- B: per-ring selection, one warp per ring
    int* hk = hkey + warp * kSelHash; int* ht = htot + warp * kSelHash; int* hr = hrun + warp * kSelHash;
    for (int h = warp; h < H; h += kScanSelWarps) {
      const int32_t* row = cand + base + static_cast<int64_t>(h) * W;
      for (int i = lane; i < kSelHash; i += 32) { hk[i] = -1; ht[i] = 0; hr[i] = 0; }
      __syncwarp();
      bool full = false;
      // both passes read the ring kSelAhead x 32 columns at a time: the loads of a batch are in flight together (the walk itself is serial
      // in the column, so one load per step would pay one trip to L2 per 32 columns)
      for (int wb = 0; wb < W; wb += 32 * kSelAhead) {   // pass 1: hits per plane
        int cbuf[kSelAhead];
#pragma unroll
        for (int u = 0; u < kSelAhead; ++u) { const int w = wb + 32 * u + lane; cbuf[u] = w < W ? __ldg(row + w) : -1; }
#pragma unroll
        for (int u = 0; u < kSelAhead; ++u) {
        if (wb + 32 * u >= W) break;
        const int c = cbuf[u];
        const unsigned hits = __ballot_sync(FULL, c >= 0);
        if (c >= 0) {
          const unsigned grp = __match_any_sync(hits, c);
          if (lane == __ffs(grp) - 1) {   // group leaders hold distinct planes: their table slots are distinct too
            const int sl = sel_hash_slot_atomic(hk, c);
            if (sl < 0) full = true; else ht[sl] += __popc(grp);
          }
        }
        __syncwarp();
        }
      }
      full = __any_sync(FULL, full);
      for (int wb = 0; wb < W; wb += 32 * kSelAhead) {   // pass 2: rank of every hit inside its plane's hit list -> the k picked ones
        int cbuf[kSelAhead];
#pragma unroll
        for (int u = 0; u < kSelAhead; ++u) { const int w = wb + 32 * u + lane; cbuf[u] = w < W ? __ldg(row + w) : -1; }
#pragma unroll
        for (int u = 0; u < kSelAhead; ++u) {
        if (wb + 32 * u >= W) break;
        const int w = wb + 32 * u + lane;
        const int c = cbuf[u];
        const unsigned hits = __ballot_sync(FULL, c >= 0);
        if (c >= 0) {
          int rank, total;
          if (!full) {
            const unsigned grp = __match_any_sync(hits, c);
            const int leader = __ffs(grp) - 1;
            int basev = 0, tot = 0;
            if (lane == leader) {
              const int sl = sel_hash_find(hk, c);
              basev = hr[sl]; tot = ht[sl];
              hr[sl] = basev + __popc(grp);
            }
            basev = __shfl_sync(grp, basev, leader); total = __shfl_sync(grp, tot, leader);
            rank = basev + __popc(grp & ((1u << lane) - 1));
          } else {   // more distinct planes than the table holds: count directly over the ring (rare; no limit in the reference)
            rank = 0; total = 0;
            for (int v = 0; v < W; ++v) { const int same = __ldg(row + v) == c; total += same; rank += same & (v < w); }
          }
          if (total >= k_per_ring * 2) {
            int step = total / (k_per_ring + 1);
            step = step > 1 ? step : 1;
            const int r1 = rank + 1;
            if (r1 % step == 0 && r1 / step >= 1 && r1 / step <= k_per_ring) {
              const double ts = raw[base + static_cast<int64_t>(h) * W + w].timestamp;
              if (ts != 0.0) { const int e = w * H + h; atomicOr(bits + (e >> 5), 1u << (e & 31)); }
            }
          }
        }
        __syncwarp();
        }
      }
    }
    __syncthreads();
    // ---- C: ordered compaction of the emission-order bitmap
    FusedSel* list = lists + base;
    int running = 0;
    for (int w0 = 0; w0 < n_words; w0 += kScanSelThreads) {
      const int wd = w0 + tid;
      const unsigned word = wd < n_words ? bits[wd] : 0u;
      const int cnt = __popc(word);
      int inc = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
      if (lane == 31) s_part[warp] = inc;
      __syncthreads();
      if (tid == 0) { int a = 0; for (int q = 0; q < kScanSelWarps; ++q) { const int t = s_part[q]; s_part[q] = a; a += t; } s_part[kScanSelWarps] = a; }
      __syncthreads();
      int o = running + s_part[warp] + inc - cnt;
      unsigned wbits = word;
      while (wbits) {
        const int b = __ffs(wbits) - 1;
        wbits &= wbits - 1;
        const int e = 32 * wd + b;
        const int w = e / H, h = e - w * H;
        list[o++] = FusedSel{e, __ldg(cand + base + static_cast<int64_t>(h) * W + w)};
      }
      running += s_part[kScanSelWarps];
      __syncthreads();
    }
    if (tid == 0) counts[scan] = running;
  }
}

__global__ void __launch_bounds__(256) assoc_emit_list_kernel(const FusedSel* __restrict__ lists, const int32_t* __restrict__ counts, const int32_t* __restrict__ bases,
                                                              int n_scans, int W, int H, int time_step, const char* __restrict__ scan_map, size_t stride,
                                                              const lvi_point_xyzit* __restrict__ raw, lvi_surfel_point* __restrict__ out, int64_t cap) {
  const int HW = W * H;
  for (int scan = blockIdx.x; scan < n_scans; scan += gridDim.x) {
    const int cnt = counts[scan], b0 = bases[scan];
    const FusedSel* list = lists + static_cast<int64_t>(scan) * HW;
    // first j with (b0 + j) % time_step == 0
    const int j0 = (time_step - b0 % time_step) % time_step;
    for (int j = j0 + threadIdx.x * time_step; j < cnt; j += blockDim.x * time_step) {   // averageTimeDownSmaple :240-244
      const int64_t o = (static_cast<int64_t>(b0) + j) / time_step;
      if (o >= cap) continue;
      const FusedSel fs = list[j];
      const int w = fs.e / H, h = fs.e - w * H;
      const int64_t idx = static_cast<int64_t>(scan) * HW + static_cast<int64_t>(h) * W + w;
      const lvi_point_xyzit p = raw[idx];
      const float* q = reinterpret_cast<const float*>(scan_map + idx * stride);
      lvi_surfel_point sp;
      sp.timestamp = p.timestamp;
      sp.point[0] = p.x; sp.point[1] = p.y; sp.point[2] = p.z;
      sp.point_in_map[0] = q[0]; sp.point_in_map[1] = q[1]; sp.point_in_map[2] = q[2];
      sp.plane_id = fs.plane;
      out[o] = sp;
    }
  }
}

// a-3'  associateVisualPointsWithPlanes inner test (L/src/core/surfel_association.cpp:196-210): thread per landmark sweeps the
// plane list (<= 1e4 landmarks x <= 1e4 planes, all planes L2-resident); the LAST matching plane wins (:206).
__global__ void __launch_bounds__(128) assoc_landmark_kernel(const double* __restrict__ pts, int64_t n, int n_planes, const double* __restrict__ p4,
                                                             const double* __restrict__ bmin, const double* __restrict__ bmax, double radius2,
                                                             int32_t* __restrict__ out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
  int best = -1;
  for (int k = 0; k < n_planes; ++k) {
    const double* mn = bmin + 3 * k; const double* mx = bmax + 3 * k;
    if (x > mn[0] && x < mx[0] && y > mn[1] && y < mx[1] && z > mn[2] && z < mx[2]) {
      if (p2plane(x, y, z, p4 + 4 * k) <= radius2) best = k;
    }
  }
  out[i] = best;
}

static void associate_device(lvi_ctx* ctx, const lvi_voxel_map* m, const lvi_surfel_set* s, const void* map_d, size_t stride,
                             const lvi_point_xyzit* raw_d, int n_scans, int W, int H, double radius, int k, int time_step,
                             lvi_surfel_point* out_d, int64_t cap, int64_t* n_out, int64_t* n_all) {
  LVI_REQUIRE(W > 0 && H > 0 && n_scans > 0, LVI_ERR_INVALID, "lvi_associate: empty scan batch");
  LVI_REQUIRE(W <= 4096, LVI_ERR_INVALID, "lvi_associate: scan width > 4096 not supported");
  LVI_REQUIRE(time_step >= 1 && k >= 1, LVI_ERR_INVALID, "lvi_associate: bad k_per_ring/time_step");
  const int64_t n = static_cast<int64_t>(n_scans) * W * H;
  LVI_REQUIRE(n < 2147483647LL, LVI_ERR_INVALID, "lvi_associate: batch too large (split the scans)");
  cudaStream_t st = ctx->stream;
  if (s->n_planes == 0) { if (n_out) *n_out = 0; if (n_all) *n_all = 0; return; }
  DBuf<int32_t> cand(n);
  DBuf<PlaneRec> prec(static_cast<size_t>(s->n_planes));
  LVI_LAUNCH(ctx, assoc_plane_rec_kernel, static_cast<int>((s->n_planes + 255) / 256), 256, 0, s->p4.p, s->bmin.p, s->bmax.p, static_cast<int>(s->n_planes), prec.p);
  static const int variant = std::getenv("LVI_ASSOC_HIT_VARIANT") ? std::atoi(std::getenv("LVI_ASSOC_HIT_VARIANT")) : 0;
  const char* map_c = static_cast<const char*>(map_d);
  const int32_t* nul = nullptr;
#define LVI_HIT(D, I, E, c2l, lk, nl, l2p) \
  LVI_LAUNCH_AS(ctx, "assoc_hit_kernel", (assoc_hit_kernel<D, I, E>), grid_for(n, 256 * I, ctx->sm_count, 8), 256, 0, map_c, stride, n, m->grid_d.p, c2l, lk, nl, l2p, prec.p, radius, cand.p)
  if (m->cell2leaf.n && variant != 9) {
    if (s->cell2plane.n != m->cell2leaf.n) {   // first association with this surfel set
      s->cell2plane.alloc(m->cell2leaf.n);
      LVI_CUDA(cudaMemsetAsync(s->cell2plane.p, 0xff, sizeof(int32_t) * s->cell2plane.n, st));
      LVI_LAUNCH(ctx, assoc_cell2plane_kernel, static_cast<int>((s->n_planes + 255) / 256), 256, 0, s->key.p, static_cast<int>(s->n_planes), s->cell2plane.p);
    }
    if (variant == 1) LVI_HIT(true, 8, true, s->cell2plane.p, nul, 0, nul);
    else if (variant == 2) LVI_HIT(true, 8, false, s->cell2plane.p, nul, 0, nul);
    else if (variant == 3) LVI_HIT(true, 2, true, s->cell2plane.p, nul, 0, nul);
    else if (variant == 4) LVI_HIT(true, 2, false, s->cell2plane.p, nul, 0, nul);
    else if (variant == 5) LVI_HIT(true, 4, true, s->cell2plane.p, nul, 0, nul);
    else LVI_HIT(true, 4, false, s->cell2plane.p, nul, 0, nul);
  } else {
    LVI_HIT(false, 4, false, m->cell2leaf.n ? m->cell2leaf.p : nul, m->leaf_key.p, static_cast<int>(m->n_leaves), s->leaf2plane.p);
  }
#undef LVI_HIT
  if (static_cast<int64_t>(W) * H <= kScanSelMaxPts && !std::getenv("LVI_ASSOC_UNFUSED")) {   // selection + emission order of a scan inside one CTA
    const int HW = W * H;
    const size_t smem = static_cast<size_t>((HW + 31) / 32) * 4 + static_cast<size_t>(kScanSelWarps) * kSelHash * 12;
    if (smem > ctx->ks.assoc_attr) {
      LVI_CUDA(cudaFuncSetAttribute(assoc_scan_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      ctx->ks.assoc_attr = smem;
    }
    DBuf<FusedSel> lists(static_cast<size_t>(n));
    DBuf<int32_t> counts(n_scans), bases(n_scans);
    LVI_LAUNCH(ctx, assoc_scan_select_kernel, std::min(n_scans, ctx->sm_count * 4), kScanSelThreads, smem, cand.p, raw_d, n_scans, W, H, k, lists.p, counts.p);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, counts.p, bases.p, n_scans, st);
    DBuf<char> tmp(tb + 16);
    LVI_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, counts.p, bases.p, n_scans, st));
    ctx->launches += 1;
    // the emission needs no host decision: it runs before the host learns the total
    if (out_d && cap > 0)
      LVI_LAUNCH(ctx, assoc_emit_list_kernel, std::min(n_scans, ctx->sm_count * 8), 256, 0, lists.p, counts.p, bases.p, n_scans, W, H, time_step,
                 static_cast<const char*>(map_d), stride, raw_d, out_d, cap);
    int last_base = 0, last_count = 0;
    LVI_CUDA(cudaMemcpyAsync(&last_base, bases.p + n_scans - 1, 4, cudaMemcpyDeviceToHost, st));
    LVI_CUDA(cudaMemcpyAsync(&last_count, counts.p + n_scans - 1, 4, cudaMemcpyDeviceToHost, st));
    LVI_CUDA(cudaStreamSynchronize(st));
    const int64_t total = static_cast<int64_t>(last_base) + last_count;
    if (n_all) *n_all = total;
    if (n_out) *n_out = (total + time_step - 1) / time_step;
    return;
  }
  DBuf<int32_t> sel(n), flag(n), rank(n);
  const int n_rings = n_scans * H;
  DBuf<int> overflow(1);
  DBuf<int32_t> ring_overflow(n_rings);
  overflow.zero(st); ring_overflow.zero(st);
  LVI_LAUNCH(ctx, assoc_select_kernel, (n_rings + kSelWarps - 1) / kSelWarps, kSelWarps * 32, 0, cand.p, raw_d, n_rings, W, H, k, sel.p, overflow.p,
             ring_overflow.p);
  // rings with more distinct surfels than the warp table holds (rare; none at C2): every CTA looks at its rings' flags and leaves at once
  // when there is nothing to do, so the launch costs a few microseconds and saves a host round trip
  LVI_LAUNCH(ctx, assoc_select_dense_kernel, std::min(n_rings, ctx->sm_count * 8), 256, 0, cand.p, raw_d, n_rings, W, H, k, sel.p, ring_overflow.p);
  LVI_LAUNCH(ctx, assoc_flag_kernel, grid_for(n, 256, ctx->sm_count, 8), 256, 0, sel.p, n, flag.p);
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, flag.p, rank.p, static_cast<int>(n), st);
  DBuf<char> tmp(tb + 16);
  LVI_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, flag.p, rank.p, static_cast<int>(n), st));
  ctx->launches += 2;
  int last_rank = 0, last_flag = 0;
  LVI_CUDA(cudaMemcpyAsync(&last_rank, rank.p + n - 1, 4, cudaMemcpyDeviceToHost, st));
  LVI_CUDA(cudaMemcpyAsync(&last_flag, flag.p + n - 1, 4, cudaMemcpyDeviceToHost, st));
  LVI_CUDA(cudaStreamSynchronize(st));
  const int64_t total = static_cast<int64_t>(last_rank) + last_flag;
  if (n_all) *n_all = total;
  if (n_out) *n_out = (total + time_step - 1) / time_step;
  if (out_d && cap > 0 && total > 0)
    LVI_LAUNCH(ctx, assoc_emit_kernel, grid_for(n, 256, ctx->sm_count, 8), 256, 0, sel.p, rank.p, n, W, H, time_step, static_cast<const char*>(map_d),
               stride, raw_d, out_d, cap);
  LVI_CUDA(cudaStreamSynchronize(st));
}

}  // namespace lvi

using namespace lvi;

extern "C" {

int lvi_associate_d(lvi_ctx* ctx, const lvi_voxel_map* m, const lvi_surfel_set* s, const void* scans_in_map_d, size_t map_stride_bytes,
                    const lvi_point_xyzit* scans_raw_d, int32_t n_scans, int32_t W, int32_t H, double radius, int32_t k_per_ring, int32_t time_step,
                    lvi_surfel_point* out_d, int64_t cap, int64_t* n_out, int64_t* n_all) {
  return guarded([&] {
    LVI_REQUIRE(ctx && m && s && scans_in_map_d && scans_raw_d, LVI_ERR_INVALID, "lvi_associate_d: null argument");
    activate(ctx);
    associate_device(ctx, m, s, scans_in_map_d, map_stride_bytes, scans_raw_d, n_scans, W, H, radius, k_per_ring, time_step, out_d, cap, n_out, n_all);
  });
}

int lvi_associate_batch(lvi_ctx* ctx, const lvi_voxel_map* m, const lvi_surfel_set* s, const lvi_scan_batch* batch, const lvi_point_xyzit* scans_raw_d,
                        int32_t W, int32_t H, double radius, int32_t k_per_ring, int32_t time_step, lvi_surfel_point* out_d, int64_t cap, int64_t* n_out,
                        int64_t* n_all) {
  return guarded([&] {
    LVI_REQUIRE(ctx && m && s && batch && scans_raw_d, LVI_ERR_INVALID, "lvi_associate_batch: null argument");
    LVI_REQUIRE(static_cast<int64_t>(W) * H == batch->pts_per_scan && batch->n == static_cast<int64_t>(batch->n_scans) * batch->pts_per_scan, LVI_ERR_INVALID,
                "lvi_associate_batch: W x H does not match the batch");
    activate(ctx);
    associate_device(ctx, m, s, batch->pts.p, sizeof(float4), scans_raw_d, batch->n_scans, W, H, radius, k_per_ring, time_step, out_d, cap, n_out, n_all);
  });
}

int lvi_associate(lvi_ctx* ctx, const lvi_voxel_map* m, const lvi_surfel_set* s, const void* scans_in_map, size_t map_stride_bytes,
                  const lvi_point_xyzit* scans_raw, int32_t n_scans, int32_t W, int32_t H, double radius, int32_t k_per_ring, int32_t time_step,
                  lvi_surfel_point* out, int64_t cap, int64_t* n_out, int64_t* n_all) {
  return guarded([&] {
    LVI_REQUIRE(ctx && m && s && scans_in_map && scans_raw, LVI_ERR_INVALID, "lvi_associate: null argument");
    activate(ctx);
    const int64_t n = static_cast<int64_t>(n_scans) * W * H;
    DBuf<char> map_d(static_cast<size_t>(n) * map_stride_bytes);
    DBuf<lvi_point_xyzit> raw_d(n);
    LVI_CUDA(cudaMemcpyAsync(map_d.p, scans_in_map, map_d.n, cudaMemcpyHostToDevice, ctx->stream));
    LVI_CUDA(cudaMemcpyAsync(raw_d.p, scans_raw, sizeof(lvi_point_xyzit) * n, cudaMemcpyHostToDevice, ctx->stream));
    DBuf<lvi_surfel_point> out_d(out && cap > 0 ? cap : 0);
    int64_t no = 0, na = 0;
    associate_device(ctx, m, s, map_d.p, map_stride_bytes, raw_d.p, n_scans, W, H, radius, k_per_ring, time_step, out_d.p, out ? cap : 0, &no, &na);
    if (out && cap > 0 && no > 0) {
      LVI_CUDA(cudaMemcpyAsync(out, out_d.p, sizeof(lvi_surfel_point) * std::min(no, cap), cudaMemcpyDeviceToHost, ctx->stream));
      LVI_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (n_out) *n_out = no;
    if (n_all) *n_all = na;
  });
}

int lvi_associate_landmarks(lvi_ctx* ctx, const lvi_surfel_set* s, const double* pts, int64_t n, double radius, int32_t* plane_out) {
  return guarded([&] {
    LVI_REQUIRE(ctx && s && plane_out && (pts || n == 0), LVI_ERR_INVALID, "lvi_associate_landmarks: null argument");
    if (n == 0) return;
    activate(ctx);
    DBuf<double> pd(3 * static_cast<size_t>(n));
    DBuf<int32_t> od(n);
    pd.upload(pts, 3 * static_cast<size_t>(n), ctx->stream);
    LVI_LAUNCH(ctx, assoc_landmark_kernel, static_cast<int>((n + 127) / 128), 128, 0, pd.p, n, static_cast<int>(s->n_planes), s->p4.p, s->bmin.p, s->bmax.p,
               radius * 2, od.p);
    od.download(plane_out, n, ctx->stream);
    LVI_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

}  // extern "C"
