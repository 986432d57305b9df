// assemble.cu — the normal equations J^T J / J^T r by owner-computes TILE GATHER on the fp64 tensor pipe (north-star
// "J^T J as a small dense contraction on tensor cores").
//
// Replaces Ceres' block-sparse Jacobian + SchurEliminator / normal-equation build behind kontiki::TrajectoryEstimator::Solve
// (K/trajectory_estimator.h:38-68) and this library's first version, which scattered one fp64 atomic per (run of residuals, column pair)
// into the tile store (17.8 M atomics per C2 iteration for the camera table alone, 0.63 ms).
//
// The sparsity structure of a problem is fixed over the LM iterations, so it is turned into a PLAN once per problem, on the device:
//   pos_kernel        linear-system position of every Jacobian column of every residual (a second-window column on a knot the first
//                     window already holds is merged into the first, so positions are distinct within a residual)
//   pairs_kernel      the 32x32 tiles each residual touches: one key (tile << 32 | table << 29 | residual) per (tile, residual)
//   CUB radix sort    keys by tile; a max-scan + select cut the sorted list into WORK ITEMS of <= kChunk residuals of one tile
// Every iteration: jacobian_kernel<TYPE> (problem.cu) writes the loss-corrected Jacobian rows and residuals to HBM (90 MB at C2, L2
// resident), then gather_kernel gives one warp per work item: it scatters the rows of the item's residuals into two zero-filled
// shared-memory panels P_I, P_J [rows x 32] (the columns that fall into the tile's row / column range) and accumulates
// tile += P_I^T P_J with mma.sync.m8n8k4.f64 (16 accumulator fragments per warp).  A tile owned by ONE item is written with plain
// stores -- no atomics, bitwise reproducible; only tiles with more than kChunk contributions (the arrow border and its corner, which
// every surfel residual touches) are split and flushed with one atomic per element per item.  The gradient J^T r falls out of the
// diagonal tiles' P_I panels.
#include <cub/cub.cuh>

#include <cstdlib>

#include "problem.cuh"

namespace lvi {

#ifndef LVI_GATHER_CHUNK
#define LVI_GATHER_CHUNK 96
#endif
constexpr int kChunk = LVI_GATHER_CHUNK;         // residuals per work item
#ifndef LVI_GATHER_WARPS
#define LVI_GATHER_WARPS 4
#endif
constexpr int kGatherWarps = LVI_GATHER_WARPS;
constexpr int kPanelRows = 32;

// ---- plan ---------------------------------------------------------------------------------------------------------------------------
template <int TYPE>
__global__ void pos_kernel(ProblemView P, int* __restrict__ pos) {
  constexpr int COLS = TYPE == RT_GYRO ? 15 : TYPE == RT_ACCEL ? 29 : TYPE == RT_SURFEL ? 54 : TYPE == RT_CAM ? 55 : TYPE == RT_CAMSURF ? 60 : 12;
  constexpr bool two_eval = TYPE == RT_SURFEL || TYPE == RT_CAM || TYPE == RT_CAMSURF;
  const ResTable& T = P.tab[TYPE];
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(T.n) * COLS) return;
  const int i = static_cast<int>(idx / COLS), c = static_cast<int>(idx % COLS);
  int p = col_pos<TYPE>(P, i, c);
  if (TYPE == RT_CAM && c == 54) p = -1;   // inverse depth: eliminated by the Schur complement, its row lives in SchurView
  if (two_eval && c >= 24 && c < 48) {     // knot shared with the first window: jacobian_kernel has folded this column into the first window's
    const int j = T.i0b[i] + ((c - 24) % 12) / 3 - T.i0a[i];
    if (j >= 0 && j <= 3) p = -1;
  }
  pos[idx] = p;
}

__device__ __forceinline__ const int* set_pos(const RowSetView& S, int res) { return S.pos + (S.boff ? static_cast<size_t>(S.boff[res]) : static_cast<size_t>(res) * S.ncols); }
__device__ __forceinline__ int set_ncols(const RowSetView& S, int res) { return S.bcols ? S.bcols[res] : S.ncols; }
__device__ __forceinline__ const double* set_block(const RowSetView& S, int res) { return S.J + (S.boff ? static_cast<size_t>(S.boff[res]) : static_cast<size_t>(res) * S.rstride); }
__device__ __forceinline__ int set_blen(const RowSetView& S, int res) { return S.bcols ? ((S.bcols[res] + 1) & ~1) : S.rstride; }

struct TileDims { int nb, NT, T, RB, TPC; };
__device__ __forceinline__ int virtual_tile_row(const TileDims& D, int p) { return p < D.nb ? p >> kTileLog : D.NT + ((p - D.nb) >> kTileLog); }

// the distinct tile rows a residual touches (sorted) -> one (tile, residual) key per pair I >= J
template <bool EMIT>
__global__ void pairs_kernel(RowSetView S, int set, TileDims D, int res_base, const int* __restrict__ offsets,
                             int* __restrict__ counts, unsigned long long* __restrict__ keys, int* __restrict__ fail) {
  const int lo = S.lo, hi = S.hi;
  const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hi) return;
  const int* pr = set_pos(S, i);
  const int ncols = set_ncols(S, i);
  constexpr int kMaxRows = 48;   // a landmark's row spans its reference window and every observation window
  int vt[kMaxRows];
  int nv = 0;
  for (int c = 0; c < ncols; ++c) {
    const int p = pr[c];
    if (p < 0) continue;
    const int v = virtual_tile_row(D, p);
    bool seen = false;
    for (int k = 0; k < nv; ++k) seen = seen || vt[k] == v;
    if (seen) continue;
    if (nv == kMaxRows) { *fail = 3; break; }
    int k = nv++;
    while (k > 0 && vt[k - 1] > v) { vt[k] = vt[k - 1]; --k; }
    vt[k] = v;
  }
  const int slot = res_base + (i - lo);
  if (!EMIT) { counts[slot] = nv * (nv + 1) / 2; return; }
  int o = offsets[slot];
  for (int a = 0; a < nv; ++a)
    for (int b = 0; b <= a; ++b) {
      const int I = vt[a], J = vt[b];
      long long tile;
      if (J < D.NT) {
        if (I < D.NT && I - J > D.T) *fail = 4;   // outside the band: the bandwidth of the lowering is wrong
        tile = static_cast<long long>(J) * D.TPC + (I < D.NT ? I - J : D.T + 1 + (I - D.NT));
      } else {
        tile = static_cast<long long>(D.NT) * D.TPC + static_cast<long long>(J - D.NT) * D.RB + (I - D.NT);
      }
      keys[o++] = (static_cast<unsigned long long>(tile) << 32) | (static_cast<unsigned long long>(set) << 29) | static_cast<unsigned>(i);
    }
}

__global__ void seg_head_kernel(const unsigned long long* __restrict__ keys, int n, int* __restrict__ head_idx) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const bool head = e == 0 || (keys[e] >> 32) != (keys[e - 1] >> 32);
  head_idx[e] = head ? e : 0;
}
__global__ void item_flag_kernel(const int* __restrict__ seg_start, int n, unsigned char* __restrict__ flag) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  flag[e] = ((e - seg_start[e]) % kChunk) == 0 ? 1 : 0;   // a segment head has e == seg_start[e]
}

// ---- gather -------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma_m8n8k4(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// Per (tile, residual) entry: which Jacobian column feeds each of the 32 panel columns -- bytes [0,32) for the tile's row range, [32,64) for
// its column range (255: none).  Static, so the per-iteration gather does no position arithmetic at all.
// DescT = unsigned char for the Jacobian sets (<= 60 columns), unsigned short for the Schur rows (a long track couples > 255 parameters)
template <class DescT> struct DescNone;
template <> struct DescNone<unsigned char> { static constexpr unsigned value = 0xffu; };
template <> struct DescNone<unsigned short> { static constexpr unsigned value = 0xffffu; };
constexpr int kDescEntries = 64;
__device__ __forceinline__ void tile_ranges(const BandSys& H, long long tile, int& rowbase, int& rowend, int& colbase, int& colend) {
  const long long n_band_tiles = static_cast<long long>(H.NT) * H.TPC;
  if (tile < n_band_tiles) {
    const int J = static_cast<int>(tile / H.TPC), s = static_cast<int>(tile % H.TPC);
    colbase = J << kTileLog; colend = min(colbase + kTile, H.nb);        // the last band tile ends at nb: border positions start there
    if (s <= H.T) { rowbase = (J + s) << kTileLog; rowend = min(rowbase + kTile, H.nb); }
    else { rowbase = H.nb + ((s - H.T - 1) << kTileLog); rowend = rowbase + kTile; }
  } else {
    const int c = static_cast<int>(tile - n_band_tiles), bj = c / H.RB, bi = c % H.RB;
    rowbase = H.nb + (bi << kTileLog); rowend = rowbase + kTile;
    colbase = H.nb + (bj << kTileLog); colend = colbase + kTile;
  }
}
template <class DescT>
__global__ void desc_kernel(AsmSets sets, BandSys H, const unsigned long long* __restrict__ keys, int n_entries, DescT* __restrict__ desc) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_entries) return;
  const unsigned long long key = keys[e];
  const RowSetView& S = sets.s[(key >> 29) & 7u];
  const int res = static_cast<int>(key & 0x1fffffffu);
  int rowbase, rowend, colbase, colend;
  tile_ranges(H, static_cast<long long>(key >> 32), rowbase, rowend, colbase, colend);
  const int* pr = set_pos(S, res);
  const int ncols = set_ncols(S, res);
  DescT* d = desc + static_cast<size_t>(e) * kDescEntries;
  for (int c = 0; c < ncols; ++c) {
    const int p = pr[c];
    if (p >= rowbase && p < rowend) d[p - rowbase] = static_cast<DescT>(c);
    if (p >= colbase && p < colend && !(S.row_only_last && c == ncols - 1)) d[32 + p - colbase] = static_cast<DescT>(c);
  }
}

// ---- bulk async copies (TMA, 1-D) global -> shared with mbarrier completion
__device__ __forceinline__ unsigned g_smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void g_mbar_init(unsigned long long* b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(g_smem_u32(b)), "r"(count) : "memory"); }
__device__ __forceinline__ void g_mbar_expect_tx(unsigned long long* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(g_smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void g_tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(g_smem_u32(dst)), "l"(src), "r"(bytes), "r"(g_smem_u32(b)) : "memory");
}
__device__ __forceinline__ void g_mbar_wait(unsigned long long* b, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(g_smem_u32(b)), "r"(parity) : "memory");
}

// One warp's staging area: the Jacobian blocks ([rows x ncols | r | pad], 16-byte multiples) and the descriptors of one batch of entries land
// here by bulk async copies -- every load of the batch is in flight at once; the tensor-core fragments are then read straight out of it.
#ifndef LVI_STAGE_DOUBLES
#define LVI_STAGE_DOUBLES 2048
#endif
constexpr int kStageDoubles = LVI_STAGE_DOUBLES;   // 2048: 16 cam blocks of 112, 32 surfel blocks of 56, 32 camera-surfel blocks of 62
template <class DescT>
struct WarpStage {
  alignas(128) double J[kStageDoubles];
  alignas(16) DescT desc[32 * kDescEntries];
  unsigned short rowoff[kPanelRows];   // panel row -> offset of its Jacobian row in J
  unsigned short roff[kPanelRows];     //           -> offset of its residual value
  unsigned char rowent[kPanelRows];    //           -> entry slot of the batch (descriptor row)
  unsigned long long bar;
};

template <class DescT>
__global__ void __launch_bounds__(kGatherWarps * 32) gather_kernel(AsmSets sets, BandSys H, const unsigned long long* __restrict__ keys,
                                                                   const DescT* __restrict__ desc, const int* __restrict__ item_start, int n_items,
                                                                   int n_entries, double* __restrict__ g, int subtract) {
  // subtract == 0: tile = sum (rows)(rows)^T, g += rows^T r  (normal equations; tiles pre-zeroed where items share them)
  // subtract == 1: tile -= sum (rows)(rows)^T              (Schur complement on top of the damped system; no gradient)
  extern __shared__ __align__(128) unsigned char gather_smem[];
  constexpr unsigned FULL = 0xffffffffu;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr unsigned kNone = DescNone<DescT>::value;
  constexpr unsigned kDescBytes = kDescEntries * sizeof(DescT);
  WarpStage<DescT>& W = reinterpret_cast<WarpStage<DescT>*>(gather_smem)[warp];
  const int fr = lane >> 2, fk = lane & 3;   // fragment coordinates: A[row fr][k fk], B[k fk][col fr], C[row fr][cols 2 fk, 2 fk + 1]
  const long long n_band_tiles = static_cast<long long>(H.NT) * H.TPC;
  if (lane == 0) {
    g_mbar_init(&W.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  unsigned phase = 0;
  for (int item = blockIdx.x * kGatherWarps + warp; item < n_items; item += gridDim.x * kGatherWarps) {
    const int e0 = item_start[item], e1 = item_start[item + 1];
    const long long tile = static_cast<long long>(keys[e0] >> 32);
    const bool sole = (e0 == 0 || static_cast<long long>(keys[e0 - 1] >> 32) != tile) && (e1 == n_entries || static_cast<long long>(keys[e1] >> 32) != tile);
    int rowbase, rowend, colbase, colend, ld;
    tile_ranges(H, tile, rowbase, rowend, colbase, colend);
    double* dst;
    if (tile < n_band_tiles) { dst = H.tiles + (static_cast<size_t>(tile) << (2 * kTileLog)); ld = kTile; }
    else { dst = H.C + (rowbase - H.nb) + static_cast<size_t>(H.ldc) * (colbase - H.nb); ld = H.ldc; }
    const bool diag = rowbase == colbase;
    double acc[4][4][2];
#pragma unroll
    for (int mb = 0; mb < 4; ++mb)
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
    double gacc = 0.0;
    for (int eb = e0; eb < e1; eb += 32) {
      const int nk = min(32, e1 - eb);
      unsigned long long key = 0;
      if (lane < nk) key = keys[eb + lane];
      const int set = static_cast<int>((key >> 29) & 7u), res = static_cast<int>(key & 0x1fffffffu);
      const RowSetView& S = sets.s[set];
      const int rows = lane < nk ? S.rows : 0, blk = lane < nk ? set_blen(S, res) : 0;
      const int ncols = lane < nk ? set_ncols(S, res) : 0;
      int incl = rows, sincl = blk;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, o), t2 = __shfl_up_sync(FULL, sincl, o);
        if (lane >= o) { incl += t; sincl += t2; }
      }
      const int excl = incl - rows, sexcl = sincl - blk;
      int consumed = 0, base_rows = 0, base_stage = 0;
      while (consumed < nk) {
        const bool fits = lane >= consumed && lane < nk && (incl - base_rows) <= kPanelRows && (sincl - base_stage) <= kStageDoubles;
        const int cnt = __popc(__ballot_sync(FULL, fits));   // a contiguous run of lanes starting at `consumed`
        const int last = consumed + cnt - 1;
        const int total_rows = __shfl_sync(FULL, incl, last) - base_rows;
        const int total_stage = __shfl_sync(FULL, sincl, last) - base_stage;
        const int padded = (total_rows + 3) & ~3;
        // every load of the batch in flight at once: one bulk copy per Jacobian block, one for the descriptors
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the previous batch was read through the generic proxy
        if (lane == 0) g_mbar_expect_tx(&W.bar, static_cast<unsigned>(total_stage) * 8u + static_cast<unsigned>(cnt) * kDescBytes);
        __syncwarp();
        if (fits) {
          const int soff = sexcl - base_stage, row0 = excl - base_rows;
          g_tma_load_1d(W.J + soff, set_block(S, res), static_cast<unsigned>(blk) * 8u, &W.bar);
          for (int k = 0; k < rows; ++k) {
            W.rowoff[row0 + k] = static_cast<unsigned short>(soff + k * ncols);
            W.roff[row0 + k] = static_cast<unsigned short>(soff + rows * ncols + k);
            W.rowent[row0 + k] = static_cast<unsigned char>(lane - consumed);
          }
        }
        if (lane == 0) g_tma_load_1d(W.desc, desc + static_cast<size_t>(eb + consumed) * kDescEntries, static_cast<unsigned>(cnt) * kDescBytes, &W.bar);
        __syncwarp();
        g_mbar_wait(&W.bar, phase);
        phase ^= 1u;
        for (int k0 = 0; k0 < padded; k0 += 4) {
          const int row = k0 + fk;
          const bool live = row < total_rows;
          const int off = live ? W.rowoff[row] : 0;
          const DescT* dr = W.desc + (live ? W.rowent[row] : 0) * kDescEntries;
          double a[4], bf[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const unsigned cI = dr[8 * q + fr];
            a[q] = (live && cI != kNone) ? W.J[off + cI] : 0.0;
            if (diag) bf[q] = a[q];
            else { const unsigned cJ = dr[32 + 8 * q + fr]; bf[q] = (live && cJ != kNone) ? W.J[off + cJ] : 0.0; }
          }
#pragma unroll
          for (int mb = 0; mb < 4; ++mb)
#pragma unroll
            for (int nb = 0; nb < 4; ++nb)
              dmma_m8n8k4(acc[mb][nb], a[mb], bf[nb]);
        }
        if (diag && !subtract)
          for (int r = 0; r < total_rows; ++r) {
            const unsigned c = W.desc[W.rowent[r] * kDescEntries + lane];
            if (c != kNone) gacc = fma(W.J[W.rowoff[r] + c], W.J[W.roff[r]], gacc);
          }
        __syncwarp();   // everybody is done with the staging area before the next batch's copies land in it
        base_rows += total_rows; base_stage += total_stage;
        consumed += cnt;
      }
    }
#pragma unroll
    for (int mb = 0; mb < 4; ++mb)
#pragma unroll
      for (int nb = 0; nb < 4; ++nb)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          double* d = dst + (8 * mb + fr) + static_cast<size_t>(ld) * (8 * nb + 2 * fk + q);
          const double v = acc[mb][nb][q];
          if (subtract) {   // A holds its diagonal tiles as lower triangles: the mirrored half of the product is not applied
            if (v != 0.0 && !(diag && 8 * mb + fr < 8 * nb + 2 * fk + q)) { if (sole) *d -= v; else atomicAdd(d, -v); }
          }
          else if (sole) *d = v;
          else if (v != 0.0) atomicAdd(d, v);
        }
    if (diag && !subtract && gacc != 0.0) atomicAdd(g + rowbase + lane, gacc);
  }
}

// ---- host ---------------------------------------------------------------------------------------------------------------------------
template <int TYPE>
static void launch_pos(lvi_problem* p, AsmPlan& A) {
  const ResTable& T = p->view.tab[TYPE];
  if (T.n == 0 || !T.active) return;
  const int cols = rt_cols(TYPE), rows = rt_rows(TYPE), rstride = asm_block_doubles(TYPE);
  A.pos[TYPE].alloc(static_cast<size_t>(T.n) * cols);
  A.J[TYPE].alloc(static_cast<size_t>(T.n) * rstride);
  const long long n = static_cast<long long>(T.n) * cols;
  LVI_LAUNCH(p->ctx, pos_kernel<TYPE>, static_cast<int>((n + 255) / 256), 256, 0, p->view, A.pos[TYPE].p);
  A.sets.s[TYPE] = RowSetView{A.J[TYPE].p, A.pos[TYPE].p, cols, rows, rstride, T.lo, T.hi, nullptr, nullptr, 0};
}

// keys, work items and descriptors of a plan whose row sets (positions) are in place
static void build_plan_structure(lvi_problem* p, AsmPlan& A, const BandSys& H, int desc_width) {
  lvi_ctx* ctx = p->ctx;
  cudaStream_t st = ctx->stream;
  const TileDims D{H.nb, H.NT, H.T, H.RB, H.TPC};
  int n_res = 0, base[RT_COUNT];
  for (int t = 0; t < RT_COUNT; ++t) { base[t] = n_res; n_res += A.sets.s[t].hi - A.sets.s[t].lo; }
  A.built = true; A.n_items = 0; A.n_entries = 0;
  if (n_res == 0) return;
  DBuf<int> counts(static_cast<size_t>(n_res) + 1), offsets(static_cast<size_t>(n_res) + 1);
  counts.zero(st);
  LVI_CUDA(cudaMemsetAsync(p->fail.p + 2, 0, sizeof(int), st));
  for (int t = 0; t < RT_COUNT; ++t) {
    const RowSetView& S = A.sets.s[t];
    if (S.hi > S.lo) LVI_LAUNCH(ctx, pairs_kernel<false>, (S.hi - S.lo + 127) / 128, 128, 0, S, t, D, base[t], nullptr, counts.p, nullptr, p->fail.p + 2);
  }
  size_t tb = 0, tb2 = 0, tb3 = 0, tb4 = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, counts.p, offsets.p, n_res + 1, st);
  DBuf<unsigned char> tmp(std::max<size_t>(tb, 1));
  LVI_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, counts.p, offsets.p, n_res + 1, st));
  int n_entries = 0;
  LVI_CUDA(cudaMemcpyAsync(&n_entries, offsets.p + n_res, sizeof(int), cudaMemcpyDeviceToHost, st));
  LVI_CUDA(cudaStreamSynchronize(st));
  if (n_entries == 0) return;
  DBuf<unsigned long long> unsorted(n_entries);
  A.keys.alloc(n_entries);
  for (int t = 0; t < RT_COUNT; ++t) {
    const RowSetView& S = A.sets.s[t];
    if (S.hi > S.lo) LVI_LAUNCH(ctx, pairs_kernel<true>, (S.hi - S.lo + 127) / 128, 128, 0, S, t, D, base[t], offsets.p, nullptr, unsorted.p, p->fail.p + 2);
  }
  cub::DeviceRadixSort::SortKeys(nullptr, tb2, unsorted.p, A.keys.p, n_entries, 0, 64, st);
  DBuf<unsigned char> tmp2(std::max<size_t>(tb2, 1));
  LVI_CUDA(cub::DeviceRadixSort::SortKeys(tmp2.p, tb2, unsorted.p, A.keys.p, n_entries, 0, 64, st));
  // work items: every tile's run of keys cut into pieces of <= kChunk
  DBuf<int> head_idx(n_entries), seg_start(n_entries);
  DBuf<unsigned char> flag(n_entries);
  LVI_LAUNCH(ctx, seg_head_kernel, (n_entries + 255) / 256, 256, 0, A.keys.p, n_entries, head_idx.p);
  cub::DeviceScan::InclusiveScan(nullptr, tb3, head_idx.p, seg_start.p, cub::Max(), n_entries, st);
  DBuf<unsigned char> tmp3(std::max<size_t>(tb3, 1));
  LVI_CUDA(cub::DeviceScan::InclusiveScan(tmp3.p, tb3, head_idx.p, seg_start.p, cub::Max(), n_entries, st));
  LVI_LAUNCH(ctx, item_flag_kernel, (n_entries + 255) / 256, 256, 0, seg_start.p, n_entries, flag.p);
  A.item_start.alloc(static_cast<size_t>(n_entries) + 1);
  DBuf<int> n_sel(1);
  cub::CountingInputIterator<int> iota(0);
  cub::DeviceSelect::Flagged(nullptr, tb4, iota, flag.p, A.item_start.p, n_sel.p, n_entries, st);
  DBuf<unsigned char> tmp4(std::max<size_t>(tb4, 1));
  LVI_CUDA(cub::DeviceSelect::Flagged(tmp4.p, tb4, iota, flag.p, A.item_start.p, n_sel.p, n_entries, st));
  int n_items = 0, fail = 0;
  LVI_CUDA(cudaMemcpyAsync(&n_items, n_sel.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  LVI_CUDA(cudaMemcpyAsync(&fail, p->fail.p + 2, sizeof(int), cudaMemcpyDeviceToHost, st));
  LVI_CUDA(cudaStreamSynchronize(st));
  LVI_REQUIRE(fail == 0, LVI_ERR_INVALID, fail == 4 ? "assembly plan: a row couples positions further apart than the half bandwidth" : "assembly plan: a row touches more than 48 tile rows");
  LVI_CUDA(cudaMemcpyAsync(A.item_start.p + n_items, &n_entries, sizeof(int), cudaMemcpyHostToDevice, st));
  LVI_CUDA(cudaStreamSynchronize(st));
  A.n_items = n_items; A.n_entries = n_entries;
  A.desc_width = desc_width;
  A.desc.alloc(static_cast<size_t>(n_entries) * kDescEntries * desc_width);
  LVI_CUDA(cudaMemsetAsync(A.desc.p, 0xFF, A.desc.n, st));
  if (desc_width == 1) LVI_LAUNCH(ctx, desc_kernel<unsigned char>, (n_entries + 127) / 128, 128, 0, A.sets, H, A.keys.p, n_entries, A.desc.p);
  else LVI_LAUNCH(ctx, desc_kernel<unsigned short>, (n_entries + 127) / 128, 128, 0, A.sets, H, A.keys.p, n_entries, reinterpret_cast<unsigned short*>(A.desc.p));
}

void assemble_build_plan(lvi_problem* p) {
  AsmPlan& A = p->asmp;
  if (A.built) return;
  for (int t = 0; t < RT_COUNT; ++t) A.sets.s[t] = RowSetView{nullptr, nullptr, rt_cols(t), rt_rows(t), asm_block_doubles(t), 0, 0, nullptr, nullptr, 0};
  launch_pos<RT_GYRO>(p, A); launch_pos<RT_ACCEL>(p, A); launch_pos<RT_SURFEL>(p, A); launch_pos<RT_CAM>(p, A); launch_pos<RT_CAMSURF>(p, A); launch_pos<RT_ORIENT>(p, A);
  build_plan_structure(p, A, p->H, 1);
}

// ---- the Schur complement on the inverse depths through the same gather: one "residual" per free inverse depth, its row = the merged
// coupling row of the landmark (distinct positions, SchurView::urow_pos) + one more column at the position of the right-hand-side row
__global__ void schur_set_kernel(SchurView SV, int rhs_pos, int* __restrict__ boff, int* __restrict__ bcols, int* __restrict__ pos, int* __restrict__ max_cols) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= SV.n_rho) return;
  const int lm = SV.lm_of_rho[k];
  const int rs = SV.row_start[lm], lu = SV.ulen[lm];
  const int off = rs + 2 * k;   // even; blocks of lu + 1 doubles (rounded up to even) never overlap: lu <= slots of the landmark
  boff[k] = off; bcols[k] = lu + 1;
  for (int i = 0; i < lu; ++i) pos[off + i] = SV.urow_pos[rs + i];
  pos[off + lu] = rhs_pos;
  atomicMax(max_cols, lu + 1);
}

void assemble_build_schur_plan(lvi_problem* p) {
  AsmPlan& A = p->schur_plan;
  p->schur_gather = false;
  if (A.built) return;
  A.built = true;
  const int n_rho = p->L.n_rho;
  // Measured at C2 (gpurun_out/r2w): 0.31 ms through the gather against 0.28 ms for schur_eliminate_kernel's atomics -- a landmark's row has
  // ~300 columns spread over ~24 tile rows (its observation windows are 10 knots apart), so 888 k (tile, landmark) keys each stage a 2.6 KB
  // row for <= 64 useful columns.  Kept as an option (LVI_SCHUR_GATHER=1: no atomics, bitwise reproducible A); the default is the atomic kernel.
  if (n_rho == 0 || !std::getenv("LVI_SCHUR_GATHER")) return;
  lvi_ctx* ctx = p->ctx;
  cudaStream_t st = ctx->stream;
  const size_t n_blocks = p->L.row_pos.size() + 2 * static_cast<size_t>(n_rho) + 2;
  p->schur_boff.alloc(n_rho); p->schur_bcols.alloc(n_rho); p->schur_pos.alloc(n_blocks); p->schur_rows.alloc(n_blocks);
  DBuf<int> mx(1);
  mx.zero(st);
  LVI_LAUNCH(ctx, schur_set_kernel, (n_rho + 127) / 128, 128, 0, p->schur, p->H.nb + p->H.nbo, p->schur_boff.p, p->schur_bcols.p, p->schur_pos.p, mx.p);
  int max_cols = 0;
  LVI_CUDA(cudaMemcpyAsync(&max_cols, mx.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  LVI_CUDA(cudaStreamSynchronize(st));
  if (max_cols + 2 > kStageDoubles) return;   // a row has to fit the staging area: longer rows keep the atomic kernel
  for (int t = 0; t < RT_COUNT; ++t) A.sets.s[t] = RowSetView{nullptr, nullptr, 0, 1, 0, 0, 0, nullptr, nullptr, 0};
  A.sets.s[0] = RowSetView{p->schur_rows.p, p->schur_pos.p, 0, 1, 0, 0, n_rho, p->schur_boff.p, p->schur_bcols.p, 1};
  build_plan_structure(p, A, p->A, 2);
  p->schur_gather = A.n_items > 0;
}

static void launch_gather(lvi_problem* p, AsmPlan& A, const BandSys& target, double* g, int subtract) {
  static const int per_sm = std::getenv("LVI_GATHER_CTAS") ? std::atoi(std::getenv("LVI_GATHER_CTAS")) : 3;
  const int grid = std::min((A.n_items + kGatherWarps - 1) / kGatherWarps, p->ctx->sm_count * per_sm);
  if (A.desc_width == 1) {
    constexpr size_t smem = sizeof(WarpStage<unsigned char>) * kGatherWarps;
    if (!p->ctx->ks.gather_attr[0]) {
      LVI_CUDA(cudaFuncSetAttribute(gather_kernel<unsigned char>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      p->ctx->ks.gather_attr[0] = true;
    }
    LVI_LAUNCH_AS(p->ctx, subtract ? "gather_kernel(schur)" : "gather_kernel", gather_kernel<unsigned char>, grid, kGatherWarps * 32, smem, A.sets, target, A.keys.p, A.desc.p,
                  A.item_start.p, A.n_items, A.n_entries, g, subtract);
  } else {
    constexpr size_t smem = sizeof(WarpStage<unsigned short>) * kGatherWarps;
    if (!p->ctx->ks.gather_attr[1]) {
      LVI_CUDA(cudaFuncSetAttribute(gather_kernel<unsigned short>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      p->ctx->ks.gather_attr[1] = true;
    }
    LVI_LAUNCH_AS(p->ctx, subtract ? "gather_kernel(schur)" : "gather_kernel", gather_kernel<unsigned short>, grid, kGatherWarps * 32, smem, A.sets, target, A.keys.p,
                  reinterpret_cast<const unsigned short*>(A.desc.p), A.item_start.p, A.n_items, A.n_entries, g, subtract);
  }
}
void assemble_schur_gather(lvi_problem* p) {
  if (p->schur_plan.n_items) launch_gather(p, p->schur_plan, p->A, nullptr, 1);
}

__global__ void mark_tiles_kernel(const unsigned long long* __restrict__ keys, int n, long long n_band_tiles, unsigned char* __restrict__ flags) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const long long tile = static_cast<long long>(keys[e] >> 32);
  if (tile < n_band_tiles) flags[tile] = 1;
}
void assemble_mark_tiles(lvi_problem* p, unsigned char* flags_d) {
  AsmPlan& A = p->asmp;
  if (A.n_entries == 0) return;
  LVI_LAUNCH(p->ctx, mark_tiles_kernel, (A.n_entries + 255) / 256, 256, 0, A.keys.p, A.n_entries, static_cast<long long>(p->H.NT) * p->H.TPC, flags_d);
}

// tile += gathered J^T J, g += J^T r  (H, g zeroed by the caller; the Jacobian rows are in the plan's buffers)
void assemble_gather(lvi_problem* p) {
  if (p->asmp.n_items) launch_gather(p, p->asmp, p->H_lin, p->g_lin, 0);
}

}  // namespace lvi
