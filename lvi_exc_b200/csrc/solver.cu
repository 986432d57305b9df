// solver.cu — on-device Levenberg-Marquardt with an exact band+arrow Cholesky solve.
//
// Replaces ceres::Solve as configured by kontiki::TrajectoryEstimator::Solve (K/trajectory_estimator.h:38-68):
// TRUST_REGION / LEVENBERG_MARQUARDT / SPARSE_SCHUR (an exact solve), Jacobi scaling, Ceres (<= 2.1) default
// tolerances.  Ceres is not vendored; the loop below restates TrustRegionMinimizer +
// LevenbergMarquardtStrategy step by step and is kept line-for-line comparable with the CPU oracle (oracle/orc_solve.cpp).
//
// Linear algebra: (S H S + D^2) y = -S g with H in 32x32 tile storage (problem.cuh), inverse depths eliminated first (Schur).
//   band_factor_ll_kernel     left-looking tile Cholesky without a grid-wide barrier.  One persistent CTA per pivot chain runs the serial
//                             recurrence D_j -> W_j = L_jj^-1 -> X_j = L(j+1,j) -> D_{j+1} out of its shared memory (tile_chol.cuh: two
//                             factoring warps + an inverse follower); every other tile of the factor (band + arrow border + the rhs row, so
//                             the forward substitution comes for free) is one task of a worker CTA: accumulator in registers, older
//                             source tiles staged by TMA bulk copies behind ld.acquire ready flags, the freshest pair and W_j taken
//                             from flagged ("LL") copies, X = P W^T, each tile written once.
//   corner_solve_kernel       dense Cholesky of the (<= ~100)^2 Schur complement of the arrow border + its triangular solves.
//   band_backsolve_ll_kernel  backward substitution: one persistent CTA per chain on TMA-staged tiles, far products and the border part
//                             by one worker task per block column, x_m / u_m handed over as flagged vectors.
// All fp64: the normal matrix of a 0.02 s-knot spline is too ill-conditioned for fp32/bf16 factors.
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>

#include "nccl_dyn.hpp"
#include "problem.cuh"
#include "tile_chol.cuh"

namespace lvi {

constexpr unsigned FULL = kFullWarp;
constexpr int kLP = kCholLd;  // padded leading dimension of the 32x32 shared-memory tiles

// ---- small elementwise kernels -------------------------------------------------------------------------------------
__global__ void diag_kernel(BandSys H, SchurView SV, int nt, double* __restrict__ d) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nt) d[t] = t < SV.base ? *band_addr(H, t, t) : SV.Hrr[t - SV.base];
}
__global__ void scale_init_kernel(const double* __restrict__ dH, int nt, int jacobi, double* __restrict__ scale) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nt) scale[t] = jacobi ? 1.0 / (1.0 + sqrt(dH[t])) : 1.0;  // ceres: 1 / (1 + ||J_col||)
}
__global__ void lm_diag_kernel(const double* __restrict__ dH, const double* __restrict__ scale, int nt, double lo, double hi, double* __restrict__ diag) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nt) diag[t] = fmin(fmax(dH[t] * scale[t] * scale[t], lo), hi);  // LevenbergMarquardtStrategy: clamp(diag(J^T J))
}

// A = S H S + diag(D2) in tile storage; the extra border row carries rhs = -S g.  One CTA per tile (+ corner CTAs): 228 MB of pure
// streaming at C2, so each thread moves 4 consecutive rows of one column with two 16-byte loads/stores and reads its scale factors once.
__global__ void __launch_bounds__(256) build_system_kernel(BandSys H, BandSys A, const double* __restrict__ scale, const double* __restrict__ diag,
                                                           double inv_radius, const double* __restrict__ g) {
  static_assert(kTile == 32, "thread mapping below assumes 32x32 tiles");
  const int ntile = A.NT * A.TPC;
  const int nb = A.nb, nbo = A.nbo;
  if (static_cast<int>(blockIdx.x) < ntile) {
    const int J = blockIdx.x / A.TPC, q = blockIdx.x % A.TPC;
    if (q <= A.T && J + q >= A.NT) return;
    const int b = threadIdx.x >> 3, a0 = (threadIdx.x & 7) * 4;   // column b, rows a0..a0+3
    const size_t off = static_cast<size_t>(blockIdx.x) * kTileElems + b * kTile + a0;
    const double2 s0 = __ldcs(reinterpret_cast<const double2*>(H.tiles + off));
    const double2 s1 = __ldcs(reinterpret_cast<const double2*>(H.tiles + off) + 1);
    const double src[4] = {s0.x, s0.y, s1.x, s1.y};
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    const int j = J * kTile + b;
    if (q <= A.T) {
      const int i0 = (J + q) * kTile + a0;
      const double sj = j < nb ? scale[j] : 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i0 + k;
        if (i < nb && j < nb) {
          if (i >= j) v[k] = src[k] * scale[i] * sj;
          if (i == j) v[k] += diag[i] * inv_radius;
        } else if (i == j) v[k] = 1.0;  // padding rows keep the factorisation well defined
      }
    } else if (j < nb) {
      const int bi0 = (q - A.T - 1) * kTile + a0;
      const double sj = scale[j];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int bi = bi0 + k;
        if (bi < nbo) v[k] = src[k] * scale[nb + bi] * sj;
        else if (bi == nbo) v[k] = -g[j] * sj;
      }
    }
    double2* dst = reinterpret_cast<double2*>(A.tiles + off);
    __stcs(dst, make_double2(v[0], v[1]));
    __stcs(dst + 1, make_double2(v[2], v[3]));
  } else {
    const int ldc = A.ldc;
    for (int e = (blockIdx.x - ntile) * blockDim.x + threadIdx.x; e < ldc * ldc; e += (gridDim.x - ntile) * blockDim.x) {
      const int bi = e % ldc, bj = e / ldc;
      double v = 0.0;
      if (bi < nbo && bj < nbo) {
        if (bi >= bj) v = H.C[e] * scale[nb + bi] * scale[nb + bj];
        if (bi == bj) v += diag[nb + bi] * inv_radius;
      } else if (bi == nbo && bj < nbo) v = -g[nb + bj] * scale[nb + bj];
      else if (bi == bj) v = 1.0;
      A.C[e] = v;
    }
  }
}

// ---- Schur complement on the inverse depths (1x1 diagonal blocks, eliminated first like Ceres' SPARSE_SCHUR e-blocks) ----------
// One CTA per free inverse depth r:  A_xx -= h h^T / d,  rhs_x -= h b_r / d   with h = S_x H_xr s_r (scaled coupling row),
// d = s_r^2 H_rr + D_r^2, b_r = -s_r g_r.  Rows are short (<= 30 + 24 x track length), updates go to the tile store with fp64 atomics.
__global__ void __launch_bounds__(128) schur_eliminate_kernel(BandSys A, SchurView SV, const double* __restrict__ scale, const double* __restrict__ diag,
                                                              double inv_radius, const double* __restrict__ g) {
  extern __shared__ double sm[];
  const int k = blockIdx.x;
  const int lm = SV.lm_of_rho[k];
  const int rs = SV.row_start[lm], len = SV.row_start[lm + 1] - rs;
  if (len == 0) return;
  const int lu = SV.ulen[lm];   // distinct positions of the row (the windows of consecutive observations overlap), ascending
  double* hv = sm;
  int* hp = reinterpret_cast<int*>(sm + len);
  const int t = SV.base + k;
  const double sr = scale[t];
  const double d = SV.Hrr[k] * sr * sr + diag[t] * inv_radius;
  const double dinv = 1.0 / d;
  const double br = -g[t] * sr;
  for (int i = threadIdx.x; i < lu; i += blockDim.x) { hv[i] = 0.0; hp[i] = SV.urow_pos[rs + i]; }
  __syncthreads();
  for (int i = threadIdx.x; i < len; i += blockDim.x) {   // slots of one parameter add up: h is the coupling to the PARAMETER
    const int p = SV.row_pos[rs + i];
    if (p < 0) continue;
    const double v = SV.Hrx[rs + i] * scale[p] * sr;
    if (v != 0.0) atomicAdd(&hv[SV.slot2u[rs + i]], v);
  }
  __syncthreads();
  const int rhs_row = A.nb + A.nbo;
  for (int i = threadIdx.x; i < lu; i += blockDim.x)
    if (hv[i] != 0.0) atomicAdd(band_addr(A, rhs_row, hp[i]), -hv[i] * br * dinv);
  // rows i and lu-1-i together hold lu+1 pairs: every thread gets the same amount of work and no index decoding is needed
  for (int r = threadIdx.x; 2 * r < lu; r += blockDim.x) {
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      const int i = half == 0 ? r : lu - 1 - r;
      if (half == 1 && i == r) break;
      const int pi = hp[i];
      const double hi = hv[i] * dinv;
      if (hi == 0.0) continue;
      for (int j = 0; j <= i; ++j) {
        const double v = hi * hv[j];
        if (v != 0.0) atomicAdd(band_addr(A, pi, hp[j]), -v);   // positions ascend: pi >= hp[j]
      }
    }
  }
}

// The same elimination through the tile gather (assemble.cu): this kernel only writes, per free inverse depth, the block
//   sqrt(1/d) * [ merged scaled coupling row h | b_r ]
// and gather_kernel subtracts (block)(block)^T from the tiles of A -- the column at the position of the right-hand-side row turns the
// rhs update into one more row of the same product.  No fp64 atomics on the band.
__global__ void __launch_bounds__(128) schur_rows_kernel(BandSys A, SchurView SV, const double* __restrict__ scale, const double* __restrict__ diag,
                                                         double inv_radius, const double* __restrict__ g, const int* __restrict__ boff, double* __restrict__ rows) {
  extern __shared__ double sm[];
  const int k = blockIdx.x;
  const int lm = SV.lm_of_rho[k];
  const int rs = SV.row_start[lm], len = SV.row_start[lm + 1] - rs;
  const int lu = SV.ulen[lm];
  double* out = rows + boff[k];
  double* hv = sm;
  const int t = SV.base + k;
  const double sr = scale[t];
  const double d = SV.Hrr[k] * sr * sr + diag[t] * inv_radius;
  const double sq = sqrt(1.0 / d);
  for (int i = threadIdx.x; i < lu; i += blockDim.x) hv[i] = 0.0;
  __syncthreads();
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    const int p = SV.row_pos[rs + i];
    if (p < 0) continue;
    const double v = SV.Hrx[rs + i] * scale[p] * sr;
    if (v != 0.0) atomicAdd(&hv[SV.slot2u[rs + i]], v);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < lu; i += blockDim.x) out[i] = hv[i] * sq;
  if (threadIdx.x == 0) { out[lu] = -g[t] * sr * sq; if (((lu + 1) & 1) != 0) out[lu + 1] = 0.0; }
}

// y_r = (b_r - h . x) / d  after the reduced system has been solved
__global__ void __launch_bounds__(128) schur_back_kernel(BandSys A, SchurView SV, const double* __restrict__ scale, const double* __restrict__ diag,
                                                         double inv_radius, const double* __restrict__ g) {
  const int k = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (k >= SV.n_rho) return;
  const int lm = SV.lm_of_rho[k];
  const int rs = SV.row_start[lm], len = SV.row_start[lm + 1] - rs;
  const int t = SV.base + k;
  const double sr = scale[t];
  const double d = SV.Hrr[k] * sr * sr + diag[t] * inv_radius;
  double acc = 0.0;
  for (int i = lane; i < len; i += 32) {
    const int p = SV.row_pos[rs + i];
    if (p < 0) continue;
    const double xv = p < A.nb ? A.x[p] : A.x[static_cast<size_t>(A.NT) * kTile + (p - A.nb)];
    acc += SV.Hrx[rs + i] * scale[p] * sr * xv;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
  if (lane == 0) SV.yrho[k] = (-g[t] * sr - acc) / d;
}

// ---- primitives of the flag-driven tile kernels
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// (the __syncwarp matters: without it the stamping lane can stay diverged from its warp through the code that follows, which made
// the traced factorisation of the diagonal block 3x slower than the untraced one)
#define LVI_TRACE(slot) do { if (S.trace) { if (tid == 0) S.trace[static_cast<size_t>(tq) * 8 + (slot)] = gtime(); __syncwarp(); } } while (0)
#define LVI_TRACE_AT(row, slot) do { if (S.trace) { if (tid == 0) S.trace[static_cast<size_t>(row) * 8 + (slot)] = gtime(); __syncwarp(); } } while (0)
// Tasks on the pivot chain (diagonal tile and first sub-diagonal tile) poll tightly; every other task backs off between polls so that the
// CTAs that merely wait do not steal issue slots and L2 bandwidth from the ones that work (two CTAs share an SM).
__device__ __forceinline__ void spin_until_set(const int* f, bool critical = true) {
  if (critical) { while (ld_acquire(f) == 0) {} }
  else { while (ld_acquire(f) == 0) __nanosleep(400); }
}

// issue order of the block columns: the two chains of the two-sided ordering are interleaved so that both advance together
__device__ __forceinline__ int ordered_column(int p, int NT0, int NT) {
  const int n0 = NT0, n1 = NT - NT0, mn = min(n0, n1);
  if (p < 2 * mn) return (p & 1) ? NT0 + (p >> 1) : (p >> 1);
  const int r = p - 2 * mn;
  return n0 >= n1 ? mn + r : NT0 + mn + r;
}
// same for the back substitution, which walks each chain from its end
__device__ __forceinline__ int ordered_column_desc(int p, int NT0, int NT) {
  const int c = ordered_column(p, NT0, NT);
  return c < NT0 ? NT0 - 1 - c : NT - 1 - (c - NT0);
}

// ---- flagged ("LL") transfers for the hand-offs on the serial chains
// The contribution of column m to column m-1 is the serial chain of the substitution.  It travels through a flagged mailbox (each 8-byte
// word = 32 bits of the value | a flag, in the style of NCCL's LL protocol): the consumer sees data and readiness in ONE L2 round trip
// instead of the three of "atomicAdd partial sum, fence, bump arrival counter / poll counter, load partial sum".
__device__ __forceinline__ void ll_store(unsigned long long* slot, double v, unsigned flag = 1u) {
  const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v));
  const unsigned long long w0 = (b << 32) | flag, w1 = (b & 0xffffffff00000000ull) | flag;
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ bool ll_valid(unsigned long long w0, unsigned long long w1, unsigned flag) {
  return static_cast<unsigned>(w0) == flag && static_cast<unsigned>(w1) == flag;   // each 8-byte half carries its own flag: no reliance on 16-byte atomicity
}
__device__ __forceinline__ double ll_value(unsigned long long w0, unsigned long long w1) {
  return __longlong_as_double(static_cast<long long>((w0 >> 32) | (w1 & 0xffffffff00000000ull)));
}
__device__ __forceinline__ double ll_load(const unsigned long long* slot, unsigned flag = 1u) {
  unsigned long long w0, w1;
  do {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot) : "memory");
  } while (!ll_valid(w0, w1, flag));
  return ll_value(w0, w1);
}
// N flagged words with all loads in flight at once (a late word costs ONE more round trip, not one per word)
template <int N>
__device__ __forceinline__ void ll_load_n(const unsigned long long* slot, int stride_words, double (&v)[N], unsigned flag) {
  unsigned long long w0[N], w1[N];
  bool ok;
  do {
#pragma unroll
    for (int i = 0; i < N; ++i)
      asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0[i]), "=l"(w1[i]) : "l"(slot + static_cast<size_t>(i) * stride_words) : "memory");
    ok = true;
#pragma unroll
    for (int i = 0; i < N; ++i) ok = ok && ll_valid(w0[i], w1[i], flag);
  } while (!ok);
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = ll_value(w0[i], w1[i]);
}

constexpr int kFacThreads = 256;
static_assert(kTile == 32, "band_factor_ll_kernel is written for 32x32 tiles (4 outputs per thread)");

// Flagged copies ("LL": every 8-byte word = half of a value | a 32-bit flag) of every tile, of W_j and of the pre-accumulated first
// sub-diagonal tile.  A consumer that needs a tile the moment it is made polls the data itself: one L2 round trip instead of "store,
// fence, flag / poll flag, load" (measured, tools/bench_handoff.cu: 0.76 us same die / 1.19 us across dies against 1.9 - 2.3 us for a
// 32x32 tile).  The flag is a per-factorisation epoch, so the copies never need clearing.
__device__ __forceinline__ unsigned long long* ll_tile(const BandSys& S, int tq) { return S.ll + static_cast<size_t>(tq) * 2 * kTileElems; }
__device__ __forceinline__ unsigned long long* ll_W(const BandSys& S, int j) { return ll_tile(S, S.NT * S.TPC + j); }
__device__ __forceinline__ unsigned long long* ll_P(const BandSys& S, int j) { return ll_tile(S, S.NT * S.TPC + S.NT + j); }

// CTA-wide fetch of one flagged tile into shared memory (element e of the tile -> dst[e], optionally a second copy).  `eager` consumers
// sit next to the pivot chain and poll the whole tile; the others first wait, with back-off, for one word, so that two dozen waiting CTAs
// per chain do not spend L2 bandwidth on polling.
__device__ __forceinline__ void fetch_ll_tile(const unsigned long long* src, unsigned flag, bool eager, double* dst, double* dst2, double* plain = nullptr) {
  const int tid = threadIdx.x;
  if (!eager) {
    if (tid == 0) {
      unsigned long long w0, w1;
      while (true) {
        asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src) : "memory");
        if (ll_valid(w0, w1, flag)) break;
        __nanosleep(200);
      }
    }
    __syncthreads();
  }
  double v[4];
  ll_load_n<4>(src + 2 * tid, 2 * kFacThreads, v, flag);
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) dst[tid + kFacThreads * q4] = v[q4];
  if (dst2) {
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) dst2[tid + kFacThreads * q4] = v[q4];
  }
  if (plain) {   // the tile's plain copy in the tile store (only the back-substitution kernel reads it)
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) plain[tid + kFacThreads * q4] = v[q4];
  }
}

// the warps that idle during the diagonal block's Cholesky fetch flagged tiles into shared memory, all their loads in flight at once;
// the non-blocking form gives up as soon as the Cholesky has finished (its barrier must not wait for a tile that is still being made)
template <bool BLOCKING, int NTHREADS, int LD = 32>
__device__ __forceinline__ bool helper_fetch_tile(const unsigned long long* src, unsigned flag, double* dst, int ht, volatile int* progress) {
  constexpr int kPer = kTileElems / NTHREADS;
  static_assert(kPer * NTHREADS == kTileElems, "helper count must divide the tile");
  unsigned long long w0[kPer], w1[kPer];
  while (true) {
#pragma unroll
    for (int q = 0; q < kPer; ++q)
      asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0[q]), "=l"(w1[q]) : "l"(src + 2 * (ht + NTHREADS * q)) : "memory");
    bool ok = true;
#pragma unroll
    for (int q = 0; q < kPer; ++q) ok = ok && ll_valid(w0[q], w1[q], flag);
    if (ok) break;
    if (!BLOCKING && *progress >= 32) return false;
    __nanosleep(100);   // the Cholesky warps share this SM's load/store pipe: do not hammer it while waiting
  }
#pragma unroll
  for (int q = 0; q < kPer; ++q) {
    const int e = ht + NTHREADS * q;
    dst[LD == 32 ? e : (e & 31) + LD * (e >> 5)] = ll_value(w0[q], w1[q]);
  }
  return true;
}

// The chain CTA's products run on the fp64 tensor pipe (mma.sync.m8n8k4.f64): lane l of a warp holds element (l / 4, l % 4) of an 8x4
// operand fragment.  With a leading dimension of kXL = 36 doubles the 16 lanes of a half-warp (4 rows x 4 k) fall into 16 distinct
// 8-byte bank pairs ((row + 36 k) mod 16 = row + 4 k), so every fragment load is conflict-free; 32 or 33 would serialise them 4-way.
constexpr int kXL = 36;
#ifndef LVI_CHAIN_DMMA
#define LVI_CHAIN_DMMA 1
#endif
constexpr bool kChainDmma = LVI_CHAIN_DMMA != 0;

// Who writes the PLAIN copies of the chain's own tiles X_j = L(j+1,j) and W_j (read by the back-substitution kernel only).  Once the workers
// had slack again (2x4 updates), the 32 KB of stores per column at the end of the chain CTA's column were 0.7 us of a 5.6 us period.  The
// first consumer among the workers has the values in registers anyway: the task of tile (j+2, j) fetches W_j, and the task of tile
// (j+3, j+1) fetches X_j for its last update -- they write the copies.  Where those tasks do not exist (end of a chain, narrow bands) the
// chain CTA still does.
__device__ __forceinline__ bool worker_writes_W(const BandSys& S, int j, int c_end) { return kChainDmma && S.T >= 2 && j + 2 < c_end; }
__device__ __forceinline__ bool worker_writes_X(const BandSys& S, int j, int c_end) { return kChainDmma && S.T >= 3 && j + 3 < c_end; }
constexpr int kBufLd = 32 * kXL;   // one staging buffer: a 32x32 tile (8 KB, TMA destination) or a padded 32x33 / 32x36 tile
#ifndef LVI_FAC_STAGES
#define LVI_FAC_STAGES 4
#endif
#ifndef LVI_FAC_PAIR
#define LVI_FAC_PAIR 2
#endif
constexpr int kPair = LVI_FAC_PAIR;
static_assert(LVI_FAC_PAIR % 2 == 0, "the two halves of a worker CTA take alternate updates of a pair");       // rank-32 updates per CTA barrier in the workers (kStages >= kPair + 2): 1 -> 2 with 4 stages 1.798 -> 1.782 ms
constexpr int kStages = LVI_FAC_STAGES;   // staging depth of the workers' tile pipeline (each stage: two 8 KB tiles)
struct FacShared {
  double buf[2 * kStages][kBufLd];   // workers: kStages x (A tile, B tile) filled by bulk async copies; chain CTA: sA, sB, sD, sM
  double sW[32 * kXL], sR[32];
  unsigned long long full[kStages];  // mbarriers: "stage filled" (one arrival + the copies' bytes)
  int q;
  volatile int progress;
  int leave;
  int nready;
};

// ---- bulk async copies (TMA, 1-D) global -> shared with mbarrier completion
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(b))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
// warp-uniform test: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(unsigned long long* b, unsigned parity) {
  unsigned ok;
  asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  return __shfl_sync(0xffffffffu, ok, 0) != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}

// acc(2x2 block of rows 2rp.., columns 2cp..) -= A(rows, :) B(cols, :)^T over one 32-wide k-step, A in sA[r + 32 m], B in sB[c + 32 m].
// 2x2 register blocking with 16-byte shared-memory loads: 3 shared-memory wavefronts per 4 DFMA instead of the 6 of a 1x4 blocking -- the
// shared-memory pipe, not the FP64 pipe, bounds this loop (measured: 0.75 us -> per 32^3 step and SM with 1x4).
__device__ __forceinline__ void rank32_update_2x2(const double* sA, const double* sB, int rp, int cp, double (&acc)[4]) {
#pragma unroll 8
  for (int m = 0; m < 32; ++m) {
    const double2 xa = *reinterpret_cast<const double2*>(sA + 2 * rp + 32 * m);
    const double2 xb = *reinterpret_cast<const double2*>(sB + 2 * cp + 32 * m);
    acc[0] = fma(-xa.x, xb.x, acc[0]);
    acc[1] = fma(-xa.y, xb.x, acc[1]);
    acc[2] = fma(-xa.x, xb.y, acc[2]);
    acc[3] = fma(-xa.y, xb.y, acc[3]);
  }
}
// The workers' form: 2x4 register blocks on HALF the CTA (128 threads: rows 2rq.., columns 4cq..) over the k-range [m0, m1): 4 shared-memory
// wavefronts per 8 DFMA instead of 3 per 4.  Two CTAs per SM made the 2x2 form shared-memory bound at the SM level (2 x 768 wavefronts =
// 1536 cycles per update against 2 x 512 cycles of fp64 issue; measured 1.5 us per update); the two halves of the CTA take alternate
// updates of a pair (or the two k-halves of one update) into PARTIAL accumulators that are added once, at the end of the task.
__device__ __forceinline__ void rank32_update_2x4(const double* sA, const double* sB, int rq, int cq, double (&p8)[8], int m0, int m1) {
#pragma unroll 8
  for (int m = m0; m < m1; ++m) {
    const double2 xa = *reinterpret_cast<const double2*>(sA + 2 * rq + 32 * m);
    const double2 x01 = *reinterpret_cast<const double2*>(sB + 4 * cq + 32 * m);
    const double2 x23 = *reinterpret_cast<const double2*>(sB + 4 * cq + 2 + 32 * m);
    p8[0] = fma(-xa.x, x01.x, p8[0]); p8[1] = fma(-xa.y, x01.x, p8[1]);
    p8[2] = fma(-xa.x, x01.y, p8[2]); p8[3] = fma(-xa.y, x01.y, p8[3]);
    p8[4] = fma(-xa.x, x23.x, p8[4]); p8[5] = fma(-xa.y, x23.x, p8[5]);
    p8[6] = fma(-xa.x, x23.y, p8[6]); p8[7] = fma(-xa.y, x23.y, p8[7]);
  }
}
// element index of accumulator q of thread (rp, cp) in a column-major 32x32 tile
__device__ __forceinline__ int acc_elem(int rp, int cp, int q) { return 2 * rp + (q & 1) + 32 * (2 * cp + (q >> 1)); }

// X(a, c) = sum_{m <= c} P(a, m) W(c, m) for c = c0 + 8 jj: four accumulation chains per thread with warp-uniform bounds (W is lower
// triangular).  P in sP[a + 32 m], W in sW[c * kLP + m].
__device__ __forceinline__ void panel_times_winv_t(const double* sP, const double* sW, int a, int c0, double (&out)[4]) {
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) out[jj] = 0.0;
#pragma unroll
  for (int seg = 0; seg < 4; ++seg) {
#pragma unroll
    for (int mm = 0; mm < 8; ++mm) {
      const int m = seg * 8 + mm;
      const double xa = sP[a + 32 * m];
#pragma unroll
      for (int jj = seg + 1; jj < 4; ++jj) out[jj] = fma(xa, sW[(c0 + 8 * jj) * kLP + m], out[jj]);
      if (mm <= c0) out[seg] = fma(xa, sW[(c0 + 8 * seg) * kLP + m], out[seg]);
    }
  }
}

__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b) {   // C(8x8) += A(8x4, row) B(4x8, col)
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// dst(lower triangle) = src - X X^T (- Y Y^T) for 32x32 tiles, on the tensor pipe, by FOUR warps (wq = 0..3): the ten 8x8 blocks of the lower
// triangle as rows {3: 0 1 2}, {2: 0 1 2}, {(1,0) (1,1) (3,3)}, {(0,0)}.  src / dst have leading dimension 32, X and Y kXL.  `mirror` also
// writes the transposed off-diagonal blocks (a full symmetric tile).
__device__ __forceinline__ void lower_downdate_dmma(const double* src, double* dst, const double* X, const double* Y, int wq, int fg, int ft, bool mirror) {
  const int nt = wq == 3 ? 1 : 3;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    if (q >= nt) break;   // warp-uniform
    const int R = wq == 0 ? 3 : wq == 1 ? 2 : wq == 2 ? (q == 2 ? 3 : 1) : 0;
    const int C = wq <= 1 ? q : wq == 2 ? (q == 2 ? 3 : q) : 0;
    double d0 = src[(8 * R + fg) + 32 * (8 * C + 2 * ft)], d1 = src[(8 * R + fg) + 32 * (8 * C + 2 * ft + 1)];
    const double* xr = X + (8 * R + fg) + kXL * ft;
    const double* xc = X + (8 * C + fg) + kXL * ft;
    if (Y) {   // Y first: the same order of accumulation as when the helper warps took the Y term earlier (bitwise the same D either way)
      const double* yr = Y + (8 * R + fg) + kXL * ft;
      const double* yc = Y + (8 * C + fg) + kXL * ft;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) dmma_884(d0, d1, -yr[kXL * 4 * kk], yc[kXL * 4 * kk]);
    }
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) dmma_884(d0, d1, -xr[kXL * 4 * kk], xc[kXL * 4 * kk]);
    dst[(8 * R + fg) + 32 * (8 * C + 2 * ft)] = d0; dst[(8 * R + fg) + 32 * (8 * C + 2 * ft + 1)] = d1;
    if (mirror && R != C) { dst[(8 * C + 2 * ft) + 32 * (8 * R + fg)] = d0; dst[(8 * C + 2 * ft + 1) + 32 * (8 * R + fg)] = d1; }
  }
}

// ---- the pivot chain: one CTA per chain walks its block columns ---------------------------------------------------------------
//   D_j = Dpre_j - X_{j-1} X_{j-1}^T ; L_jj L_jj^T = D_j, W_j = L_jj^-1 ; X_j = L(j+1,j) = Ppre_j W_j^T
// Dpre_j and Ppre_j (the contributions of columns <= j-2 to the diagonal tile and to tile (j+1,j)) are pre-accumulated by worker CTAs
// and arrive as flagged copies; X_{j-1} and W_j never leave this CTA's shared memory on their way to the next step, so the serial chain
// itself has no global-memory hand-off.  While warps 0-2 factor and invert the diagonal block (tile_chol.cuh), warps 3, 5, 6, 7 fetch
// Ppre_j and the freshest tile L(j+1,j-1), apply the last update  Ppre_j -= L(j+1,j-1) X_{j-1}^T  and try to fetch Dpre_{j+1}; warp 4
// shares its scheduler with the head warp and stays idle.
__device__ void factor_chain(const BandSys& S, const int chain, FacShared& sh) {
  const int tid = threadIdx.x, a = tid & 31, c0 = tid >> 5;
  const int rp = tid & 15, cp = tid >> 4;   // 2x2 accumulator block: rows 2rp, 2rp+1, columns 2cp, 2cp+1
  const unsigned ep = S.epoch;
  const int c_start = chain == 0 ? 0 : S.NT0, c_end = chain == 0 ? S.NT0 : S.NT;
  const bool coupled = S.T >= 1;
  double* sA = sh.buf[0]; double* sB = sh.buf[1]; double* sD = sh.buf[2]; double* sM = sh.buf[3]; double* sX = sh.buf[4]; double* sT = sh.buf[5];
  double* sW = sh.sW;
  const int warp = tid >> 5;
  // leading dimensions of the operand tiles (sB = Ppre, sT = L(j+1,j-1), sX = X_j: element (r, c) at r + ld c; sW: W(r, c) at r ld + c)
  constexpr int XL = kChainDmma ? kXL : 32, WL = kChainDmma ? kXL : kLP;
  const int fg = a >> 2, ft = a & 3;   // this lane's row and k index inside an m8n8k4 operand fragment
  bool have_next = false;   // Dpre_{j+1} already in sD (fetched under column j's Cholesky)
  {  // the first column of a chain has nothing before it: its tile is already final
    const double* tile = S.tiles + static_cast<size_t>(c_start) * S.TPC * kTileElems;
    for (int e = tid; e < kTileElems; e += kFacThreads) sA[e] = tile[e];
    if (tid == 96) sh.progress = 0;
    __syncthreads();
  }
  for (int j = c_start; j < c_end; ++j) {
    const int tq = j * S.TPC;
    LVI_TRACE(2);
    const bool has_panel = coupled && j + 1 < c_end;
    const bool have_T = kChainDmma && has_panel && j != c_start && S.T >= 2;   // sT holds L(j+1,j-1) during this column
    int got_next = 1;
    if (warp == 0) {         // columns 0..15 and pivots 0..15
      const long long cyc0 = clock64();
      const bool ok = warp_potrf_head(sA, 32, sM, sh.sR, &sh.progress);
      if (!ok && tid == 0) *S.fail = 1;
      if (S.trace) {   // slot 7: wall time; slot 1: SM cycles the head took (their ratio is the SM clock under this load)
        if (a == 0) { S.trace[static_cast<size_t>(tq + S.T + 1) * 8 + 7] = gtime(); S.trace[static_cast<size_t>(tq + S.T + 1) * 8 + 1] = static_cast<unsigned long long>(clock64() - cyc0); }
        __syncwarp();
      }
    } else if (warp == 1) {  // columns 16..31: follows the head, then pivots 16..31
      const bool ok = warp_potrf_tail(sA, 32, sM, sh.sR, &sh.progress, S.trace ? S.trace + static_cast<size_t>(tq + S.T + 1) * 8 + 0 : nullptr);
      if (!ok && a == 0) *S.fail = 1;
      if (S.trace && a == 0) S.trace[static_cast<size_t>(tq + S.T + 1) * 8 + 5] = gtime();
    } else if (warp == 2) {  // W = L^-1, one chunk of pivots behind
      warp_inverse_cols<WL>(sM, sh.sR, &sh.progress, sW);
      if (S.trace && a == 0) S.trace[static_cast<size_t>(tq + S.T + 1) * 8 + 6] = gtime();
    } else if (warp == 4) {   // shares its scheduler with the head warp: stays idle
    } else if (has_panel) {   // warps 3, 5, 6, 7: Ppre_j, and its last update  Ppre_j -= L(j+1,j-1) X_{j-1}^T  (the freshest pair of tiles)
      const int h = (warp == 3 ? 0 : warp - 4) * 32 + a;   // 0..127
      if (j == c_start) {
        const double* tile = S.tiles + static_cast<size_t>(tq + 1) * kTileElems;
        for (int e = h; e < kTileElems; e += 128) sB[(e & 31) + XL * (e >> 5)] = tile[e];
      } else {
        const int hrow = tq + S.T + 1;   // diagnostics: the helpers stamp into the (otherwise unstamped) first border row of the column
        helper_fetch_tile<true, 128, XL>(ll_P(S, j), ep, sB, h, &sh.progress);
        if (S.T >= 2 && kChainDmma) {
          // Ppre_j -= L(j+1,j-1) X_{j-1}^T on the tensor pipe: helper warp hw owns the 8 rows 8 hw .. 8 hw + 7 of the tile (four 8x8
          // blocks); the negated row fragment of T is shared by the four.  32 DMMA + 40 conflict-free fragment loads per warp instead of
          // 256 DFMA + 96 16-byte loads per thread: the Cholesky warps next door keep the shared-memory pipe.
          helper_fetch_tile<true, 128, XL>(ll_tile(S, tq - S.TPC + 2), ep, sT, h, &sh.progress);
          if (S.trace && h == 0) S.trace[static_cast<size_t>(hrow) * 8 + 2] = gtime();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          const int hw = h >> 5;
          double acc[4][2];
#pragma unroll
          for (int C = 0; C < 4; ++C) { acc[C][0] = sB[(8 * hw + fg) + XL * (8 * C + 2 * ft)]; acc[C][1] = sB[(8 * hw + fg) + XL * (8 * C + 2 * ft + 1)]; }
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const double ta = -sT[(8 * hw + fg) + XL * (4 * kk + ft)];
#pragma unroll
            for (int C = 0; C < 4; ++C) dmma_884(acc[C][0], acc[C][1], ta, sX[(8 * C + fg) + XL * (4 * kk + ft)]);
          }
#pragma unroll
          for (int C = 0; C < 4; ++C) { sB[(8 * hw + fg) + XL * (8 * C + 2 * ft)] = acc[C][0]; sB[(8 * hw + fg) + XL * (8 * C + 2 * ft + 1)] = acc[C][1]; }
        } else if (S.T >= 2) {
          helper_fetch_tile<true, 128>(ll_tile(S, tq - S.TPC + 2), ep, sT, h, &sh.progress);
          if (S.trace && h == 0) S.trace[static_cast<size_t>(hrow) * 8 + 2] = gtime();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          const int rq = h & 15, cq = h >> 4;   // 2 x 4 block: rows 2rq.., columns 4cq..
          double p8[8];
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) { p8[2 * cc] = sB[2 * rq + 32 * (4 * cq + cc)]; p8[2 * cc + 1] = sB[2 * rq + 1 + 32 * (4 * cq + cc)]; }
#pragma unroll 8
          for (int m = 0; m < 32; ++m) {
            const double2 xa = *reinterpret_cast<const double2*>(sT + 2 * rq + 32 * m);
            const double2 x01 = *reinterpret_cast<const double2*>(sX + 4 * cq + 32 * m);
            const double2 x23 = *reinterpret_cast<const double2*>(sX + 4 * cq + 2 + 32 * m);
            p8[0] = fma(-xa.x, x01.x, p8[0]); p8[1] = fma(-xa.y, x01.x, p8[1]);
            p8[2] = fma(-xa.x, x01.y, p8[2]); p8[3] = fma(-xa.y, x01.y, p8[3]);
            p8[4] = fma(-xa.x, x23.x, p8[4]); p8[5] = fma(-xa.y, x23.x, p8[5]);
            p8[6] = fma(-xa.x, x23.y, p8[6]); p8[7] = fma(-xa.y, x23.y, p8[7]);
          }
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) { sB[2 * rq + 32 * (4 * cq + cc)] = p8[2 * cc]; sB[2 * rq + 1 + 32 * (4 * cq + cc)] = p8[2 * cc + 1]; }
        }
      }
      if (S.trace && h == 0) S.trace[static_cast<size_t>(tq + S.T + 1) * 8 + 3] = gtime();
      got_next = helper_fetch_tile<false, 128>(ll_tile(S, tq + S.TPC), ep, sD, h, &sh.progress) ? 1 : 0;
      if (kChainDmma) {
        // The worker task of the next diagonal tile stops TWO columns back (its inputs are then a whole column period old when the chain
        // gets here, so this fetch finds the tile waiting); the contribution of column j-1, L(j+1,j-1) L(j+1,j-1)^T, is taken here from
        // the tile the helpers already hold for the Ppre update -- still under the Cholesky's shadow.
        unsigned all_ok;
        asm volatile("{ .reg .pred p, q; setp.ne.u32 q, %1, 0; barrier.red.and.pred p, 1, 128, q; selp.u32 %0, 1, 0, p; }" : "=r"(all_ok) : "r"(static_cast<unsigned>(got_next)) : "memory");
        got_next = static_cast<int>(all_ok);
        if (all_ok && have_T) lower_downdate_dmma(sD, sD, sT, nullptr, h >> 5, fg, ft, false);
      }
      if (S.trace && h == 0) S.trace[static_cast<size_t>(tq + S.T + 1) * 8 + 4] = gtime();
    }
    have_next = __syncthreads_and(got_next) != 0 && has_panel;
    LVI_TRACE(3);
    {  // W_j goes out at once as a flagged copy: a whole block column of panel tasks is waiting for it
      unsigned long long* wll = ll_W(S, j);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int e = tid + kFacThreads * q4;
        ll_store(wll + 2 * e, sW[(e & 31) * WL + (e >> 5)], ep);
      }
    }
    LVI_TRACE(4);
    if (has_panel && kChainDmma) {
      // X_j = Ppre_j W_j^T on the tensor pipe.  W is lower triangular, so the 8 columns 8 C .. 8 C + 7 of X only need k < 8 (C + 1): warp w
      // takes row block w / 2 and the column blocks {0, 3} or {1, 2} (10 k-steps of 4 either way), sharing the P fragment between its two.
      const int R = warp >> 1, Ca = (warp & 1) ? 1 : 0, Cb = (warp & 1) ? 2 : 3;
      double xa0 = 0.0, xa1 = 0.0, xb0 = 0.0, xb1 = 0.0;
      const double* pP = sB + (8 * R + fg) + XL * ft;
      const double* wA = sW + (8 * Ca + fg) * WL + ft;
      const double* wB = sW + (8 * Cb + fg) * WL + ft;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        if (kk < 2 * (Cb + 1)) {   // warp-uniform
          const double pa = pP[XL * 4 * kk];
          if (kk < 2 * (Ca + 1)) dmma_884(xa0, xa1, pa, wA[4 * kk]);
          dmma_884(xb0, xb1, pa, wB[4 * kk]);
        }
      }
      sX[(8 * R + fg) + XL * (8 * Ca + 2 * ft)] = xa0; sX[(8 * R + fg) + XL * (8 * Ca + 2 * ft + 1)] = xa1;   // stays here for the next diagonal tile
      sX[(8 * R + fg) + XL * (8 * Cb + 2 * ft)] = xb0; sX[(8 * R + fg) + XL * (8 * Cb + 2 * ft + 1)] = xb1;   // and the next Ppre update
      {  // the flagged copy of X_j = L(j+1,j) leaves straight from the accumulator fragments: a block column of worker tasks is waiting for it
        unsigned long long* sll = ll_tile(S, tq + 1);
        const int ea = (8 * R + fg) + 32 * (8 * Ca + 2 * ft), eb = (8 * R + fg) + 32 * (8 * Cb + 2 * ft);
        ll_store(sll + 2 * ea, xa0, ep); ll_store(sll + 2 * (ea + 32), xa1, ep);
        ll_store(sll + 2 * eb, xb0, ep); ll_store(sll + 2 * (eb + 32), xb1, ep);
      }
    } else if (has_panel) {
      double out[4];
      panel_times_winv_t(sB, sW, a, c0, out);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) sX[a + 32 * (c0 + 8 * jj)] = out[jj];   // stays here for the next diagonal tile and the next Ppre update
    }
    __syncthreads();  // X_j is in sX
    LVI_TRACE(5);
    // The CTA splits: warps 4-7 write X_j (flagged + plain) and the plain copy of W_j to global memory (48 KB of stores: pure issue time),
    // warps 0-3 build the next diagonal tile  D_{j+1} = Dpre_{j+1} - X_j X_j^T  with 2 x 4 register blocks (four warps with 8 outputs per
    // thread finish this shared-memory-bound update sooner than eight warps with 4).
    if (warp >= 4) {
      const int h = tid - 128;
      if (has_panel) {
        unsigned long long* sll = ll_tile(S, tq + 1);
        double* tile = S.tiles + static_cast<size_t>(tq + 1) * kTileElems;
#pragma unroll
        for (int q8 = 0; q8 < 8; ++q8) {
          const int e = h + 128 * q8;
          const double x = sX[(e & 31) + XL * (e >> 5)];
          if (!kChainDmma) ll_store(sll + 2 * e, x, ep);
          if (!worker_writes_X(S, j, c_end)) tile[e] = x;
        }
      }
      if (!worker_writes_W(S, j, c_end)) {
        double* Wg = S.Linv + static_cast<size_t>(j) * kTileElems;
#pragma unroll
        for (int q8 = 0; q8 < 8; ++q8) {
          const int e = h + 128 * q8;
          Wg[e] = sW[(e & 31) * WL + (e >> 5)];
        }
      }
    } else if (j + 1 < c_end) {
      const int h = tid;                    // 0..127
      const int rq = h & 15, cq = h >> 4;   // 2 x 4 block: rows 2rq.., columns 4cq..
      double p8[8];
      if (!coupled) {
        const double* tile = S.tiles + static_cast<size_t>(tq + S.TPC) * kTileElems;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) { p8[2 * cc] = tile[2 * rq + 32 * (4 * cq + cc)]; p8[2 * cc + 1] = tile[2 * rq + 1 + 32 * (4 * cq + cc)]; }
      } else {
        if (!have_next) {
          helper_fetch_tile<true, 128>(ll_tile(S, tq + S.TPC), ep, sD, h, &sh.progress);
          asm volatile("bar.sync 2, 128;" ::: "memory");
        }
        LVI_TRACE(6);
        if (kChainDmma) {
          // D_{j+1} = Dpre_{j+1} - X_j X_j^T on the tensor pipe (lower triangle, mirrored on the way out: the head / tail Cholesky read whole
          // columns).  When the helpers could not fetch Dpre_{j+1} under the Cholesky, the column j-1 term L(j+1,j-1) L(j+1,j-1)^T is taken here.
          lower_downdate_dmma(sD, sA, sX, (!have_next && have_T) ? sT : nullptr, warp, fg, ft, true);
          goto d_done;
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) { p8[2 * cc] = sD[2 * rq + 32 * (4 * cq + cc)]; p8[2 * cc + 1] = sD[2 * rq + 1 + 32 * (4 * cq + cc)]; }
#pragma unroll 8
        for (int m = 0; m < 32; ++m) {
          const double2 xa = *reinterpret_cast<const double2*>(sX + 2 * rq + 32 * m);
          const double2 x01 = *reinterpret_cast<const double2*>(sX + 4 * cq + 32 * m);
          const double2 x23 = *reinterpret_cast<const double2*>(sX + 4 * cq + 2 + 32 * m);
          p8[0] = fma(-xa.x, x01.x, p8[0]); p8[1] = fma(-xa.y, x01.x, p8[1]);
          p8[2] = fma(-xa.x, x01.y, p8[2]); p8[3] = fma(-xa.y, x01.y, p8[3]);
          p8[4] = fma(-xa.x, x23.x, p8[4]); p8[5] = fma(-xa.y, x23.x, p8[5]);
          p8[6] = fma(-xa.x, x23.y, p8[6]); p8[7] = fma(-xa.y, x23.y, p8[7]);
        }
      }
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) { sA[2 * rq + 32 * (4 * cq + cc)] = p8[2 * cc]; sA[2 * rq + 1 + 32 * (4 * cq + cc)] = p8[2 * cc + 1]; }
    d_done:;
    }
    if (tid == 96) sh.progress = 0;
    __syncthreads();  // D_{j+1} is in sA; the stores of column j are issued
    LVI_TRACE(7);
  }
  // No ready flags for the chain's tiles: inside this kernel W_j and X_j are only ever consumed as flagged copies (they are the
  // freshest inputs of their consumers by construction), and the plain copies are for the back-substitution kernel.  A fence + release
  // here is not just unnecessary: the membar stalled the whole CTA's shared-memory traffic for ~1.7 us per column (measured).
}

// ---- flag-driven left-looking band Cholesky -----------------------------------------------------------------------------------
// Every 32x32 tile of the factor off the pivot chain is ONE task: a worker CTA fetches tasks in column-major order from a global
// counter, keeps the tile's accumulator in registers and subtracts L(i,k) L(j,k)^T for k ascending AS SOON AS the two source tiles are
// published, then multiplies by W_j^T and publishes the tile.  Each tile is written once by one CTA (no read-modify-write on HBM, no
// grid-wide barrier).  Older source tiles are read plainly behind their ready flag (ld.acquire / st.release); the LAST update of a task
// -- from the column that has only just been finished -- and W_j come through the flagged copies, because every tile follows the
// recurrence  W_j -> L(i,j) -> update of L(i,j+1) -> W_{j+1} -> ...  and that loop has to close within one column period of the chain.
// The tasks of the diagonal tile and of the first sub-diagonal tile only PRE-ACCUMULATE (they hand Dpre / Ppre to the chain CTA, see
// factor_chain).  Tasks are fetched in dependency order and the chain CTAs are resident from the start, so a fetched task only ever
// waits on work that is already running: no deadlock for any grid size.
__global__ void __launch_bounds__(kFacThreads, 2) band_factor_ll_kernel(BandSys S) {
  extern __shared__ __align__(128) unsigned char fac_smem[];
  FacShared& sh = *reinterpret_cast<FacShared*>(fac_smem);
  int* flags = S.work_i;
  int* counter = S.work_i + static_cast<size_t>(S.NT) * S.TPC + S.NT;
  int* chain_sm = counter + 2;   // [2] SM id + 1 of the chain CTAs
  const unsigned ep = S.epoch;
  const int n_chains = (S.NT0 > 0 && S.NT0 < S.NT) ? 2 : 1;
  const int ntask = (S.NT + S.pre_shift) * S.TPC;
  const int tid = threadIdx.x;
  const int a = tid & 31, c0 = tid >> 5;
  const int rp = tid & 15, cp = tid >> 4;   // 2x2 accumulator block: rows 2rp, 2rp+1, columns 2cp, 2cp+1
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  // chain CTAs: blocks 0 and 2 (consecutive blocks land on the two SMs of one TPC; the chains get a TPC each)
  const int chain_id = blockIdx.x == 0 ? 0 : (blockIdx.x == 2 && n_chains == 2 && gridDim.x > 3) ? 1 : (blockIdx.x == 1 && n_chains == 2 && gridDim.x <= 3) ? 1 : -1;
  if (chain_id >= 0) {
    if (tid == 0) st_release(chain_sm + chain_id, static_cast<int>(smid) + 1);
    factor_chain(S, chain_id, sh);
    return;
  }
  // a worker that shares its SM with a chain CTA steps aside: the chain is the critical path and runs ~1.5x faster alone on the SM
  if (tid == 0) {
    int leave = 0;
    for (int c = 0; c < n_chains; ++c) {
      int v;
      while ((v = ld_acquire(chain_sm + c)) == 0) {}
      leave |= (v == static_cast<int>(smid) + 1);
    }
    sh.leave = leave && gridDim.x > static_cast<unsigned>(n_chains + 2);
  }
  __syncthreads();
  if (sh.leave) return;
  double* sA = sh.buf[0]; double* sB = sh.buf[1]; double* sW = sh.sW;
  if (tid == 0) {
    for (int st = 0; st < kStages; ++st) mbar_init(&sh.full[st], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned g = 0;   // pipelined update steps this CTA has run so far: step n uses stage n % kStages, mbarrier phase (n / kStages) & 1
  while (true) {
    if (tid == 0) sh.q = atomicAdd(counter, 1);
    __syncthreads();
    const int q = sh.q;
    __syncthreads();
    if (q >= ntask) break;
    // The two pre-accumulation tasks of a column are what the chain waits for, and like every task they run ~T sequential updates: they
    // are queued pre_shift column slots EARLIER than the other tiles of their column, so that only their last update is left when the
    // chain arrives.  (Their inputs from the last two columns are then produced by tasks LATER in the queue; the host picks pre_shift so
    // that the resident workers always cover that span, which keeps the no-deadlock argument intact.)
    const int slot = q / S.TPC, s = q - slot * S.TPC;
    // The tiles right below them (s = 2 .. 4) run the longest serial update sequences of a column (T - s rank-32 updates, ~30 us) and feed
    // the chain's helper warps directly: queued with everything else they finished their old updates only 2.6 us (median; +3.6 us at the
    // 95th percentile) around the moment W_j appeared (tools/diag/trace_s2.py), and every late one stalls the chain.  They are queued one
    // (s = 3, 4) or two (s = 2) slots earlier; their producers then sit at most two slots later in the queue, like those of the s <= 1 tasks.
    const int early = s < 64 ? static_cast<int>(S.lead[s]) : 0;
    const int jo = (s <= S.T) ? slot - (S.pre_shift - early) : slot - S.pre_shift;
    if (jo < 0 || jo >= S.NT) continue;
    const int j = ordered_column(jo, S.NT0, S.NT);
    const int c_start = j < S.NT0 ? 0 : S.NT0, c_end = j < S.NT0 ? S.NT0 : S.NT;   // the chain this column belongs to
    const bool band = s <= S.T;
    const int i = j + s;
    if (band && i >= c_end) continue;  // tile below the end of the chain: never referenced
    if (band && s <= 1 && (j == c_start || S.T == 0)) continue;   // first column of a chain: the chain CTA reads those tiles directly
    // separator rows of the two-sided ordering couple only to the last <= bw positions of a chain: their border tiles are
    // structurally zero (and stay zero in the factor) before block column sep_first, so those tasks and products are skipped
    const int RBsep = S.n_mid >> kTileLog;                       // border tile rows made of separator dims only
    const int sep_first = max(c_start, c_end - S.T - 1);
    const bool sep_row = !band && (s - S.T - 1) < RBsep;
    if (sep_row && j < sep_first) continue;
    const int tq = j * S.TPC + s;      // storage / flag index of this tile
    const bool eager = band && s <= 3; // next to the pivot chain
    const bool stamp = band;                                 // diagnostics (LVI_TRACE_FACTOR)
    const int trow = j * S.TPC + (s == 0 ? S.TPC - 1 : s);   // the diagonal task stamps into the (unused) last row of its column
    double* tile = S.tiles + static_cast<size_t>(tq) * kTileElems;
    if (stamp) LVI_TRACE_AT(trow, 0);
    double acc[4];
    const int kmin = max(sep_row ? sep_first : c_start, band ? i - S.T : j - S.T);
    // the two pre-accumulation tasks leave the update from column j-1 to the chain CTA; the diagonal one also leaves column j-2 (the chain's
    // helper warps hold L(j,j-2) anyway, factor_chain), so its last input is a whole column period old when the chain asks for the tile
    const int kend = (band && s == 0 && kChainDmma && S.T >= 2) ? j - 2 : (band && s <= 1) ? j - 1 : j;
    // Updates from the older columns: their source tiles are staged into shared memory by bulk async copies (TMA) running two steps
    // ahead of the arithmetic, one elected thread issuing them as soon as the ready flags allow (one warp looks at the flags of up to 32
    // steps at once: one L2 round trip per batch, not per step).  The LAST update takes the flagged copies (below).
    auto tile_of = [&](int k, int& fi, int& fj) { fi = k * S.TPC + (band ? i - k : s); fj = k * S.TPC + (j - k); };
    const int n_old = max(kend - 1 - kmin, 0);
    const bool any_update = kend - 1 >= kmin;
    const int half = tid >> 7, rq = tid & 15, cq = (tid >> 4) & 7;   // partial accumulators: half 0 starts from the tile, half 1 from zero
    double p8[8];
    if (any_update) {
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        p8[2 * cc] = half == 0 ? tile[2 * rq + 32 * (4 * cq + cc)] : 0.0;
        p8[2 * cc + 1] = half == 0 ? tile[2 * rq + 1 + 32 * (4 * cq + cc)] : 0.0;
      }
    } else {
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) acc[q4] = tile[acc_elem(rp, cp, q4)];
    }
    int issued = 0, ready_n = 0;   // warp 0: steps issued / steps known to be published
    auto try_issue = [&](int upto) {   // warp 0 only
      upto = min(upto, n_old);
      while (issued < upto) {
        if (issued >= ready_n) {
          const int kk = kmin + issued + a;
          bool ok = false;
          if (kk < kend - 1) { int fi, fj; tile_of(kk, fi, fj); ok = ld_acquire(flags + fi) != 0 && ld_acquire(flags + fj) != 0; }
          const unsigned mask = __ballot_sync(FULL, ok);
          const int n = mask == FULL ? 32 : __ffs(~mask) - 1;
          if (n == 0) return;
          ready_n = issued + n;
        }
        if (a == 0) {
          int fi, fj;
          tile_of(kmin + issued, fi, fj);
          const unsigned st = (g + issued) % kStages;
          asm volatile("fence.proxy.async;" ::: "memory");   // the acquire above orders the generic proxy; the copy reads through the async proxy
          mbar_expect_tx(&sh.full[st], fi == fj ? 8192u : 16384u);
          tma_load_1d(sh.buf[2 * st], S.tiles + static_cast<size_t>(fi) * kTileElems, 8192u, &sh.full[st]);
          if (fi != fj) tma_load_1d(sh.buf[2 * st + 1], S.tiles + static_cast<size_t>(fj) * kTileElems, 8192u, &sh.full[st]);
        }
        ++issued;
      }
    };
    if (tid < 32) try_issue(kStages - 1);
    // kPair updates per CTA barrier: the barrier releases the stages of the pair, so steps t .. t + kStages - 1 may be in flight at its start
    for (int t = 0; t < n_old; t += kPair) {
      if (tid < 32) {
        const int need = min(t + kPair, n_old);   // the steps of this pair must be on their way before anybody waits for them
        try_issue(t + kStages);
        while (issued < need) { __nanosleep(200); try_issue(t + kStages); }
      }
#pragma unroll
      for (int u = 0; u < kPair; ++u) {
        if (t + u >= n_old) break;
        const unsigned st = (g + t + u) % kStages;
        mbar_wait(&sh.full[st], ((g + t + u) / kStages) & 1u);
        if (stamp && t + u == n_old - 1) LVI_TRACE_AT(trow, 6);   // inputs of the second-to-last update in shared memory
        if ((u & 1) == half) rank32_update_2x4(sh.buf[2 * st], (band && s == 0) ? sh.buf[2 * st] : sh.buf[2 * st + 1], rq, cq, p8, 0, 32);
      }
      __syncthreads();   // everybody is done with these stages (per-stage `empty` mbarriers instead of this barrier let the warps drift apart
                         // and measured slower: 2.08 against 1.92 ms)
    }
    g += n_old;
    if (kend - 1 >= kmin) {  // the freshest inputs -- the column that has only just been finished -- come as flagged copies
      int fi, fj;
      tile_of(kend - 1, fi, fj);
      __syncthreads();    // the previous k-step's reads of sA/sB are complete
      if (stamp) LVI_TRACE_AT(trow, 2);
      fetch_ll_tile(ll_tile(S, fi), ep, eager, sA, fj == fi ? sB : nullptr);
      // fj is tile (j, kend-1); for the s = 2 task that is X_{j-1} = L(j, j-1), whose plain copy this task writes (worker_writes_X)
      const bool plain_x = band && s == 2 && kend == j && fj != fi && worker_writes_X(S, j - 1, c_end);
      if (fj != fi) fetch_ll_tile(ll_tile(S, fj), ep, eager, sB, nullptr, plain_x ? S.tiles + static_cast<size_t>(fj) * kTileElems : nullptr);
      __syncthreads();
      if (stamp) LVI_TRACE_AT(trow, 3);
      rank32_update_2x4(sA, sB, rq, cq, p8, 16 * half, 16 * half + 16);   // the freshest update: its two k-halves on the two halves of the CTA
    }
    if (any_update) {   // add the two partial accumulators (through shared memory) and go back to the 2x2 layout of the rest of the task
      __syncthreads();
      double* sP = half == 0 ? sA : sB;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) *reinterpret_cast<double2*>(sP + 2 * rq + 32 * (4 * cq + cc)) = make_double2(p8[2 * cc], p8[2 * cc + 1]);
      __syncthreads();
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) acc[q4] = sA[acc_elem(rp, cp, q4)] + sB[acc_elem(rp, cp, q4)];
    }
    if (stamp) LVI_TRACE_AT(trow, 4);
    if (band && s <= 1) {  // pre-accumulated diagonal / first sub-diagonal tile: hand it to the chain CTA
      unsigned long long* dst = s == 0 ? ll_tile(S, tq) : ll_P(S, j);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) ll_store(dst + 2 * acc_elem(rp, cp, q4), acc[q4], ep);
      if (stamp) LVI_TRACE_AT(trow, 7);
      continue;
    }
    __syncthreads();
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) sA[acc_elem(rp, cp, q4)] = acc[q4];
    {  // W_j, transposed into sW[c * kLP + m]
      const unsigned long long* wll = ll_W(S, j);
      if (!eager) {
        if (tid == 0) {
          unsigned long long w0, w1;
          while (true) {
            asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(wll) : "memory");
            if (ll_valid(w0, w1, ep)) break;
            __nanosleep(200);
          }
        }
        __syncthreads();
      }
      double v[4];
      ll_load_n<4>(wll + 2 * tid, 2 * kFacThreads, v, ep);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) sW[a * kLP + c0 + 8 * q4] = v[q4];   // element e = tid + 256 q4: row e & 31 = a, column e >> 5
      if (band && s == 2 && worker_writes_W(S, j, c_end)) {   // plain copy of W_j (see worker_writes_W)
        double* Wg = S.Linv + static_cast<size_t>(j) * kTileElems;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) Wg[tid + kFacThreads * q4] = v[q4];
      }
    }
    if (stamp) LVI_TRACE_AT(trow, 5);
    __syncthreads();
    double out[4];
    panel_times_winv_t(sA, sW, a, c0, out);
    {
      unsigned long long* sll = ll_tile(S, tq);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int e = a + 32 * (c0 + 8 * jj);
        ll_store(sll + 2 * e, out[jj], ep);
        tile[e] = out[jj];
      }
    }
    if (stamp) LVI_TRACE_AT(trow, 7);
    __syncthreads();
    if (tid == 0) { __threadfence(); st_release(flags + tq, 1); }
    if (stamp) LVI_TRACE_AT(trow, 1);   // ready flag released
    if (!band) {  // border tile rb of column j: its share of the corner's Schur complement, C(rb, bj) -= Lb(rb,j) Lb(bj,j)^T for bj <= rb
      const int rb = s - S.T - 1;
      __syncthreads();
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) sA[a + 32 * (c0 + 8 * jj)] = out[jj];   // own finished tile stays in shared memory
      for (int bj = 0; bj <= rb; ++bj) {
        if (j < sep_first && bj < RBsep) continue;   // zero separator tile: the product vanishes
        const double* Xb_s = sA;
        if (bj < rb) {
          const int fb = j * S.TPC + S.T + 1 + bj;
          if (tid == 0) spin_until_set(flags + fb, false);
          __syncthreads();
          const double2* Xb = reinterpret_cast<const double2*>(S.tiles + static_cast<size_t>(fb) * kTileElems);
          for (int e = tid; e < kTileElems / 2; e += kFacThreads) reinterpret_cast<double2*>(sB)[e] = __ldcg(Xb + e);
          Xb_s = sB;
        }
        __syncthreads();
        double pr[4] = {0.0, 0.0, 0.0, 0.0};
        for (int m = 0; m < 32; ++m) {
          const double xa = sA[a + 32 * m];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) pr[jj] = fma(xa, Xb_s[c0 + 8 * jj + 32 * m], pr[jj]);
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          if (pr[jj] != 0.0) atomicAdd(S.C + 32 * rb + a + static_cast<size_t>(S.ldc) * (32 * bj + c0 + 8 * jj), -pr[jj]);
        __syncthreads();
      }
    }
  }
}

// ---- backward substitution: x_m = W_m^T (z_m - Lb(:,m)^T x2 - sum_{d=1..T} L(m+d,m)^T x_{m+d}),  m descending along each chain ------
// The recurrence is serial in m, so its cost is (number of block columns) x (latency of one step).  One CTA per chain walks its columns
// and does only what depends on the LATEST solutions: the products with the kBsLocal nearest sub-diagonal tiles and with W_m, on tiles
// that bulk async copies (TMA) stage into shared memory two columns ahead -- x_{m+1} never leaves the CTA on its way to x_m.  Everything
// else of a column (the border part, the rhs, and the products with the tiles further down, whose x are at least kBsLocal + 1 steps
// old) is one task of a worker CTA, which hands u_m = z_m - Lb^T x2 - sum_{d > kBsLocal} ... to the chain as a flagged vector; workers
// pick up each x_m as a flagged vector as well.  (The previous version passed x_m from CTA to CTA through a flagged mailbox: ~1 us of
// hand-off plus a one-warp product per column, 4.4 us per column; this one is bound by the chain SM's copy bandwidth, ~0.5 us per column.)
#ifndef LVI_BS_LOCAL
#define LVI_BS_LOCAL 4
#endif
#ifndef LVI_BS_STAGES
#define LVI_BS_STAGES 3
#endif
constexpr int kBsLocal = LVI_BS_LOCAL;
constexpr int kBsStages = LVI_BS_STAGES;
struct BsShared {
  double tiles[kBsStages][kBsLocal + 1][kTileElems];   // [stage][0] = W_m, [stage][d] = L(m+d, m)
  double x2s[1024];                                      // border solution (workers)
  double xring[kBsLocal + 1][kTile];                     // chain: the last kBsLocal + 1 solutions, slot = column % (kBsLocal + 1)
  double su[kTile], sl[kTile];
  double sred[8][kTile], xv[8][kTile];                   // workers: per-warp partial sums / per-warp copy of one x_{m+d}
  unsigned long long full[kBsStages];
  int q;
};
__device__ __forceinline__ unsigned long long* bs_xll(const BandSys& S) { return S.ll + (static_cast<size_t>(S.NT) * S.TPC + 2 * static_cast<size_t>(S.NT)) * 2 * kTileElems; }
__device__ __forceinline__ unsigned long long* bs_ull(const BandSys& S) { return bs_xll(S) + static_cast<size_t>(S.NT) * 2 * kTile; }

__device__ void backsolve_chain(const BandSys& S, const int chain, BsShared& sh) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned ep = S.epoch;
  const int c_start = chain == 0 ? 0 : S.NT0, c_end = chain == 0 ? S.NT0 : S.NT;
  const int n_cols = c_end - c_start;
  unsigned long long* xll = bs_xll(S);
  const unsigned long long* ull = bs_ull(S);
  auto issue = [&](int t) {   // one thread: stage W_m and the nearest sub-diagonal tiles of step t
    const int m = c_end - 1 - t;
    const int Dm = min(kBsLocal, min(S.T, t));
    const int st = t % kBsStages;
    mbar_expect_tx(&sh.full[st], static_cast<unsigned>(1 + Dm) * 8192u);
    tma_load_1d(sh.tiles[st][0], S.Linv + static_cast<size_t>(m) * kTileElems, 8192u, &sh.full[st]);
    if (Dm > 0) tma_load_1d(sh.tiles[st][1], S.tiles + (static_cast<size_t>(m) * S.TPC + 1) * kTileElems, static_cast<unsigned>(Dm) * 8192u, &sh.full[st]);
  };
  if (tid == 0) {
    for (int st = 0; st < kBsStages; ++st) mbar_init(&sh.full[st], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int t = 0; t < min(kBsStages - 1, n_cols); ++t) issue(t);
  }
  // u_m arrives as a flagged vector: warp 0 keeps the loads of the next two steps in flight
  unsigned long long a0 = 0, a1 = 0, b0 = 0, b1 = 0;
  auto u_issue = [&](int t, unsigned long long& w0, unsigned long long& w1) {
    w0 = w1 = 0;
    if (t < n_cols) {
      const unsigned long long* src = ull + (static_cast<size_t>(c_end - 1 - t) * kTile + lane) * 2;
      asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src) : "memory");
    }
  };
  if (warp == 0) { u_issue(0, a0, a1); u_issue(1, b0, b1); }
  __syncthreads();
  for (int t = 0; t < n_cols; ++t) {
    const int m = c_end - 1 - t;
    const int Dm = min(kBsLocal, min(S.T, t));
    const int st = t % kBsStages;
    if (tid == 0 && t + kBsStages - 1 < n_cols) issue(t + kBsStages - 1);   // its stage was released by the barrier that ended step t - 1
    if (warp == 0) {
      unsigned long long w0 = a0, w1 = a1;
      a0 = b0; a1 = b1;
      u_issue(t + 2, b0, b1);
      const unsigned long long* src = ull + (static_cast<size_t>(m) * kTile + lane) * 2;
      while (!ll_valid(w0, w1, ep)) asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src) : "memory");
      sh.su[lane] = ll_value(w0, w1);
    }
    mbar_wait(&sh.full[st], (t / kBsStages) & 1u);
    double part[4] = {0.0, 0.0, 0.0, 0.0};   // lane = row r of the tiles, this warp's 4 columns; summed over the lanes afterwards
    for (int d = 1; d <= Dm; ++d) {
      const double* tl = sh.tiles[st][d];
      const double xr = sh.xring[(m + d) % (kBsLocal + 1)][lane];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) part[cc] = fma(tl[lane + 32 * (4 * warp + cc)], xr, part[cc]);
    }
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part[cc] += __shfl_xor_sync(FULL, part[cc], o);
    }
    if (lane < 4) sh.sl[4 * warp + lane] = lane == 0 ? part[0] : lane == 1 ? part[1] : lane == 2 ? part[2] : part[3];
    __syncthreads();
    const double vr = sh.su[lane] - sh.sl[lane];
    const double* W = sh.tiles[st][0];
    double xs[4];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      xs[cc] = W[lane + 32 * (4 * warp + cc)] * vr;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) xs[cc] += __shfl_xor_sync(FULL, xs[cc], o);
    }
    if (lane < 4) {
      const int c = 4 * warp + lane;
      const double xc = lane == 0 ? xs[0] : lane == 1 ? xs[1] : lane == 2 ? xs[2] : xs[3];
      sh.xring[m % (kBsLocal + 1)][c] = xc;
      ll_store(xll + (static_cast<size_t>(m) * kTile + c) * 2, xc, ep);
      S.x[static_cast<size_t>(m) * kTile + c] = xc;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) band_backsolve_ll_kernel(BandSys S) {
  extern __shared__ __align__(128) unsigned char bs_smem[];
  BsShared& sh = *reinterpret_cast<BsShared*>(bs_smem);
  const int n_chains = (S.NT0 > 0 && S.NT0 < S.NT) ? 2 : 1;
  if (static_cast<int>(blockIdx.x) < n_chains) { backsolve_chain(S, blockIdx.x, sh); return; }
  int* counter = S.work_i + static_cast<size_t>(S.NT) * S.TPC + S.NT + 1;
  const unsigned ep = S.epoch;
  const unsigned long long* xll = bs_xll(S);
  unsigned long long* ull = bs_ull(S);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < S.ldc; i += 256) sh.x2s[i] = S.x[static_cast<size_t>(S.NT) * kTile + i];
  const int zrb = S.nbo >> kTileLog, zrow = S.nbo & (kTile - 1);
  const int RBsep = S.n_mid >> kTileLog;
  while (true) {
    __syncthreads();
    if (tid == 0) sh.q = atomicAdd(counter, 1);
    __syncthreads();
    const int q = sh.q;
    if (q >= S.NT) break;
    const int m = ordered_column_desc(q, S.NT0, S.NT);
    const int c_start = m < S.NT0 ? 0 : S.NT0, c_end = m < S.NT0 ? S.NT0 : S.NT;
    const int Tm = min(S.T, c_end - 1 - m);
    const double* col = S.tiles + static_cast<size_t>(m) * S.TPC * kTileElems;
    const int rb0 = (m < max(c_start, c_end - S.T - 1)) ? RBsep : 0;   // zero separator tiles are skipped
    const int n_border = S.RB - rb0, n_far = max(Tm - kBsLocal, 0);
    double acc = 0.0;   // lane c: column c of this warp's share of the products
    for (int it = warp; it < n_border + n_far; it += 8) {
      const bool border = it < n_border;
      const int d = border ? 0 : Tm - (it - n_border);   // far tiles from the oldest x to the newest
      const double* tc = col + static_cast<size_t>(border ? S.T + 1 + rb0 + it : d) * kTileElems + kTile * lane;
      double tv[kTile];
#pragma unroll
      for (int r = 0; r < kTile; ++r) tv[r] = tc[r];   // all loads in flight at once: one memory round trip per tile
      const double* vec;
      if (border) {
        vec = sh.x2s + kTile * (rb0 + it);
      } else {
        sh.xv[warp][lane] = ll_load(xll + (static_cast<size_t>(m + d) * kTile + lane) * 2, ep);
        __syncwarp();
        vec = sh.xv[warp];
      }
      double u0 = 0.0, u1 = 0.0;
#pragma unroll
      for (int r = 0; r < kTile; r += 2) { u0 = fma(tv[r], vec[r], u0); u1 = fma(tv[r + 1], vec[r + 1], u1); }
      acc += u0 + u1;
      __syncwarp();
    }
    sh.sred[warp][lane] = acc;
    __syncthreads();
    if (warp == 0) {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) tot += sh.sred[w][lane];
      const double z = col[static_cast<size_t>(S.T + 1 + zrb) * kTileElems + zrow + kTile * lane];
      ll_store(ull + (static_cast<size_t>(m) * kTile + lane) * 2, z - tot, ep);
    }
  }
}

// dense Cholesky of the border Schur complement S (nbo x nbo, lower, column-major ld = ldc) with the rhs carried as row nbo,
// then x2 = L^-T z2.  One CTA.
__global__ void __launch_bounds__(256) corner_solve_kernel(BandSys S, int in_smem) {
  extern __shared__ double corner_smem[];
  const int n = S.nbo, tid = threadIdx.x;
  double* x2 = S.x + static_cast<size_t>(S.NT) * kTile;
  __shared__ double xs[1024];
  __shared__ int bad;
  if (tid == 0) bad = 0;
  // the corner is small (45 rows at C2) and every pivot is three dependent passes over it: from shared memory a pass costs a barrier,
  // from global memory an L2 round trip (64 us -> 10 us at C2); large corners (no shared-memory budget) stay in global memory
  double* C = S.C;
  int ld = S.ldc;
  if (in_smem) {
    const int lds = n + 2;
    for (int e = tid; e < (n + 1) * (n + 1); e += 256) {
      const int i = e % (n + 1), c = e / (n + 1);
      corner_smem[i + lds * c] = S.C[i + static_cast<size_t>(S.ldc) * c];
    }
    C = corner_smem; ld = lds;
  }
  __syncthreads();
  for (int j = 0; j < n; ++j) {
    double d = C[j + static_cast<size_t>(ld) * j];
    if (!(d > 0.0) || !isfinite(d)) { if (tid == 0) bad = 1; d = 1.0; }
    const double s = sqrt(d);
    __syncthreads();
    if (tid == 0) C[j + static_cast<size_t>(ld) * j] = s;
    for (int i = j + 1 + tid; i <= n; i += 256) C[i + static_cast<size_t>(ld) * j] /= s;
    __syncthreads();
    const int m = n - j;  // rows j+1..n (n = rhs row), columns j+1..n-1
    for (int e = tid; e < m * m; e += 256) {
      const int i = j + 1 + e % m, c = j + 1 + e / m;
      if (c <= i && c < n) C[i + static_cast<size_t>(ld) * c] -= C[i + static_cast<size_t>(ld) * j] * C[c + static_cast<size_t>(ld) * j];
    }
    __syncthreads();
  }
  for (int i = tid; i < S.ldc; i += 256) xs[i] = 0.0;
  __syncthreads();
  if (tid < 32) {
    for (int c = n - 1; c >= 0; --c) {
      double part = 0.0;
      for (int i = c + 1 + tid; i < n; i += 32) part += C[i + static_cast<size_t>(ld) * c] * xs[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
      if (tid == 0) xs[c] = (C[n + static_cast<size_t>(ld) * c] - part) / C[c + static_cast<size_t>(ld) * c];
      __syncwarp();
    }
  }
  __syncthreads();
  for (int i = tid; i < S.ldc; i += 256) x2[i] = xs[i];
  if (tid == 0 && bad) *S.fail = 1;
}

// ---- second-level system of the two-sided ordering ------------------------------------------------------------------------------
// After both chains are factored the corner C = [separator | map-time knots + sensors | rhs]^2 holds the Schur complement of the whole
// border.  The separator part (n_mid ~ the half bandwidth) is too large for the single-CTA dense corner solve, so it is re-packed as a
// (dense) band system B with the remaining border as ITS border and goes through the same tile factorisation.
__global__ void __launch_bounds__(256) corner_to_second_level_kernel(BandSys A, BandSys B) {
  const int ntile = B.NT * B.TPC;
  const int nm = A.n_mid;
  if (static_cast<int>(blockIdx.x) < ntile) {
    const int J = blockIdx.x / B.TPC, q = blockIdx.x % B.TPC;
    if (q <= B.T && J + q >= B.NT) return;
    double* dst = B.tiles + static_cast<size_t>(blockIdx.x) * kTileElems;
    for (int e = threadIdx.x; e < kTileElems; e += blockDim.x) {
      const int a = e & (kTile - 1), b = e >> kTileLog;
      const int j = J * kTile + b;
      double v = 0.0;
      if (q <= B.T) {
        const int i = (J + q) * kTile + a;
        if (i < nm && j < nm) { if (i >= j) v = A.C[i + static_cast<size_t>(A.ldc) * j]; }
        else if (i == j) v = 1.0;
      } else if (j < nm) {
        const int bi = (q - B.T - 1) * kTile + a;      // row nm + bi of the corner; bi == B.nbo is the rhs row
        if (bi <= B.nbo) v = A.C[(nm + bi) + static_cast<size_t>(A.ldc) * j];
      }
      dst[e] = v;
    }
  } else {
    for (int e = (blockIdx.x - ntile) * blockDim.x + threadIdx.x; e < B.ldc * B.ldc; e += (gridDim.x - ntile) * blockDim.x) {
      const int bi = e % B.ldc, bj = e / B.ldc;
      double v = 0.0;
      if (bi <= B.nbo && bj < B.nbo) { if (bi >= bj) v = A.C[(nm + bi) + static_cast<size_t>(A.ldc) * (nm + bj)]; }
      else if (bi == bj) v = 1.0;
      B.C[e] = v;
    }
  }
}
// border solution of A = [x of the separator | border solution of B]
__global__ void second_level_solution_kernel(BandSys A, BandSys B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.ldc) return;
  double v = 0.0;
  if (i < A.n_mid) v = B.x[i];
  else if (i < A.nbo) v = B.x[static_cast<size_t>(B.NT) * kTile + (i - A.n_mid)];
  A.x[static_cast<size_t>(A.NT) * kTile + i] = v;
}

// y (tangent order) from the solver's x, delta = S y, and the scalars of the step: [2] y.g_s  [3] sum D2 y^2  [4] #non-finite
__global__ void __launch_bounds__(256) finish_step_kernel(BandSys A, SchurView SV, int nt, const double* __restrict__ scale, const double* __restrict__ diag,
                                                          double inv_radius, const double* __restrict__ g, double* __restrict__ y,
                                                          double* __restrict__ delta, double* __restrict__ scal) {
  double yg = 0.0, dy = 0.0, nf = 0.0;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += gridDim.x * blockDim.x) {
    const double v = t < A.nb ? A.x[t] : t < SV.base ? A.x[static_cast<size_t>(A.NT) * kTile + (t - A.nb)] : SV.yrho[t - SV.base];
    y[t] = v;
    delta[t] = v * scale[t];
    if (!isfinite(v)) nf += 1.0;
    yg += v * g[t] * scale[t];
    dy += diag[t] * inv_radius * v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { yg += __shfl_xor_sync(FULL, yg, o); dy += __shfl_xor_sync(FULL, dy, o); nf += __shfl_xor_sync(FULL, nf, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(scal + 2, yg); atomicAdd(scal + 3, dy); if (nf != 0.0) atomicAdd(scal + 4, nf); }
}

// ---- parameter update (Program::Plus): EigenQuaternionParameterization + bounds projection -------------------------------------
__global__ void __launch_bounds__(256) plus_kernel(const FreeBlock* __restrict__ blocks, int n_blocks, const double* __restrict__ X,
                                                   const double* __restrict__ delta, double step, double sign, double* __restrict__ XC) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_blocks) return;
  const FreeBlock fb = blocks[b];
  const double* x = X + fb.off;
  double* o = XC + fb.off;
  const double* d = delta + fb.pos;
  const double f = step * sign;
  if (fb.kind == 1) {  // q+ = [sin|d|/|d| d, cos|d|] * q
    const double d0 = f * d[0], d1 = f * d[1], d2 = f * d[2];
    const double n = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    if (n > 0.0) {
      const double s = sin(n) / n;
      const Q4 r = qmul(q4(s * d0, s * d1, s * d2, cos(n)), q4(x[0], x[1], x[2], x[3]));
      o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w;
    } else { o[0] = x[0]; o[1] = x[1]; o[2] = x[2]; o[3] = x[3]; }
  } else {
    for (int c = 0; c < fb.size; ++c) {
      double v = x[c] + f * d[c];
      if (fb.kind == 2) v = fmax(v, 0.0);  // inverse depth lower bound (K/measurements/static_rscamera_measurement.h:185)
      o[c] = v;
    }
  }
}

// scal[5] += sum (x - xc)^2 ; scal[6] = max |x - xc| ; scal[7] += sum x^2   over the free blocks
__global__ void __launch_bounds__(256) diff_kernel(const FreeBlock* __restrict__ blocks, int n_blocks, const double* __restrict__ X,
                                                   const double* __restrict__ XC, double* __restrict__ scal) {
  double ss = 0.0, mx = 0.0, xx = 0.0;
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < n_blocks; b += gridDim.x * blockDim.x) {
    const FreeBlock fb = blocks[b];
    for (int c = 0; c < fb.size; ++c) {
      const double xv = X[fb.off + c], dv = xv - XC[fb.off + c];
      ss += dv * dv; mx = fmax(mx, fabs(dv)); xx += xv * xv;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { ss += __shfl_xor_sync(FULL, ss, o); xx += __shfl_xor_sync(FULL, xx, o); mx = fmax(mx, __shfl_xor_sync(FULL, mx, o)); }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(scal + 5, ss); atomicAdd(scal + 7, xx);
    atomicMax(reinterpret_cast<unsigned long long*>(scal + 6), static_cast<unsigned long long>(__double_as_longlong(mx)));
  }
}

__global__ void __launch_bounds__(256) dot_kernel(const double* __restrict__ a, const double* __restrict__ b, int n, double* __restrict__ out) {
  double s = 0.0;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) s += a[t] * b[t];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

// diagnostics (LVI_POTRF_SELFTEST): the three Cholesky warps on one tile, alone on the device, timed with clock64
__global__ void __launch_bounds__(kFacThreads, 2) potrf_selftest_kernel(const double* tile_g, long long* cycles, const double2* bg, size_t bg_n, volatile int* stop) {
  extern __shared__ __align__(128) unsigned char selftest_pad[];   // sized at launch so that every block has an SM to itself
  if (blockIdx.x != 0) {   // background: stream a large buffer through L2 until block 0 is done
    double2 acc = make_double2(0.0, 0.0);
    size_t i = (static_cast<size_t>(blockIdx.x) * 7919 * 256 + threadIdx.x) % bg_n;
    while (*stop == 0) {
#pragma unroll
      for (int u = 0; u < 8; ++u) { const double2 v = __ldcg(bg + i); acc.x += v.x; acc.y += v.y; i += 256 * 149; if (i >= bg_n) i -= bg_n; }
    }
    if (acc.x == 12345.678) cycles[7] = 1;
    return;
  }
  if (bg) __nanosleep(30000);
  // same shared-memory layout as the factor kernel's chain CTA
  FacShared& sh = *reinterpret_cast<FacShared*>(selftest_pad);
  double* sA = sh.buf[0]; double* sM = sh.buf[3]; double* sW = sh.sW; double* sR = sh.sR;
  volatile int& s_progress = sh.progress;
  const int tid = threadIdx.x;
  for (int r = 0; r < 4; ++r) {
    for (int e = tid; e < kTileElems; e += kFacThreads) sA[e] = tile_g[e];
    if (tid == 96) s_progress = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (tid < 32) warp_potrf_head(sA, 32, sM, sR, &s_progress);
    else if (tid < 64) warp_potrf_tail(sA, 32, sM, sR, &s_progress);
    else if (tid < 96) warp_inverse_cols(sM, sR, &s_progress, sW);
    __syncthreads();
    if (tid == 0) cycles[r] = clock64() - t0;
  }
  if (tid == 0) { cycles[4] = static_cast<long long>(sW[5 * kLP + 3] * 1e6); __threadfence(); *stop = 1; }
}

// ---- host side ----------------------------------------------------------------------------------------------------------
static int resident_ctas(lvi_ctx* ctx, const void* kernel, int threads, size_t smem) {
  int per_sm = 0;
  LVI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
  return std::max(1, per_sm) * ctx->sm_count;
}
static unsigned next_epoch() {
  static std::atomic<unsigned> epoch_counter{0};
  unsigned e;
  do { e = ++epoch_counter; } while (e == 0);
  return e;
}
static void band_factor_only(lvi_ctx* ctx, BandSys& A, bool allow_trace = true) {
  cudaStream_t st = ctx->stream;
  if (A.NT == 0) return;
  LVI_REQUIRE(A.work_i && A.work_d && A.ll, LVI_ERR_INVALID, "band solver workspace missing");
  A.epoch = next_epoch();   // flag value of this factorisation's flagged tile copies (never cleared)
  if (std::getenv("LVI_POTRF_SELFTEST")) {
    bool& done = ctx->ks.selftest_done;
    if (!done) {
      done = true;
      std::vector<double> h(kTileElems);
      for (int i = 0; i < 32; ++i) for (int j = 0; j < 32; ++j) h[i + 32 * j] = i == j ? 40.0 : 1.0 / (1 + i + j);
      DBuf<double> dt(kTileElems); DBuf<long long> dc(8);
      dt.upload(h.data(), kTileElems, st);
      LVI_CUDA(cudaStreamSynchronize(st));
      DBuf<int> stop(1);
      LVI_CUDA(cudaFuncSetAttribute(potrf_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
      for (int bgload = 0; bgload < 2; ++bgload) {
        stop.zero(st);
        potrf_selftest_kernel<<<bgload ? 148 : 1, kFacThreads, 120 * 1024, st>>>(dt.p, dc.p, bgload ? reinterpret_cast<const double2*>(A.ll) : nullptr,
                                                                                A.ll_count() / 2, stop.p);
        long long hc[8] = {0};
        LVI_CUDA(cudaMemcpyAsync(hc, dc.p, sizeof(hc), cudaMemcpyDeviceToHost, st));
        LVI_CUDA(cudaStreamSynchronize(st));
        std::fprintf(stderr, "[lvi] potrf selftest (head + tail + inverse on an SM of its own, %s): %lld %lld %lld %lld cycles\n",
                     bgload ? "147 other SMs streaming through L2" : "device otherwise idle", hc[0], hc[1], hc[2], hc[3]);
      }
    }
  }
  LVI_CUDA(cudaMemsetAsync(A.work_i, 0, A.work_i_count() * sizeof(int), st));
  LVI_CUDA(cudaMemsetAsync(A.work_d, 0, A.work_d_count() * sizeof(double), st));
  constexpr size_t smem = sizeof(FacShared);
  int& resident = ctx->ks.fac_resident;
  if (!resident) {
    LVI_CUDA(cudaFuncSetAttribute(band_factor_ll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    resident = resident_ctas(ctx, reinterpret_cast<const void*>(band_factor_ll_kernel), kFacThreads, smem);
  }
  const int n_chains = (A.NT0 > 0 && A.NT0 < A.NT) ? 2 : 1;
  // Column slots by which the pre-accumulation tasks (s <= 1) are queued ahead of the far tiles of their column.  A task may then wait for
  // tiles whose tasks sit up to (pre_shift - 1) slots LATER in the queue; tasks are handed out in queue order, so nothing can block as long
  // as the resident workers outnumber the live tasks of that span: (pre_shift - 1) * live_per_slot <= workers - live_per_slot / 2.
  // More lead is better up to there (C2: 3 -> 1.92 ms, 4 -> 1.87, 5 -> 1.82, 6 -> 1.80): the s <= 4 tasks run ~20 serial updates (~30 us).
  const int live_per_slot = n_chains * (A.T + 1 + std::max(1, A.RB - (A.n_mid >> kTileLog)));
  const int workers = resident - 2 * n_chains;   // the chain CTAs, and the workers that leave the chains' SMs to them
  A.pre_shift = std::max(0, std::min(6, (workers - live_per_slot / 2) / live_per_slot + 1));
  if (std::getenv("LVI_PRE_SHIFT")) A.pre_shift = std::atoi(std::getenv("LVI_PRE_SHIFT"));   // diagnostics
  {
    int e2 = std::max(A.pre_shift - 2, std::min(2, A.pre_shift)), e34 = std::max(A.pre_shift - 3, std::min(1, A.pre_shift)), prof = 0;
    if (std::getenv("LVI_EARLY2")) e2 = std::min(A.pre_shift, std::atoi(std::getenv("LVI_EARLY2")));
    if (std::getenv("LVI_EARLY34")) e34 = std::min(A.pre_shift, std::atoi(std::getenv("LVI_EARLY34")));
    if (std::getenv("LVI_LEAD_PROFILE")) prof = std::atoi(std::getenv("LVI_LEAD_PROFILE"));
    for (int s = 0; s < 64; ++s) {
      int l = s <= 1 ? A.pre_shift : s == 2 ? e2 : s <= 4 ? e34 : 0;
      if (prof == 1 && s >= 2) l = s <= A.T ? (A.pre_shift * (A.T - s + 2) + (A.T + 2) / 2) / (A.T + 2) : 0;                 // linear in the number of updates
      if (prof == 2 && s >= 2) l = s <= A.T ? std::min(A.pre_shift - 2, (A.pre_shift * (A.T - s + 2) + (A.T + 2) / 2) / (A.T + 2)) : 0;
      if (prof == 3 && s >= 2) l = s <= A.T ? std::max(0, std::min(A.pre_shift - 1, ((A.pre_shift + 1) * (A.T - s)) / A.T)) : 0;
      A.lead[s] = static_cast<unsigned char>(std::max(0, std::min(l, A.pre_shift)));
    }
  }
  const int grid = std::max(n_chains + 1, std::min(resident, (A.NT + A.pre_shift) * A.TPC + n_chains));   // chain CTAs (always resident) + workers
  const char* trace_path = allow_trace ? std::getenv("LVI_TRACE_FACTOR") : nullptr;
  if (trace_path) {  // diagnostics: per-task timestamps of ONE factorisation, dumped as uint64[ntask][8]
    const size_t n = static_cast<size_t>(A.NT) * A.TPC * 8;
    DBuf<unsigned long long> tr(n);
    tr.zero(st);
    BandSys At = A;
    At.trace = tr.p;
    LVI_LAUNCH(ctx, band_factor_ll_kernel, grid, kFacThreads, smem, At);
    std::vector<unsigned long long> h(n);
    tr.download(h.data(), n, st);
    int chain_sm[2] = {0, 0};
    LVI_CUDA(cudaMemcpyAsync(chain_sm, A.work_i + static_cast<size_t>(A.NT) * A.TPC + A.NT + 2, sizeof(chain_sm), cudaMemcpyDeviceToHost, st));
    LVI_CUDA(cudaStreamSynchronize(st));
    if (FILE* f = std::fopen(trace_path, "wb")) { const int hdr[4] = {A.NT, A.TPC, A.T, A.RB}; std::fwrite(hdr, sizeof(int), 4, f); std::fwrite(h.data(), 8, n, f); std::fclose(f); }
    std::fprintf(stderr, "[lvi] factor trace: chain CTAs on SM %d and %d, grid %d\n", chain_sm[0] - 1, chain_sm[1] - 1, grid);
    return;
  }
  LVI_LAUNCH(ctx, band_factor_ll_kernel, grid, kFacThreads, smem, A);
}
static void band_backsolve_only(lvi_ctx* ctx, BandSys& A) {
  if (A.NT == 0) return;
  LVI_REQUIRE(A.work_i && A.ll, LVI_ERR_INVALID, "band solver workspace missing");
  constexpr size_t smem = sizeof(BsShared);
  int& resident = ctx->ks.bs_resident;
  if (!resident) {
    LVI_CUDA(cudaFuncSetAttribute(band_backsolve_ll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    resident = resident_ctas(ctx, reinterpret_cast<const void*>(band_backsolve_ll_kernel), 256, smem);
  }
  A.epoch = next_epoch();   // flag value of this substitution's flagged vectors
  const int n_chains = (A.NT0 > 0 && A.NT0 < A.NT) ? 2 : 1;
  LVI_LAUNCH(ctx, band_backsolve_ll_kernel, std::max(n_chains + 1, std::min(resident, A.NT + n_chains)), 256, smem, A);
}
void init_second_level(const BandSys& A, BandSys& B) {
  B = BandSys{};
  B.nb = A.n_mid; B.nbo = A.nbo - A.n_mid;
  B.NT = (B.nb + kTile - 1) / kTile; B.NT0 = B.NT; B.n_mid = 0;
  B.T = B.NT - 1;                                   // dense
  B.RB = (B.nbo + 1 + kTile - 1) / kTile; B.TPC = B.T + 1 + B.RB; B.ldc = B.RB * kTile;
}
static void launch_corner_solve(lvi_ctx* ctx, const BandSys& S) {
  const size_t need = static_cast<size_t>(S.nbo + 1) * (S.nbo + 2) * sizeof(double);
  const bool in_smem = need <= 160 * 1024;
  size_t& attr = ctx->ks.corner_attr;
  if (in_smem && need > attr) { LVI_CUDA(cudaFuncSetAttribute(corner_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(need))); attr = need; }
  LVI_LAUNCH(ctx, corner_solve_kernel, 1, 256, in_smem ? need : 0, S, in_smem ? 1 : 0);
}
// everything after the factorisation of A: corner (directly, or through the second-level system), then the back substitution
static void band_solve_only(lvi_ctx* ctx, BandSys& A, BandSys& A2) {
  if (A.n_mid > 0) {
    const int ntile = A2.NT * A2.TPC;
    LVI_LAUNCH(ctx, corner_to_second_level_kernel, ntile + 4, 256, 0, A, A2);
    band_factor_only(ctx, A2, false);
    launch_corner_solve(ctx, A2);
    band_backsolve_only(ctx, A2);
    LVI_LAUNCH(ctx, second_level_solution_kernel, (A.ldc + 255) / 256, 256, 0, A, A2);
  } else {
    launch_corner_solve(ctx, A);
  }
  band_backsolve_only(ctx, A);
}
void band_factor_solve(lvi_ctx* ctx, BandSys& A, BandSys& A2) {
  LVI_CUDA(cudaMemsetAsync(A.fail, 0, sizeof(int), ctx->stream));
  band_factor_only(ctx, A);
  band_solve_only(ctx, A, A2);
}

struct Scalars {  // mirrors p->scal (+ the factorisation failure flag)
  double cost, cand_cost, yg, d2y2, nonfinite, diff_ss, diff_max, xnorm_ss, gd;
  int failed;
};

// gather / scatter of the structurally non-zero tiles of H around the all-reduce (one CTA per tile)
__global__ void __launch_bounds__(256) pack_tiles_kernel(double* __restrict__ tiles, const int* __restrict__ map, double* __restrict__ packed, int unpack) {
  double2* t = reinterpret_cast<double2*>(tiles + static_cast<size_t>(map[blockIdx.x]) * kTileElems);
  double2* q = reinterpret_cast<double2*>(packed + static_cast<size_t>(blockIdx.x) * kTileElems);
  for (int e = threadIdx.x; e < kTileElems / 2; e += 256) {
    if (unpack) t[e] = q[e];
    else q[e] = t[e];
  }
}

static void allreduce_sum(lvi_ctx* ctx, double* buf, size_t count) {
  if (ctx->world <= 1 || count == 0) return;
  ncclResult_t r = nccl().AllReduce(buf, buf, count, ncclDouble, ncclSum, static_cast<ncclComm_t>(ctx->nccl), ctx->stream);
  LVI_REQUIRE(r == ncclSuccess, LVI_ERR_NCCL, std::string("ncclAllReduce: ") + nccl().GetErrorString(r));
}

static void read_scalars(lvi_problem* p, Scalars& s) {
  cudaStream_t st = p->ctx->stream;
  LVI_CUDA(cudaMemcpyAsync(p->h_scal, p->scal.p, 15 * sizeof(double), cudaMemcpyDeviceToHost, st));
  LVI_CUDA(cudaMemcpyAsync(p->h_scal + 15, p->fail.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  LVI_CUDA(cudaStreamSynchronize(st));   // the ONLY host wait of the LM loop: twice per iteration
  const double* h = p->h_scal;
  std::memcpy(&s.failed, p->h_scal + 15, sizeof(int));
  s.cost = h[0]; s.cand_cost = h[1]; s.yg = h[2]; s.d2y2 = h[3]; s.nonfinite = h[4]; s.diff_ss = h[5]; s.diff_max = h[6]; s.xnorm_ss = h[7]; s.gd = h[8];
}

// residuals + Jacobians + normal equations at X (+ NCCL all-reduce of {H, g, cost} across the data-parallel ranks)
static void linearize(lvi_problem* p) {
  problem_linearize(p, nullptr);
  lvi_ctx* ctx = p->ctx;
  if (ctx->world > 1 && p->p2p.active) {   // fused reduction over NVLink peer memory (p2p.cu): no pack / unpack, peers' tiles read in place
    p2p_reduce(p);
  } else if (ctx->world > 1) {   // ONE all-reduce per iteration over [non-zero tiles | corner | g | Schur rows | Schur diagonal | cost]
    cudaStream_t st = ctx->stream;
    double* buf = p->pack_buf.p;
    if (p->n_pack > 0) LVI_LAUNCH(ctx, pack_tiles_kernel, p->n_pack, 256, 0, p->H_tiles.p, p->pack_map.p, buf, 0);
    struct Seg { double* ptr; size_t n; };
    const Seg segs[5] = {{p->H_C.p, p->H_C.n}, {p->g.p, p->g.n}, {p->Hrx.p, p->Hrx.n}, {p->Hrr.p, p->Hrr.n}, {p->scal.p, 1}};
    size_t off = static_cast<size_t>(p->n_pack) * kTileElems;
    for (const Seg& s : segs) { if (s.n) LVI_CUDA(cudaMemcpyAsync(buf + off, s.ptr, s.n * sizeof(double), cudaMemcpyDeviceToDevice, st)); off += s.n; }
    allreduce_sum(ctx, buf, off);
    if (p->n_pack > 0) LVI_LAUNCH(ctx, pack_tiles_kernel, p->n_pack, 256, 0, p->H_tiles.p, p->pack_map.p, buf, 1);
    off = static_cast<size_t>(p->n_pack) * kTileElems;
    for (const Seg& s : segs) { if (s.n) LVI_CUDA(cudaMemcpyAsync(s.ptr, buf + off, s.n * sizeof(double), cudaMemcpyDeviceToDevice, st)); off += s.n; }
  }
}
static void trial_cost(lvi_problem* p, const double* x_d, double* cost_d, bool active, bool inactive) {
  problem_cost(p, x_d, cost_d, active, inactive);
  allreduce_sum(p->ctx, cost_d, 1);
}

static inline int blocks_for(int n) { return std::max(1, (n + 255) / 256); }

static size_t max_row_len(const lvi_problem* p) {
  size_t m = 0;
  for (size_t l = 0; l + 1 < p->L.row_start.size(); ++l) m = std::max<size_t>(m, p->L.row_start[l + 1] - p->L.row_start[l]);
  return m;
}
static void schur_eliminate(lvi_problem* p, double inv_radius) {
  if (p->L.n_rho == 0) return;
  if (p->schur_gather) {   // rows only; the products go through the tile gather (fp64 tensor pipe, single-owner tiles)
    const size_t smem_rows = max_row_len(p) * 8 + 16;
    LVI_REQUIRE(smem_rows <= 48 * 1024, LVI_ERR_INVALID, "Schur row too long");
    LVI_LAUNCH(p->ctx, schur_rows_kernel, p->L.n_rho, 128, smem_rows, p->A, p->schur, p->scale.p, p->diag.p, inv_radius, p->g.p, p->schur_boff.p, p->schur_rows.p);
    assemble_schur_gather(p);
    return;
  }
  size_t& attr = p->ctx->ks.schur_attr;
  const size_t smem = max_row_len(p) * 12 + 16;
  if (smem > attr) { LVI_CUDA(cudaFuncSetAttribute(schur_eliminate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))); attr = smem; }
  LVI_LAUNCH(p->ctx, schur_eliminate_kernel, p->L.n_rho, 128, smem, p->A, p->schur, p->scale.p, p->diag.p, inv_radius, p->g.p);
}
static void schur_back(lvi_problem* p, double inv_radius) {
  if (p->L.n_rho == 0) return;
  LVI_LAUNCH(p->ctx, schur_back_kernel, (p->L.n_rho + 3) / 4, 128, 0, p->A, p->schur, p->scale.p, p->diag.p, inv_radius, p->g.p);
}

static void compute_step(lvi_problem* p, double radius) {  // A = S H S + D^2 ; factor ; solve ; y, delta, scalars
  lvi_ctx* ctx = p->ctx;
  const int ntile = p->A.NT * p->A.TPC;
  const int corner_ctas = std::max(1, std::min(64, (p->A.ldc * p->A.ldc + 255) / 256));
  LVI_LAUNCH(ctx, build_system_kernel, ntile + corner_ctas, 256, 0, p->H, p->A, p->scale.p, p->diag.p, 1.0 / radius, p->g.p);
  schur_eliminate(p, 1.0 / radius);
  band_factor_solve(ctx, p->A, p->A2);
  schur_back(p, 1.0 / radius);
  LVI_CUDA(cudaMemsetAsync(p->scal.p + 2, 0, 3 * sizeof(double), ctx->stream));
  LVI_LAUNCH(ctx, finish_step_kernel, std::min(blocks_for(p->nt), ctx->sm_count * 4), 256, 0, p->A, p->schur, p->nt, p->scale.p, p->diag.p, 1.0 / radius, p->g.p,
             p->y.p, p->delta.p, p->scal.p);
}

static void apply_plus(lvi_problem* p, const double* delta_d, double step, double sign) {
  cudaStream_t st = p->ctx->stream;
  LVI_CUDA(cudaMemcpyAsync(p->XC.p, p->X.p, sizeof(double) * p->nx, cudaMemcpyDeviceToDevice, st));
  if (p->n_blocks) LVI_LAUNCH(p->ctx, plus_kernel, blocks_for(p->n_blocks), 256, 0, p->blocks.p, p->n_blocks, p->X.p, delta_d, step, sign, p->XC.p);
}
static void diff_norms(lvi_problem* p) {  // scal[5..7]
  LVI_CUDA(cudaMemsetAsync(p->scal.p + 5, 0, 3 * sizeof(double), p->ctx->stream));
  if (p->n_blocks) LVI_LAUNCH(p->ctx, diff_kernel, std::min(blocks_for(p->n_blocks), p->ctx->sm_count * 4), 256, 0, p->blocks.p, p->n_blocks, p->X.p, p->XC.p, p->scal.p);
}

static void refresh_diag(lvi_problem* p, DBuf<double>& dH, const lvi_solve_options& o) {
  LVI_LAUNCH(p->ctx, diag_kernel, blocks_for(p->nt), 256, 0, p->H, p->schur, p->nt, dH.p);
  LVI_LAUNCH(p->ctx, lm_diag_kernel, blocks_for(p->nt), 256, 0, dH.p, p->scale.p, p->nt, o.min_lm_diagonal, o.max_lm_diagonal, p->diag.p);
}

static void solve_lm(lvi_problem* p, const lvi_solve_options& o, lvi_solve_summary& S, lvi_iteration_callback cb = nullptr, void* cb_user = nullptr,
                     int cb_update_state = 0) {
  const auto T0 = std::chrono::steady_clock::now();
  std::memset(&S, 0, sizeof(S));
  lvi_ctx* ctx = p->ctx;
  cudaStream_t st = ctx->stream;
  const bool timing = std::getenv("LVI_TIME_SOLVE") != nullptr;   // diagnostics: host wall time between the loop's wait points, on stderr
  auto Tl = T0;
  auto lap = [&](const char* what, int it_no = -1) {
    if (!timing) return;
    const auto T1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[lvi] solve: %-24s %3d %9.3f ms\n", what, it_no, std::chrono::duration<double, std::milli>(T1 - Tl).count());
    Tl = T1;
  };
  problem_ensure_solver_buffers(p);
  lap("solver buffers");
  const int nt = p->nt;
  DBuf<double> dH(std::max(nt, 1));
  Scalars sc{};
  // phase timing with CUDA events that are only read back after the loop: the loop itself never waits for them
  struct Span { cudaEvent_t a, b; float* accu; };
  std::vector<Span> spans;
  float tj = 0, tl = 0;
  ctx->timing_used = 0;   // events come from the context's pool (created once, reused by every solve)
  auto tic = [&](float& accu) {
    Span sp{ctx->timing_event(), ctx->timing_event(), &accu};
    LVI_REQUIRE(sp.a && sp.b, LVI_ERR_CUDA, "cudaEventCreate failed");
    LVI_CUDA(cudaEventRecord(sp.a, st));
    spans.push_back(sp);
  };
  auto toc = [&] { LVI_CUDA(cudaEventRecord(spans.back().b, st)); };

  // fixed cost of the residual blocks whose parameter blocks are all constant (dropped from the reduced program)
  trial_cost(p, p->X.p, p->scal.p + 1, false, true);
  read_scalars(p, sc);
  const double fixed_cost = sc.cand_cost;
  lap("fixed cost");
  tic(tj); linearize(p); toc();
  read_scalars(p, sc);
  lap("first linearize");
  double x_cost = sc.cost;
  S.initial_cost = x_cost + fixed_cost; S.fixed_cost = fixed_cost;
  S.num_residual_blocks = p->L.n_res_blocks; S.num_residuals = p->L.n_res; S.num_effective_parameters = nt;
  S.band_width = p->L.bw; S.border_width = p->L.nbo;
  // Jacobi scaling at iteration 0
  LVI_LAUNCH(ctx, diag_kernel, blocks_for(nt), 256, 0, p->H, p->schur, nt, dH.p);
  LVI_LAUNCH(ctx, scale_init_kernel, blocks_for(nt), 256, 0, dH.p, nt, o.jacobi_scaling, p->scale.p);
  auto grad_max_norm = [&]() {  // || x - Plus(x, -g) ||_inf (with bounds projection)
    apply_plus(p, p->g.p, 1.0, -1.0);
    diff_norms(p);
    read_scalars(p, sc);
    return sc.diff_max;
  };
  double gmax = grad_max_norm();
  double x_norm = std::sqrt(sc.xnorm_ss);
  double radius = o.initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  int it = 0, invalid = 0;
  int cb_verdict = 0;   // what the last iteration callback returned
  auto log_iter = [&](double cost, double change, double gm, double step, bool ok) {
    if (S.n_log < LVI_MAX_ITER_LOG) {
      const int k = S.n_log++;
      S.log_cost[k] = cost; S.log_cost_change[k] = change; S.log_gradient_max_norm[k] = gm; S.log_step_norm[k] = step; S.log_radius[k] = radius; S.log_successful[k] = ok;
    }
    if (cb) {   // ceres::IterationCallback (TrajectoryEstimator::AddCallback, K/trajectory_estimator.h:88-94)
      if (cb_update_state) problem_download_params(p);   // update_state_every_iteration: the caller's arrays hold the current iterate
      lvi_iteration_summary its{};
      its.iteration = it; its.step_is_successful = ok ? 1 : 0; its.cost = cost; its.cost_change = change; its.gradient_max_norm = gm; its.step_norm = step;
      its.trust_region_radius = radius;
      cb_verdict = cb(&its, cb_user);
    }
  };
  log_iter(x_cost + fixed_cost, 0, gmax, 0, true);
  S.termination_type = LVI_NO_CONVERGENCE;
  if (o.verbose) std::printf("[lvi] iter %3d cost %.9e |g| %.3e\n", 0, x_cost, gmax);
  bool done = gmax <= o.gradient_tolerance;
  if (done) S.termination_type = LVI_CONVERGENCE;
  while (!done) {
    if (cb_verdict == 1) { S.termination_type = LVI_USER_FAILURE; break; }
    if (cb_verdict == 2) { S.termination_type = LVI_USER_SUCCESS; break; }
    if (it >= o.max_num_iterations) break;
    ++it;
    if (!reuse_diagonal) refresh_diag(p, dH, o);
    // the whole trial is queued before the host looks at anything: step, candidate point, its cost, the Armijo slope, step norms
    tic(tl);
    compute_step(p, radius);
    toc();
    apply_plus(p, p->delta.p, 1.0, 1.0);
    tic(tj);
    trial_cost(p, p->XC.p, p->scal.p + 1, true, false);
    toc();
    if (p->L.constrained) {
      LVI_CUDA(cudaMemsetAsync(p->scal.p + 8, 0, sizeof(double), st));
      LVI_LAUNCH(ctx, dot_kernel, std::min(blocks_for(nt), ctx->sm_count * 4), 256, 0, p->g.p, p->delta.p, nt, p->scal.p + 8);
    }
    diff_norms(p);
    read_scalars(p, sc);
    lap("step + trial cost", it);
    reuse_diagonal = true;
    bool step_valid = !sc.failed && sc.nonfinite == 0.0;
    // model_cost_change = -(J y).(r + J y / 2) = -y.g_s - y^T H_s y / 2, with H_s y = -g_s - D^2 y
    const double model_cost_change = -0.5 * sc.yg + 0.5 * sc.d2y2;
    if (step_valid && !(model_cost_change > 0.0)) step_valid = false;
    if (!step_valid) {  // HandleInvalidStep (the candidate evaluated above is simply discarded)
      ++invalid;
      if (invalid >= o.max_num_consecutive_invalid_steps) { S.termination_type = LVI_FAILURE; break; }
      radius *= 0.5;
      ++S.num_unsuccessful_steps;
      log_iter(x_cost + fixed_cost, 0, gmax, 0, false);
      continue;
    }
    invalid = 0;
    double cand_cost = sc.cand_cost;
    if (p->L.constrained) {  // projected step: Armijo check along delta (ceres DoLineSearch)
      const double gd = sc.gd;
      double step = 1.0;
      int ls = 0;
      while (cand_cost > x_cost + 1e-4 * step * gd && ls < 20 && step > 1e-9) {
        double ns = -gd * step * step / (2.0 * (cand_cost - x_cost - gd * step));
        ns = std::min(std::max(ns, 1e-3 * step), 0.6 * step);
        step = ns; ++ls;
        apply_plus(p, p->delta.p, step, 1.0);
        trial_cost(p, p->XC.p, p->scal.p + 1, true, false);
        diff_norms(p);
        read_scalars(p, sc);
        cand_cost = sc.cand_cost;
      }
    }
    const double step_norm = std::sqrt(sc.diff_ss);
    const double cost_change = x_cost - cand_cost;
    if (step_norm <= o.parameter_tolerance * (x_norm + o.parameter_tolerance)) {
      S.termination_type = LVI_CONVERGENCE; log_iter(x_cost + fixed_cost, cost_change, gmax, step_norm, false); break;
    }
    if (std::fabs(cost_change) <= o.function_tolerance * x_cost) {
      S.termination_type = LVI_CONVERGENCE; log_iter(x_cost + fixed_cost, cost_change, gmax, step_norm, false); break;
    }
    const double rel = cost_change / model_cost_change;
    if (rel > o.min_relative_decrease) {  // HandleSuccessfulStep
      LVI_CUDA(cudaMemcpyAsync(p->X.p, p->XC.p, sizeof(double) * p->nx, cudaMemcpyDeviceToDevice, st));
      x_cost = cand_cost;
      tic(tj); linearize(p); toc();
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rel - 1.0, 3));
      radius = std::min(o.max_trust_region_radius, radius);
      decrease_factor = 2.0; reuse_diagonal = false;
      ++S.num_successful_steps;
      gmax = grad_max_norm();
      lap("accept: linearize", it);
      x_norm = std::sqrt(sc.xnorm_ss);
      log_iter(x_cost + fixed_cost, cost_change, gmax, step_norm, true);
      if (o.verbose) std::printf("[lvi] iter %3d cost %.9e change %.3e |g| %.3e |step| %.3e radius %.3e\n", it, x_cost, cost_change, gmax, step_norm, radius);
      if (gmax <= o.gradient_tolerance) { S.termination_type = LVI_CONVERGENCE; break; }
    } else {  // HandleUnsuccessfulStep
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      ++S.num_unsuccessful_steps;
      log_iter(x_cost + fixed_cost, cost_change, gmax, step_norm, false);
      if (o.verbose) std::printf("[lvi] iter %3d REJECTED cost %.9e change %.3e radius %.3e\n", it, cand_cost, cost_change, radius);
    }
    if (radius <= o.min_trust_region_radius) { S.termination_type = LVI_CONVERGENCE; break; }
  }
  problem_download_params(p);
  LVI_CUDA(cudaStreamSynchronize(st));
  lap("download");
  for (Span& sp : spans) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) *sp.accu += ms;
  }
  (void)cudaGetLastError();
  S.num_iterations = it;
  S.final_cost = x_cost + fixed_cost;
  S.time_jacobian_ms = tj; S.time_linear_solve_ms = tl;
  S.time_total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T0).count();
}

}  // namespace lvi

using namespace lvi;

extern "C" {

int lvi_problem_solve(lvi_problem* p, const lvi_solve_options* opt, lvi_solve_summary* summary) {
  return guarded([&] {
    LVI_REQUIRE(p && opt && summary, LVI_ERR_INVALID, "lvi_problem_solve: null argument");
    activate(p->ctx);
    solve_lm(p, *opt, *summary);
  });
}

int lvi_problem_solve_cb(lvi_problem* p, const lvi_solve_options* opt, lvi_solve_summary* summary, lvi_iteration_callback cb, void* user, int update_state) {
  return guarded([&] {
    LVI_REQUIRE(p && opt && summary, LVI_ERR_INVALID, "lvi_problem_solve_cb: null argument");
    activate(p->ctx);
    solve_lm(p, *opt, *summary, cb, user, update_state);
  });
}

// One LM iteration's worth of work, `iters` times, without acceptance: linearise, damped system, factor + solve, candidate, trial cost.
// ms_per_phase[6]: linearize | build_system | band_factor | corner + backsolve + finish | candidate + trial cost | whole loop / iters
int lvi_problem_bench_iterations(lvi_problem* p, int iters, float* ms_per_phase) {
  return guarded([&] {
    LVI_REQUIRE(p && iters > 0, LVI_ERR_INVALID, "lvi_problem_bench_iterations: bad argument");
    activate(p->ctx);
    lvi_ctx* ctx = p->ctx;
    cudaStream_t st = ctx->stream;
    problem_ensure_solver_buffers(p);
    lvi_solve_options o;
    lvi_solve_options_default(&o);
    DBuf<double> dH(std::max(p->nt, 1));
    cudaEvent_t ev[6], t0, t1;
    for (auto& e : ev) LVI_CUDA(cudaEventCreate(&e));
    LVI_CUDA(cudaEventCreate(&t0)); LVI_CUDA(cudaEventCreate(&t1));
    float acc[5] = {0, 0, 0, 0, 0};
    linearize(p);
    LVI_LAUNCH(ctx, diag_kernel, blocks_for(p->nt), 256, 0, p->H, p->schur, p->nt, dH.p);
    LVI_LAUNCH(ctx, scale_init_kernel, blocks_for(p->nt), 256, 0, dH.p, p->nt, 1, p->scale.p);
    const double inv_radius = 1.0 / o.initial_trust_region_radius;
    LVI_CUDA(cudaStreamSynchronize(st));
    LVI_CUDA(cudaEventRecord(t0, st));
    for (int it = 0; it < iters; ++it) {
      LVI_CUDA(cudaEventRecord(ev[0], st));
      linearize(p);
      LVI_CUDA(cudaEventRecord(ev[1], st));
      refresh_diag(p, dH, o);
      const int ntile = p->A.NT * p->A.TPC;
      const int corner_ctas = std::max(1, std::min(64, (p->A.ldc * p->A.ldc + 255) / 256));
      LVI_LAUNCH(ctx, build_system_kernel, ntile + corner_ctas, 256, 0, p->H, p->A, p->scale.p, p->diag.p, inv_radius, p->g.p);
      schur_eliminate(p, inv_radius);
      LVI_CUDA(cudaEventRecord(ev[2], st));
      LVI_CUDA(cudaMemsetAsync(p->A.fail, 0, sizeof(int), st));
      band_factor_only(ctx, p->A);
      LVI_CUDA(cudaEventRecord(ev[3], st));
      band_solve_only(ctx, p->A, p->A2);
      schur_back(p, inv_radius);
      LVI_CUDA(cudaMemsetAsync(p->scal.p + 2, 0, 3 * sizeof(double), st));
      LVI_LAUNCH(ctx, finish_step_kernel, std::min(blocks_for(p->nt), ctx->sm_count * 4), 256, 0, p->A, p->schur, p->nt, p->scale.p, p->diag.p, inv_radius, p->g.p,
                 p->y.p, p->delta.p, p->scal.p);
      LVI_CUDA(cudaEventRecord(ev[4], st));
      apply_plus(p, p->delta.p, 1.0, 1.0);
      trial_cost(p, p->XC.p, p->scal.p + 1, true, false);
      diff_norms(p);
      LVI_CUDA(cudaEventRecord(ev[5], st));
      Scalars sc;
      read_scalars(p, sc);  // the per-iteration host decision point of the LM loop
      for (int k = 0; k < 5; ++k) { float ms = 0; LVI_CUDA(cudaEventElapsedTime(&ms, ev[k], ev[k + 1])); acc[k] += ms; }
    }
    LVI_CUDA(cudaEventRecord(t1, st));
    LVI_CUDA(cudaEventSynchronize(t1));
    float total = 0;
    LVI_CUDA(cudaEventElapsedTime(&total, t0, t1));
    for (auto& e : ev) cudaEventDestroy(e);
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    if (ms_per_phase) { for (int k = 0; k < 5; ++k) ms_per_phase[k] = acc[k] / iters; ms_per_phase[5] = total / iters; }
  });
}

// out[8]: band dims, border dims, half bandwidth, block columns NT, sub-diagonal tile rows T, border tile rows RB, chain1_start, n_pad
int lvi_problem_layout(lvi_problem* p, int32_t* out) {
  return guarded([&] {
    LVI_REQUIRE(p && out, LVI_ERR_INVALID, "lvi_problem_layout: null argument");
    problem_ensure_solver_buffers(p);
    out[0] = p->L.nb; out[1] = p->L.nbo; out[2] = p->L.bw; out[3] = p->A.NT; out[4] = p->A.T; out[5] = p->A.RB; out[6] = p->L.chain1_start; out[7] = p->L.n_pad;
  });
}

// test hook: solve a dense SPD system given in band+border form through the tile solver.
//   A_dense [n x n] row-major symmetric, n = nb + nbo, entries outside the band (|i-j| > bw within the first nb) must be zero.
int lvi_band_solve_dense(lvi_ctx* ctx, int nb, int nbo, int bw, int chain1_start, int n_mid, const double* A_dense, const double* rhs, double* x_out) {
  return guarded([&] {
    LVI_REQUIRE(ctx && A_dense && rhs && x_out && nb >= 0 && nbo >= 0 && nb + nbo > 0, LVI_ERR_INVALID, "lvi_band_solve_dense: bad argument");
    LVI_REQUIRE((chain1_start == nb || chain1_start % kTile == 0) && chain1_start >= 0 && chain1_start <= nb && n_mid >= 0 && n_mid <= nbo, LVI_ERR_INVALID,
                "lvi_band_solve_dense: chain1_start must be a multiple of 32 in [0, nb], n_mid in [0, nbo]");
    activate(ctx);
    cudaStream_t st = ctx->stream;
    BandSys S{};
    S.nb = nb; S.nbo = nbo;
    S.NT = (nb + kTile - 1) / kTile;
    S.NT0 = chain1_start < nb ? chain1_start / kTile : S.NT;
    S.n_mid = n_mid;
    S.T = S.NT > 0 ? std::min(S.NT - 1, (bw + kTile - 1) / kTile) : 0;
    S.RB = (nbo + 1 + kTile - 1) / kTile; S.TPC = S.T + 1 + S.RB; S.ldc = S.RB * kTile;
    LVI_REQUIRE(S.ldc <= 1024, LVI_ERR_INVALID, "lvi_band_solve_dense: border too wide");
    const size_t ntile = static_cast<size_t>(S.NT) * S.TPC;
    std::vector<double> ht(std::max<size_t>(ntile * kTileElems, 1), 0.0), hc(static_cast<size_t>(S.ldc) * S.ldc, 0.0);
    const int n = nb + nbo;
    auto at = [&](int i, int j) { return A_dense[static_cast<size_t>(i) * n + j]; };
    for (int J = 0; J < S.NT; ++J)
      for (int q = 0; q < S.TPC; ++q)
        for (int b = 0; b < kTile; ++b)
          for (int a = 0; a < kTile; ++a) {
            const int j = J * kTile + b;
            double v = 0.0;
            if (q <= S.T) {
              const int i = (J + q) * kTile + a;
              if (J + q >= S.NT) continue;
              if (J < S.NT0 && J + q >= S.NT0) v = 0.0;  // the two chains never couple directly
              else if (i < nb && j < nb) { if (i >= j) v = at(i, j); }
              else if (i == j) v = 1.0;
            } else if (j < nb) {
              const int bi = (q - S.T - 1) * kTile + a;
              if (bi < nbo) v = at(nb + bi, j);
              else if (bi == nbo) v = rhs[j];
            }
            ht[(static_cast<size_t>(J) * S.TPC + q) * kTileElems + b * kTile + a] = v;
          }
    for (int bj = 0; bj < S.ldc; ++bj)
      for (int bi = 0; bi < S.ldc; ++bi) {
        double v = 0.0;
        if (bi < nbo && bj < nbo) { if (bi >= bj) v = at(nb + bi, nb + bj); }
        else if (bi == nbo && bj < nbo) v = rhs[nb + bj];
        else if (bi == bj) v = 1.0;
        hc[bi + static_cast<size_t>(S.ldc) * bj] = v;
      }
    DBuf<double> tiles(ht.size()), C(hc.size()), Linv(std::max<size_t>(static_cast<size_t>(S.NT) * kTileElems, 1)), x(static_cast<size_t>(S.NT) * kTile + S.ldc);
    DBuf<int> fail(4);
    tiles.upload(ht.data(), ht.size(), st); C.upload(hc.data(), hc.size(), st);
    S.tiles = tiles.p; S.C = C.p; S.Linv = Linv.p; S.x = x.p; S.fail = fail.p;
    DBuf<int> wi(S.work_i_count());
    DBuf<double> wd(std::max<size_t>(S.work_d_count(), 1));
    S.work_i = wi.p; S.work_d = wd.p;
    DBuf<unsigned long long> ll1(std::max<size_t>(S.ll_count(), 1)), ll2;
    ll1.zero(ctx->stream);
    S.ll = ll1.p; S.epoch = 0;
    BandSys S2{};
    DBuf<double> t2, c2, l2, x2v, wd2;
    DBuf<int> wi2;
    if (n_mid > 0) {
      init_second_level(S, S2);
      t2.alloc(static_cast<size_t>(S2.NT) * S2.TPC * kTileElems); c2.alloc(static_cast<size_t>(S2.ldc) * S2.ldc);
      l2.alloc(static_cast<size_t>(S2.NT) * kTileElems); x2v.alloc(static_cast<size_t>(S2.NT) * kTile + S2.ldc);
      wi2.alloc(S2.work_i_count()); wd2.alloc(S2.work_d_count());
      S2.tiles = t2.p; S2.C = c2.p; S2.Linv = l2.p; S2.x = x2v.p; S2.fail = fail.p; S2.work_i = wi2.p; S2.work_d = wd2.p;
      ll2.alloc(std::max<size_t>(S2.ll_count(), 1)); ll2.zero(ctx->stream);
      S2.ll = ll2.p; S2.epoch = 0;
    }
    band_factor_solve(ctx, S, S2);
    std::vector<double> hx(x.n);
    int hf = 0;
    x.download(hx.data(), x.n, st);
    LVI_CUDA(cudaMemcpyAsync(&hf, fail.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    LVI_CUDA(cudaStreamSynchronize(st));
    LVI_REQUIRE(hf == 0, LVI_ERR_NUMERIC, "Cholesky breakdown");
    for (int t = 0; t < n; ++t) x_out[t] = t < nb ? hx[t] : hx[static_cast<size_t>(S.NT) * kTile + (t - nb)];
  });
}

}  // extern "C"
