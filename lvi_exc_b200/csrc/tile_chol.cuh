// 32x32 fp64 Cholesky + inverse of the factor for the diagonal blocks of the band factorisation (solver.cu).
//
// The diagonal block is the serial pivot chain of the whole factorisation, so this code is written for latency.
// What was measured on B200 (tools/bench_potrf.cu, tools/bench_lat.cu, tools/bench_tput.cu; one warp, 1965 MHz):
//   * a warp issues at most one DFMA every 4.2 cycles (8.2 cycles dependent latency); MUFU.RSQ64H seed 17 cycles; rsqrt() 68 cycles;
//   * a CTA-wide shared-memory Cholesky with one barrier per pivot: 10.5 us; a one-warp shared-memory loop: 30+ us;
//   * one warp, rows in registers, fully unrolled, Cholesky then inverse: 3.9 us; with the inverse on a second warp running
//     one chunk of pivots behind: 2.7 us -- the version below.
// Warp A factors: lane a owns row a in registers; column j is published to shared memory, which is also how the rank-1 update broadcasts
// it.  Each lane keeps its own diagonal element up to date (dg -= l^2), so the next pivot costs ONE shuffle after the column scale.
// Warp B builds W = L^-1 column-parallel, right-looking, from the published columns: step t needs column t of L and 1/L(t,t) only.
//   sLc[j*33 + a] = L(a,j) (zero above the diagonal), sRinv[j] = 1/L(j,j), *progress = number of published columns.
#pragma once
#include <cuda_runtime.h>

namespace lvi {

constexpr unsigned kFullWarp = 0xffffffffu;
constexpr int kCholLd = 33;     // padded leading dimension of the shared-memory tiles
constexpr int kCholChunk = 8;   // pivots per progress publication

// Publication of "columns [0, n) of L are in shared memory" to the follower warps.  The column stores and this store are shared-memory
// stores of ONE warp in program order behind a __syncwarp, and the load/store unit performs a warp's shared-memory accesses in order, so a
// compiler barrier is all that is needed.  NOT __threadfence_block(): a membar also waits for the thread's outstanding GLOBAL stores, and
// in the factor kernel the publishing lane has 24 KB of tile stores in flight from the previous column -- measured: +1.7 us per block.
__device__ __forceinline__ void publish_progress(volatile int* progress, int n) {
  asm volatile("" ::: "memory");
  *progress = n;
}

// 1/sqrt(d) for a positive normal d: 20-bit hardware seed and one third-order step y (1 + e/2 + 3 e^2/8), e = 1 - d y^2  (error < 2^-55)
__device__ __forceinline__ double rsqrt_pos(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double t = d * y;
  const double e = fma(-t, y, 1.0);
  const double p = fma(0.375, e, 0.5);
  return fma(y, p * e, y);
}

// ---- the factorisation itself, split over two warps ----------------------------------------------------------------------------
// A single warp issues one DFMA every 4.2 cycles, and the rank-1 updates of a 32x32 Cholesky are ~500 of them on top of the pivot
// chain, so one warp is issue-bound (2.55 us).  Warp "head" therefore owns columns 0..15 (all 32 rows) and runs pivots 0..15 updating only
// its own columns; warp "tail" owns columns 16..31, applies the head's published columns to them as they appear (a follower, like the
// inverse warp) and then runs pivots 16..31 itself.  Nothing moves between the two except the published columns of L.
constexpr int kCholPub = 4;   // pivots per progress publication of the factoring warps

__device__ __forceinline__ bool warp_potrf_head(const double* tile, int ld, double* sLc, double* sRinv, volatile int* progress) {
  const int a = threadIdx.x & 31;
  double A[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) A[c] = tile[a + ld * c];
  double dg = a < 16 ? tile[a + ld * a] : 1.0;
  bool bad = false;
  double d = __shfl_sync(kFullWarp, dg, 0);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (!(d > 1e-290) || !(d < 1e290)) { bad = true; d = 1.0; }   // not positive definite (or NaN / overflow): flagged, keeps running
    const double ri = rsqrt_pos(d);
    const double l = (a == j ? d : A[j]) * ri;
    dg = fma(-l, l, dg);
    if (j < 15) d = __shfl_sync(kFullWarp, dg, j + 1);
    sLc[j * kCholLd + a] = (a >= j) ? l : 0.0;
    if (a == j) sRinv[j] = ri;
    __syncwarp();
    if ((j % kCholPub) == kCholPub - 1 && a == 0) publish_progress(progress, j + 1);
#pragma unroll
    for (int c = j + 1; c < 16; ++c) A[c] = fma(-l, sLc[j * kCholLd + c], A[c]);
  }
  return !bad;
}

__device__ __forceinline__ bool warp_potrf_tail(const double* tile, int ld, double* sLc, double* sRinv, volatile int* progress,
                                             unsigned long long* stamp = nullptr) {
  const int a = threadIdx.x & 31;
  double A[16];   // columns 16..31
#pragma unroll
  for (int c = 0; c < 16; ++c) A[c] = tile[a + ld * (16 + c)];
  double dg = a >= 16 ? tile[a + ld * a] : 1.0;
#pragma unroll
  for (int t = 0; t < 16; ++t) {   // follow the head: A(:, 16..31) -= L(:, t) L(16..31, t)^T
    if ((t % kCholPub) == 0) {
      while (*progress < t + kCholPub) {}
      __syncwarp();
    }
    const double la = sLc[t * kCholLd + a];
#pragma unroll
    for (int c = 0; c < 16; ++c) A[c] = fma(-la, sLc[t * kCholLd + 16 + c], A[c]);
    dg = fma(-la, la, dg);
  }
  if (stamp) {   // diagnostics: when the follower part ended
    if (a == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); *stamp = t; }
    __syncwarp();
  }
  bool bad = false;
  double d = __shfl_sync(kFullWarp, dg, 16);
#pragma unroll
  for (int j = 16; j < 32; ++j) {
    if (!(d > 1e-290) || !(d < 1e290)) { bad = true; d = 1.0; }
    const double ri = rsqrt_pos(d);
    const double l = (a == j ? d : A[j - 16]) * ri;
    dg = fma(-l, l, dg);
    if (j < 31) d = __shfl_sync(kFullWarp, dg, j + 1);
    sLc[j * kCholLd + a] = (a >= j) ? l : 0.0;
    if (a == j) sRinv[j] = ri;
    __syncwarp();
    if ((j % kCholPub) == kCholPub - 1 && a == 0) publish_progress(progress, j + 1);
#pragma unroll
    for (int c = j + 1; c < 32; ++c) A[c - 16] = fma(-l, sLc[j * kCholLd + c], A[c - 16]);
  }
  return !bad;
}

// one-warp version (kept for the micro-benchmark, tools/bench_potrf.cu)
__device__ __noinline__ bool warp_potrf_cols(const double* tile, int ld, double* sLc, double* sRinv, volatile int* progress) {
  const int a = threadIdx.x & 31;
  double A[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) A[c] = tile[a + ld * c];
  double dg = tile[a + ld * a];
  bool bad = false;
  double d = __shfl_sync(kFullWarp, dg, 0);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (!(d > 1e-290) || !(d < 1e290)) { bad = true; d = 1.0; }   // not positive definite (or NaN / overflow): flagged, keeps running
    const double ri = rsqrt_pos(d);
    const double l = (a == j ? d : A[j]) * ri;
    dg = fma(-l, l, dg);
    if (j < 31) d = __shfl_sync(kFullWarp, dg, j + 1);
    sLc[j * kCholLd + a] = (a >= j) ? l : 0.0;
    if (a == j) sRinv[j] = ri;
    __syncwarp();
    if ((j % kCholChunk) == kCholChunk - 1 && a == 0) publish_progress(progress, j + 1);
#pragma unroll
    for (int c = j + 1; c < 32; ++c) A[c] = fma(-l, sLc[j * kCholLd + c], A[c]);
  }
  return !bad;
}

template <int WLD = kCholLd>
__device__ __forceinline__ void warp_inverse_cols(const double* sLc, const double* sRinv, volatile int* progress, double* sW) {
  const int a = threadIdx.x & 31;
  double w[32];
#pragma unroll
  for (int r = 0; r < 32; ++r) w[r] = (r == a) ? 1.0 : 0.0;
#pragma unroll
  for (int t = 0; t < 32; ++t) {
    if ((t % kCholChunk) == 0) {
      while (*progress < t + kCholChunk) __nanosleep(40);
      __syncwarp();
    }
    w[t] *= sRinv[t];
#pragma unroll
    for (int r = t + 1; r < 32; ++r) w[r] = fma(-sLc[t * kCholLd + r], w[t], w[r]);
  }
#pragma unroll
  for (int r = 0; r < 32; ++r) sW[r * WLD + a] = w[r];
}

}  // namespace lvi
