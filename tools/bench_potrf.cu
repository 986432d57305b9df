// Micro-benchmark of the 32x32 diagonal-block kernels of the band factorisation (diagnostics; not part of the library).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/bench_potrf tools/bench_potrf.cu && /tmp/bench_potrf
// Each variant factors the same SPD tile inside one CTA of 256 threads (like the factor kernel); warp 0 reports clock64 cycles between
// the CTA barrier before and after, and the result is checked against a host Cholesky + inverse.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../lvi_exc_b200/csrc/tile_chol.cuh"

constexpr unsigned FULL = 0xffffffffu;
constexpr int kLP = 33;

__device__ __forceinline__ double fast_rsqrt(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  // two Newton steps: y <- y + y * (1 - d y^2) / 2
  double h = 0.5 * y, t = d * y;
  double e = fma(-t, y, 1.0);
  y = fma(h, e, y);
  h = 0.5 * y; t = d * y;
  e = fma(-t, y, 1.0);
  y = fma(h, e, y);
  return y;
}

// ---- V0: the first pipelined version (pivot look-ahead by two shuffles, progress published every pivot) -------
__device__ __noinline__ bool potrf_v0(const double* tile, int ld, double* sLc, double* sRinv, volatile int* progress) {
  const int a = threadIdx.x & 31;
  double A[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) A[c] = tile[a + ld * c];
  bool bad = false;
  double d = __shfl_sync(FULL, A[0], 0);
  if (!(d > 0.0) || !isfinite(d)) { bad = true; d = 1.0; }
  double ri = rsqrt(d);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const double l = A[j] * ri;
    sLc[j * kLP + a] = (a >= j) ? l : 0.0;
    if (a == j) sRinv[j] = ri;
    if (j < 31) {
      const double lc1 = __shfl_sync(FULL, l, j + 1);
      A[j + 1] = fma(-l, lc1, A[j + 1]);
      d = __shfl_sync(FULL, A[j + 1], j + 1);
      if (!(d > 0.0) || !isfinite(d)) { bad = true; d = 1.0; }
      ri = rsqrt(d);
    }
    __syncwarp();
    if (a == 0) { __threadfence_block(); *progress = j + 1; }
#pragma unroll
    for (int c = j + 2; c < 32; ++c) A[c] = fma(-l, sLc[j * kLP + c], A[c]);
  }
  return !bad;
}
__device__ __noinline__ void inverse_v0(const double* sLc, const double* sRinv, volatile int* progress, double* sW) {
  const int a = threadIdx.x & 31;
  double w[32];
#pragma unroll
  for (int r = 0; r < 32; ++r) w[r] = (r == a) ? 1.0 : 0.0;
#pragma unroll
  for (int t = 0; t < 32; ++t) {
    while (*progress < t + 1) {}
    __syncwarp();
    w[t] *= sRinv[t];
#pragma unroll
    for (int r = t + 1; r < 32; ++r) w[r] = fma(-sLc[t * kLP + r], w[t], w[r]);
  }
#pragma unroll
  for (int r = 0; r < 32; ++r) sW[r * kLP + a] = w[r];
}

// ---- V1: own-diagonal trick (one shuffle per pivot), progress published every CH pivots, inverse polls with back-off ----------------
template <int CH, bool FAST>
__device__ __noinline__ bool potrf_v1(const double* tile, int ld, double* sLc, double* sRinv, volatile int* progress) {
  const int a = threadIdx.x & 31;
  double A[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) A[c] = tile[a + ld * c];
  double dg = tile[a + ld * a];
  bool bad = false;
  double d = __shfl_sync(FULL, dg, 0);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (!(d > 0.0) || !isfinite(d)) { bad = true; d = 1.0; }
    const double ri = FAST ? fast_rsqrt(d) : rsqrt(d);
    const double l = A[j] * ri;
    dg = fma(-l, l, dg);
    if (j < 31) d = __shfl_sync(FULL, dg, j + 1);
    sLc[j * kLP + a] = (a >= j) ? l : 0.0;
    if (a == j) sRinv[j] = ri;
    __syncwarp();
    if (CH > 0 && (j % CH) == CH - 1 && a == 0) { __threadfence_block(); *progress = j + 1; }
#pragma unroll
    for (int c = j + 1; c < 32; ++c) A[c] = fma(-l, sLc[j * kLP + c], A[c]);
  }
  return !bad;
}
template <int CH>
__device__ __noinline__ void inverse_v1(const double* sLc, const double* sRinv, volatile int* progress, double* sW) {
  const int a = threadIdx.x & 31;
  double w[32];
#pragma unroll
  for (int r = 0; r < 32; ++r) w[r] = (r == a) ? 1.0 : 0.0;
#pragma unroll
  for (int t = 0; t < 32; ++t) {
    if (CH > 0 && (t % CH) == 0) {
      while (*progress < t + CH) __nanosleep(40);
      __syncwarp();
    }
    w[t] *= sRinv[t];
#pragma unroll
    for (int r = t + 1; r < 32; ++r) w[r] = fma(-sLc[t * kLP + r], w[t], w[r]);
  }
#pragma unroll
  for (int r = 0; r < 32; ++r) sW[r * kLP + a] = w[r];
}

// ---- V2: like V1 but the rank-1 update takes the column from shuffles of registers?  no: keep the column in shared memory but the
// pivot row element through the own-diagonal register; the variant here drops the per-pivot __syncwarp by double-buffering nothing:
// each column has its own shared-memory slot, so only the store -> load ordering inside the warp matters (__syncwarp is that fence).


// ---- V2: the pivot chain goes through a reciprocal (d' = dg - a^2 / d: rcp seed + one cubic Newton step + fma + shuffle), the
// reciprocal square root that scales the column runs beside it, and the first column of the rank-1 update comes from a shuffle
__device__ __forceinline__ double fast_rcp3(double d) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(-d, y, 1.0);
  return fma(fma(e, e, e), y, y);      // y (1 + e + e^2): cubic, 2^-20 seed -> < 2^-58
}
__device__ __forceinline__ double fast_rsqrt3(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double t = d * y;
  const double e = fma(-t, y, 1.0);                // 1 - d y^2
  const double p = fma(0.375, e, 0.5);
  return fma(y, p * e, y);                          // y (1 + e/2 + 3 e^2 / 8)
}
template <int CH>
__device__ __noinline__ bool potrf_v2(const double* tile, int ld, double* sLc, double* sRinv, volatile int* progress) {
  const int a = threadIdx.x & 31;
  double A[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) A[c] = tile[a + ld * c];
  double dg = tile[a + ld * a];
  bool bad = false;
  double d = __shfl_sync(FULL, dg, 0);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (!(d > 0.0) || !isfinite(d)) { bad = true; d = 1.0; }
    const double rc = fast_rcp3(d);
    const double ri = fast_rsqrt3(d);
    const double sq = A[j] * A[j];
    dg = fma(-sq, rc, dg);
    double dn = 0.0;
    if (j < 31) dn = __shfl_sync(FULL, dg, j + 1);
    const double l = (a == j ? d : A[j]) * ri;
    if (j < 31) {
      const double lc1 = __shfl_sync(FULL, l, j + 1);
      A[j + 1] = fma(-l, lc1, A[j + 1]);
    }
    sLc[j * kLP + a] = (a >= j) ? l : 0.0;
    if (a == j) sRinv[j] = ri;
    __syncwarp();
    if (CH > 0 && (j % CH) == CH - 1 && a == 0) { __threadfence_block(); *progress = j + 1; }
#pragma unroll
    for (int c = j + 2; c < 32; ++c) A[c] = fma(-l, sLc[j * kLP + c], A[c]);
    d = dn;
  }
  return !bad;
}

template <int MODE>
__global__ void __launch_bounds__(256, 2) bench_kernel(const double* tile_g, double* W_out, long long* cycles, int reps) {
  __shared__ __align__(16) double sA[1024];
  __shared__ double sM[32 * kLP], sW[32 * kLP], sR[32];
  __shared__ volatile int s_progress;
  const int tid = threadIdx.x;
  long long best = 1ll << 60, first = 0;
  for (int r = 0; r < reps; ++r) {
    for (int e = tid; e < 1024; e += 256) sA[e] = tile_g[e];
    if (tid == 64) s_progress = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (MODE == 0) {
      if (tid < 32) potrf_v0(sA, 32, sM, sR, &s_progress);
      else if (tid < 64) inverse_v0(sM, sR, &s_progress, sW);
    } else if (MODE == 1) {  // v0 potrf alone
      if (tid < 32) potrf_v0(sA, 32, sM, sR, &s_progress);
    } else if (MODE == 2) {
      if (tid < 32) potrf_v1<8, false>(sA, 32, sM, sR, &s_progress);
      else if (tid < 64) inverse_v1<8>(sM, sR, &s_progress, sW);
    } else if (MODE == 3) {
      if (tid < 32) potrf_v1<0, false>(sA, 32, sM, sR, &s_progress);
    } else if (MODE == 4) {
      if (tid < 32) potrf_v1<8, true>(sA, 32, sM, sR, &s_progress);
      else if (tid < 64) inverse_v1<8>(sM, sR, &s_progress, sW);
    } else if (MODE == 5) {
      if (tid < 32) potrf_v1<0, true>(sA, 32, sM, sR, &s_progress);
    } else if (MODE == 6) {
      if (tid < 32) potrf_v1<4, true>(sA, 32, sM, sR, &s_progress);
      else if (tid < 64) inverse_v1<4>(sM, sR, &s_progress, sW);
    } else if (MODE == 10) {
      if (tid < 32) lvi::warp_potrf_cols(sA, 32, sM, sR, &s_progress);
      else if (tid < 64) lvi::warp_inverse_cols(sM, sR, &s_progress, sW);
    } else if (MODE == 11) {
      if (tid < 32) lvi::warp_potrf_head(sA, 32, sM, sR, &s_progress);
      else if (tid < 64) lvi::warp_potrf_tail(sA, 32, sM, sR, &s_progress);
      else if (tid < 96) lvi::warp_inverse_cols(sM, sR, &s_progress, sW);
    } else if (MODE == 12) {
      if (tid < 32) lvi::warp_potrf_head(sA, 32, sM, sR, &s_progress);
      else if (tid < 64) lvi::warp_potrf_tail(sA, 32, sM, sR, &s_progress);
    } else if (MODE == 8) {
      if (tid < 32) potrf_v2<8>(sA, 32, sM, sR, &s_progress);
      else if (tid < 64) inverse_v1<8>(sM, sR, &s_progress, sW);
    } else if (MODE == 9) {
      if (tid < 32) potrf_v2<0>(sA, 32, sM, sR, &s_progress);
    } else if (MODE == 7) {  // sequential in one warp
      if (tid < 32) { potrf_v1<0, true>(sA, 32, sM, sR, &s_progress); __syncwarp(); inverse_v1<0>(sM, sR, &s_progress, sW); }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (t1 - t0 < best) best = t1 - t0;
    if (r == 0) first = t1 - t0;
  }
  if (tid == 0) { cycles[blockIdx.x] = best; cycles[512 + blockIdx.x] = first; }
  if (blockIdx.x == 0)
    for (int e = tid; e < 1024; e += 256) W_out[e] = sW[(e & 31) * kLP + (e >> 5)];
}

int main() {
  std::vector<double> A(1024), L(1024, 0.0), W(1024, 0.0);
  srand(1);
  std::vector<double> B(1024);
  for (auto& v : B) v = rand() / double(RAND_MAX) - 0.5;
  for (int i = 0; i < 32; ++i)
    for (int j = 0; j < 32; ++j) {
      double s = i == j ? 4.0 : 0.0;
      for (int k = 0; k < 32; ++k) s += B[i + 32 * k] * B[j + 32 * k];
      A[i + 32 * j] = s;
    }
  for (int j = 0; j < 32; ++j) {
    double d = A[j + 32 * j];
    for (int k = 0; k < j; ++k) d -= L[j + 32 * k] * L[j + 32 * k];
    L[j + 32 * j] = std::sqrt(d);
    for (int i = j + 1; i < 32; ++i) {
      double s = A[i + 32 * j];
      for (int k = 0; k < j; ++k) s -= L[i + 32 * k] * L[j + 32 * k];
      L[i + 32 * j] = s / L[j + 32 * j];
    }
  }
  for (int c = 0; c < 32; ++c)
    for (int r = 0; r < 32; ++r) {
      double s = r == c ? 1.0 : 0.0;
      for (int k = 0; k < r; ++k) s -= L[r + 32 * k] * W[k + 32 * c];
      W[r + 32 * c] = s / L[r + 32 * r];
    }
  if (getenv("POTRF_UPPER_ZERO")) for (int i = 0; i < 32; ++i) for (int j = i + 1; j < 32; ++j) A[i + 32 * j] = 0.0;
  if (getenv("POTRF_SCALE")) { const double sc = atof(getenv("POTRF_SCALE")); for (auto& v : A) v *= sc; }
  double *dA, *dW;
  long long* dC;
  cudaMalloc(&dA, 8192); cudaMalloc(&dW, 8192); cudaMalloc(&dC, 8 * 1024);
  cudaMemcpy(dA, A.data(), 8192, cudaMemcpyHostToDevice);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const char* names[] = {"v0 pipelined (first version)", "v0 potrf alone", "v1 ch8 pipelined", "v1 potrf alone", "v1 ch8 fast-rsqrt pipelined",
                         "v1 fast-rsqrt potrf alone", "v1 ch4 fast-rsqrt pipelined", "v1 fast potrf then inverse, one warp", "v2 ch8 pipelined", "v2 potrf alone", "library one-warp potrf + inverse", "library head + tail + inverse (3 warps)", "library head + tail alone"};
  const int dyn = getenv("POTRF_DYN_SMEM") ? atoi(getenv("POTRF_DYN_SMEM")) : 0;   // extra dynamic shared memory per block (changes the L1 / shared carve-out)
  for (int grid : {1, 296}) {
    for (int mode = 0; mode < 13; ++mode) {
      cudaMemset(dW, 0, 8192);
      if (dyn) { cudaFuncSetAttribute(bench_kernel<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn); cudaFuncSetAttribute(bench_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn); }
      switch (mode) {
        case 0: bench_kernel<0><<<grid, 256, dyn>>>(dA, dW, dC, 20); break;
        case 1: bench_kernel<1><<<grid, 256, dyn>>>(dA, dW, dC, 20); break;
        case 2: bench_kernel<2><<<grid, 256, dyn>>>(dA, dW, dC, 20); break;
        case 3: bench_kernel<3><<<grid, 256, dyn>>>(dA, dW, dC, 20); break;
        case 4: bench_kernel<4><<<grid, 256, dyn>>>(dA, dW, dC, 20); break;
        case 5: bench_kernel<5><<<grid, 256, dyn>>>(dA, dW, dC, 20); break;
        case 6: bench_kernel<6><<<grid, 256, dyn>>>(dA, dW, dC, 20); break;
        case 7: bench_kernel<7><<<grid, 256, dyn>>>(dA, dW, dC, 20); break;
        case 8: bench_kernel<8><<<grid, 256, dyn>>>(dA, dW, dC, 20); break;
        case 9: bench_kernel<9><<<grid, 256, dyn>>>(dA, dW, dC, 20); break;
        case 10: bench_kernel<10><<<grid, 256, dyn>>>(dA, dW, dC, 20); break;
        case 11: bench_kernel<11><<<grid, 256, dyn>>>(dA, dW, dC, 20); break;
        case 12: bench_kernel<12><<<grid, 256, dyn>>>(dA, dW, dC, 20); break;
      }
      cudaError_t err = cudaDeviceSynchronize();
      long long cyc = 0, cyc_first = 0;
      cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(&cyc_first, dC + 512, 8, cudaMemcpyDeviceToHost);
      std::vector<double> Wg(1024);
      cudaMemcpy(Wg.data(), dW, 8192, cudaMemcpyDeviceToHost);
      double emax = 0;
      const bool has_w = mode != 1 && mode != 3 && mode != 5 && mode != 9 && mode != 12;
      if (has_w)
        for (int e = 0; e < 1024; ++e) emax = std::fmax(emax, std::fabs(Wg[e] - W[e]));
      printf("grid %3d  %-40s %7lld cycles (%.2f us at %d MHz; first call, cold instruction cache: %.2f us)  max|W - W_host| %.2e  %s\n", grid, names[mode],
             cyc, cyc / (khz / 1e3), khz / 1000, cyc_first / (khz / 1e3), emax, cudaGetErrorString(err));
    }
  }
  return 0;
}
