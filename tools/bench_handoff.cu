// Ping-pong of one 32x32 fp64 tile between two CTAs on different SMs through L2 (diagnostics): how long does a flagged hand-off take?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ll_store(unsigned long long* slot, double v) {
  const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v));
  const unsigned long long w0 = (b << 32) | 1ull, w1 = (b & 0xffffffff00000000ull) | 1ull;
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(w0), "l"(w1) : "memory");
}
template <int N>
__device__ __forceinline__ void ll_load_n(const unsigned long long* slot, int stride_words, double (&v)[N]) {
  unsigned long long w0[N], w1[N];
  bool ok;
  do {
#pragma unroll
    for (int i = 0; i < N; ++i)
      asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0[i]), "=l"(w1[i]) : "l"(slot + static_cast<size_t>(i) * stride_words) : "memory");
    ok = true;
#pragma unroll
    for (int i = 0; i < N; ++i) ok = ok && ((w0[i] & w1[i] & 1ull) != 0);
  } while (!ok);
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = __longlong_as_double(static_cast<long long>((w0[i] >> 32) | (w1[i] & 0xffffffff00000000ull)));
}
__device__ __forceinline__ int ld_acquire(const int* p) { int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// MODE 0: LL (16 B per value, every thread polls its 4 words); MODE 1: plain stores + __syncthreads + fence + flag, consumer: thread 0 polls
// flag, __syncthreads, __ldcg loads; MODE 2: LL where only warp 0 polls the LAST-written words (first 32 threads' words) ... not safe in general.
template <int MODE>
__global__ void pingpong(unsigned long long* ll, double* plain, int* flags, int rounds, long long* cyc, int partner_block) {
  if (blockIdx.x != 0 && blockIdx.x != partner_block) return;
  const int me = blockIdx.x == 0 ? 0 : 1;
  const int tid = threadIdx.x;
  __shared__ double sT[1024];
  for (int e = tid; e < 1024; e += 256) sT[e] = e;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    for (int half = 0; half < 2; ++half) {
      const int buf = 2 * r + half;
      if (half == me) {  // send
        if (MODE == 0) {
          unsigned long long* dst = ll + static_cast<size_t>(buf) * 2048;
          for (int e = tid; e < 1024; e += 256) ll_store(dst + 2 * e, sT[e]);
        } else {
          double* dst = plain + static_cast<size_t>(buf) * 1024;
          for (int e = tid; e < 1024; e += 256) dst[e] = sT[e];
          __syncthreads();
          if (tid == 0) { __threadfence(); st_release(flags + buf, 1); }
        }
      } else {  // receive
        if (MODE == 0) {
          const unsigned long long* src = ll + static_cast<size_t>(buf) * 2048;
          double v[4];
          ll_load_n<4>(src + 2 * tid, 512, v);
#pragma unroll
          for (int q = 0; q < 4; ++q) sT[tid + 256 * q] = v[q] + 1.0;
        } else {
          if (tid == 0) while (ld_acquire(flags + buf) == 0) {}
          __syncthreads();
          const double* src = plain + static_cast<size_t>(buf) * 1024;
          for (int e = tid; e < 1024; e += 256) sT[e] = __ldcg(src + e) + 1.0;
        }
        __syncthreads();
      }
    }
  }
  long long t1 = clock64();
  if (tid == 0 && me == 0) cyc[0] = t1 - t0;
}
int main() {
  const int rounds = 200;
  unsigned long long* ll; double* plain; int* flags; long long* cyc;
  cudaMalloc(&ll, sizeof(unsigned long long) * 2048 * 2 * rounds);
  cudaMalloc(&plain, 8 * 1024 * 2 * rounds);
  cudaMalloc(&flags, 4 * 2 * rounds);
  cudaMalloc(&cyc, 8);
  int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  for (int partner : {1, 2, 37, 74, 147}) {
    for (int mode = 0; mode < 2; ++mode) {
      cudaMemset(ll, 0, sizeof(unsigned long long) * 2048 * 2 * rounds);
      cudaMemset(flags, 0, 4 * 2 * rounds);
      if (mode == 0) pingpong<0><<<148, 256>>>(ll, plain, flags, rounds, cyc, partner);
      else pingpong<1><<<148, 256>>>(ll, plain, flags, rounds, cyc, partner);
      cudaError_t e = cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("partner block %3d  %-28s one-way hand-off %.0f cycles = %.2f us  (%s)\n", partner, mode == 0 ? "LL 16B/value" : "stores+fence+flag / poll+ldcg",
             c / (2.0 * rounds), c / (2.0 * rounds) / (khz / 1e3), cudaGetErrorString(e));
    }
  }
}
