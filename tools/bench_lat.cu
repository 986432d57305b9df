// Dependent-chain latencies of the instructions on the pivot chain (diagnostics).
#include <cstdio>
#include <cuda_runtime.h>
constexpr unsigned FULL = 0xffffffffu;
__device__ __forceinline__ double rcp_seed(double d) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d)); return y; }
__device__ __forceinline__ double rsq_seed(double d) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d)); return y; }
template <int MODE>
__global__ void k(double* out, long long* cyc, double x0, double c) {
  __shared__ double s[64];
  double x = x0 + threadIdx.x * 1e-9;
  const int lane = threadIdx.x;
  s[lane] = x;
  __syncwarp();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 64; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (MODE == 0) x = fma(x, c, c);
      if (MODE == 1) x = x * c;
      if (MODE == 2) x = __shfl_sync(FULL, x, (u + 1) & 31);
      if (MODE == 3) x = rcp_seed(x);
      if (MODE == 4) x = rsq_seed(x);
      if (MODE == 5) x = rsqrt(x);
      if (MODE == 6) x = 1.0 / x;
      if (MODE == 7) { s[lane] = x; __syncwarp(); x = s[(lane + 1) & 31]; __syncwarp(); }
      if (MODE == 8) x = sqrt(x);
      if (MODE == 9) { float f = __double2float_rn(x); f = rsqrtf(f); x = f; }
      if (MODE == 10) x = x + c;
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* o; long long* c; cudaMalloc(&o, 1024); cudaMalloc(&c, 8);
  const char* n[] = {"DFMA", "DMUL", "SHFL f64", "rcp.approx.f64", "rsqrt.approx.f64", "rsqrt()", "1.0/x", "STS+syncwarp+LDS+syncwarp", "sqrt()", "d2f+rsqrtf+f2d", "DADD"};
  for (int m = 0; m < 11; ++m) {
    switch (m) {
      case 0: k<0><<<1, 32>>>(o, c, 1.0, 0.999); break;
      case 1: k<1><<<1, 32>>>(o, c, 1.0, 0.9999); break;
      case 2: k<2><<<1, 32>>>(o, c, 1.0, 0.999); break;
      case 3: k<3><<<1, 32>>>(o, c, 1.3, 0.999); break;
      case 4: k<4><<<1, 32>>>(o, c, 1.3, 0.999); break;
      case 5: k<5><<<1, 32>>>(o, c, 1.3, 0.999); break;
      case 6: k<6><<<1, 32>>>(o, c, 1.3, 0.999); break;
      case 7: k<7><<<1, 32>>>(o, c, 1.3, 0.999); break;
      case 8: k<8><<<1, 32>>>(o, c, 1.3, 0.999); break;
      case 9: k<9><<<1, 32>>>(o, c, 1.3, 0.999); break;
      case 10: k<10><<<1, 32>>>(o, c, 1.3, 1e-9); break;
    }
    cudaDeviceSynchronize();
    long long cy; cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
    printf("%-28s %.1f cycles per dependent op\n", n[m], cy / 1024.0);
  }
}
