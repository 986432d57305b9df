// Does chip-wide fp64 load slow a latency-bound fp64 warp on an otherwise idle SM?  (diagnostics)
// Block 0 times a dependent DFMA chain and an independent DFMA stream of ONE warp; the other blocks either idle or run dense DFMA loops.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int busy, volatile int* stop) {
  extern __shared__ double pad[];   // 120 KB requested at launch: one block per SM, so block 0 has its SM to itself
  double x[16];
#pragma unroll
  for (int u = 0; u < 16; ++u) x[u] = 1.0 + u + threadIdx.x * 1e-9;
  const double c = 0.999;
  if (blockIdx.x != 0) {
    if (!busy) return;
    while (*stop == 0) {
#pragma unroll 1
      for (int it = 0; it < 64; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) x[u] = fma(x[u], c, c);
      }
    }
    double r = 0;
#pragma unroll
    for (int u = 0; u < 16; ++u) r += x[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    return;
  }
  if (threadIdx.x >= 32) return;
  __nanosleep(20000);   // let the other blocks get going
  long long t0 = clock64();
  double y = x[0];
#pragma unroll 1
  for (int it = 0; it < 64; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) y = fma(y, c, c);
  }
  long long t1 = clock64();
#pragma unroll 1
  for (int it = 0; it < 64; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) x[u] = fma(x[u], c, c);
  }
  long long t2 = clock64();
  double r = y;
#pragma unroll
  for (int u = 0; u < 16; ++u) r += x[u];
  out[threadIdx.x] = r;
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; }
  __threadfence();
  *stop = 1;
}
int main() {
  double* o; long long* c; int* stop;
  cudaMalloc(&o, 8 * 296 * 256); cudaMalloc(&c, 16); cudaMalloc(&stop, 4);
  for (int busy = 0; busy < 2; ++busy) {
    cudaMemset(stop, 0, 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
    k<<<148, 256, 120 * 1024>>>(o, c, busy, stop);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, c, 16, cudaMemcpyDeviceToHost);
    printf("other 147 SMs %s: dependent DFMA %.2f cycles each, independent stream %.2f cycles each (%s)\n", busy ? "running dense fp64" : "idle", h[0] / 1024.0, h[1] / 1024.0,
           cudaGetErrorString(e));
  }
}
