// Single-warp issue throughput of independent DFMA / LDS streams (diagnostics).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE, int NW>
__global__ void k(double* out, long long* cyc, double c) {
  __shared__ double s[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = i * 1e-3;
  __syncthreads();
  double x[16];
#pragma unroll
  for (int u = 0; u < 16; ++u) x[u] = 1.0 + u + threadIdx.x * 1e-9;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 64; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (MODE == 0) x[u] = fma(x[u], c, c);
      if (MODE == 1) x[u] = fma(x[u], c, s[(it * 16 + u) & 1023]);               // broadcast LDS + DFMA
      if (MODE == 2) x[u] += s[(it * 16 + u + threadIdx.x) & 1023];              // LDS + DADD
    }
  }
  long long t1 = clock64();
  double r = 0;
#pragma unroll
  for (int u = 0; u < 16; ++u) r += x[u];
  out[threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* o; long long* c; cudaMalloc(&o, 8192); cudaMalloc(&c, 8);
  const char* n[] = {"16 independent DFMA chains", "DFMA + broadcast LDS", "DADD + LDS"};
  for (int m = 0; m < 3; ++m)
    for (int nw : {1, 4, 8}) {
      if (m == 0 && nw == 1) k<0, 1><<<1, 32>>>(o, c, 0.999);
      if (m == 0 && nw == 4) k<0, 4><<<1, 128>>>(o, c, 0.999);
      if (m == 0 && nw == 8) k<0, 8><<<1, 256>>>(o, c, 0.999);
      if (m == 1 && nw == 1) k<1, 1><<<1, 32>>>(o, c, 0.999);
      if (m == 1 && nw == 4) k<1, 4><<<1, 128>>>(o, c, 0.999);
      if (m == 1 && nw == 8) k<1, 8><<<1, 256>>>(o, c, 0.999);
      if (m == 2 && nw == 1) k<2, 1><<<1, 32>>>(o, c, 0.999);
      if (m == 2 && nw == 4) k<2, 4><<<1, 128>>>(o, c, 0.999);
      if (m == 2 && nw == 8) k<2, 8><<<1, 256>>>(o, c, 0.999);
      cudaDeviceSynchronize();
      long long cy; cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
      printf("%-28s warps %d: %.2f cycles per instruction per warp\n", n[m], nw, cy / 1024.0);
    }
}
