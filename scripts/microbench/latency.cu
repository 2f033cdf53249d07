// micro-benchmark: dependent-issue latency of the ops on the solver's critical path (1 warp)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double x0) {
  __shared__ double sm[64];
  sm[threadIdx.x] = x0 + threadIdx.x; sm[threadIdx.x + 32] = 1.0;
  __syncwarp();
  double x = x0, y = 1.0000001, z = 0.5, acc = 0.0;
  long long t0, t1;
  const int N = 512;
  t0 = clock64();
  #pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, y, z);
  t1 = clock64(); if (threadIdx.x == 0) cyc[0] = (t1 - t0); acc += x;
  t0 = clock64();
  #pragma unroll 16
  for (int i = 0; i < N; ++i) x = x * y;
  t1 = clock64(); if (threadIdx.x == 0) cyc[1] = (t1 - t0); acc += x;
  t0 = clock64();
  #pragma unroll 16
  for (int i = 0; i < N; ++i) x = x + y;
  t1 = clock64(); if (threadIdx.x == 0) cyc[2] = (t1 - t0); acc += x;
  x = 1.5 + x0;
  t0 = clock64();
  #pragma unroll 16
  for (int i = 0; i < N; ++i) x = 1.0 / x + 0.25;
  t1 = clock64(); if (threadIdx.x == 0) cyc[3] = (t1 - t0); acc += x;
  double s, c; x = 0.3 * x0;
  t0 = clock64();
  #pragma unroll 4
  for (int i = 0; i < N; ++i) { sincos(x, &s, &c); x = s + c * 0.1; }
  t1 = clock64(); if (threadIdx.x == 0) cyc[4] = (t1 - t0); acc += x;
  int idx = threadIdx.x;
  sm[threadIdx.x] = (double)((threadIdx.x * 7) % 32);
  __syncwarp();
  t0 = clock64();
  #pragma unroll 16
  for (int i = 0; i < N; ++i) idx = (int)sm[idx];
  t1 = clock64(); if (threadIdx.x == 0) cyc[5] = (t1 - t0);
  t0 = clock64();
  #pragma unroll 16
  for (int i = 0; i < N; ++i) { sm[threadIdx.x] = x; __syncwarp(); x = sm[(threadIdx.x + 1) & 31] + 1.0; __syncwarp(); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[6] = (t1 - t0);
  t0 = clock64();
  #pragma unroll 16
  for (int i = 0; i < N; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0;
  t1 = clock64(); if (threadIdx.x == 0) cyc[7] = (t1 - t0); acc += x;
  float f = 1.1f;
  t0 = clock64();
  #pragma unroll 16
  for (int i = 0; i < N; ++i) f = fmaf(f, 1.0001f, 0.5f);
  t1 = clock64(); if (threadIdx.x == 0) cyc[8] = (t1 - t0);
  x = 2.0 * x0;
  t0 = clock64();
  #pragma unroll 8
  for (int i = 0; i < N; ++i) x = sqrt(x) + 1.5;
  t1 = clock64(); if (threadIdx.x == 0) cyc[9] = (t1 - t0);
  out[threadIdx.x] = x + idx + f + s + acc;
}
int main() {
  double* o; long long* c; cudaMalloc(&o, 256); cudaMalloc(&c, 128);
  k<<<1, 32>>>(o, c, 1.0); k<<<1, 32>>>(o, c, 1.0);
  long long h[10]; cudaMemcpy(h, c, 80, cudaMemcpyDeviceToHost);
  const char* nm[] = {"DFMA", "DMUL", "DADD", "1/x + DADD", "sincos + 2", "LDS.64 + F2I", "STS+sync+LDS+sync+DADD", "SHFL.64 + DADD", "FFMA", "sqrt + DADD"};
  for (int i = 0; i < 10; ++i) printf("%-26s %.1f cycles\n", nm[i], h[i] / 512.0);
  return 0;
}
