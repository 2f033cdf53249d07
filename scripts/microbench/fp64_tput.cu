// fp64 issue rate: independent DFMA streams, 1..8 warps per SM sub-partition (blockDim = 32*4*w: w warps per SMSP)
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, long long* cyc, double y, double z, int iters) {
  double a[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) a[j] = threadIdx.x + j;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < ILP; ++j) a[j] = fma(a[j], y, z);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < ILP; ++j) s += a[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP> void run(int warps_per_smsp, double* o, long long* c) {
  const int iters = 2000;
  k<ILP><<<1, 128 * warps_per_smsp>>>(o, c, 1.0000001, 0.5, iters);
  k<ILP><<<1, 128 * warps_per_smsp>>>(o, c, 1.0000001, 0.5, iters);
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("ILP=%d warps/SMSP=%d: %.2f cycles per warp-DFMA (per warp), %.2f DFMA/cycle/SMSP\n", ILP, warps_per_smsp,
         (double)h / (iters * ILP), (double)iters * ILP * warps_per_smsp / h);
}
int main() {
  double* o; long long* c; cudaMalloc(&o, 1 << 20); cudaMalloc(&c, 64);
  for (int w : {1, 2, 4, 8}) { run<1>(w, o, c); run<4>(w, o, c); run<8>(w, o, c); }
  // single warp only (1 warp in the whole SM)
  k<8><<<1, 32>>>(o, c, 1.0000001, 0.5, 2000); k<8><<<1, 32>>>(o, c, 1.0000001, 0.5, 2000);
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("single warp ILP=8: %.2f cycles per DFMA\n", (double)h / 16000);
  return 0;
}
