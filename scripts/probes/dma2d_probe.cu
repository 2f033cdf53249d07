// Probe (not part of the product): how fast does the copy engine move a padded observation tensor device -> pinned host
// when the zero padding is skipped with per-agent cudaMemcpy2DAsync calls, compared with one flat copy of everything?
// Geometry = case33: B envs x 6 agents x 50 slots (fp64), true lengths 50/50/18/14/34/34.
//   nvcc -O2 -o /tmp/dma2d scripts/probes/dma2d_probe.cu && /tmp/dma2d
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main() {
  const int B = 4096, NA = 6, OD = 50;
  const int len[NA] = {50, 50, 18, 14, 34, 34};
  const size_t row = size_t(NA) * OD * 8, total = row * B;
  double *dev, *host, *devc;
  CK(cudaMalloc(&dev, total)); CK(cudaMalloc(&devc, total)); CK(cudaMallocHost(&host, total));
  CK(cudaMemset(dev, 1, total)); CK(cudaMemset(devc, 1, total));
  cudaStream_t s; CK(cudaStreamCreate(&s));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  size_t useful = 0; for (int a = 0; a < NA; ++a) useful += size_t(len[a]) * 8 * B;
  auto report = [&](const char* name, size_t bytes, float ms, int reps) {
    printf("%-46s %7.1f us  %6.1f GB/s (%.2f MB)\n", name, 1e3 * ms / reps, bytes * reps / (ms * 1e6), bytes / 1e6);
  };
  const int reps = 50;
  float ms;
  for (int pass = 0; pass < 2; ++pass) {
    CK(cudaEventRecord(e0, s));
    for (int r = 0; r < reps; ++r) CK(cudaMemcpyAsync(host, dev, total, cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s)); CK(cudaEventElapsedTime(&ms, e0, e1));
    if (pass) report("flat copy, padding included", total, ms, reps);
    CK(cudaEventRecord(e0, s));
    for (int r = 0; r < reps; ++r) CK(cudaMemcpyAsync(host, dev, useful, cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s)); CK(cudaEventElapsedTime(&ms, e0, e1));
    if (pass) report("flat copy of the useful bytes only (bound)", useful, ms, reps);
    // padded device layout -> padded host layout, one 2D copy per agent
    CK(cudaEventRecord(e0, s));
    for (int r = 0; r < reps; ++r)
      for (int a = 0; a < NA; ++a)
        CK(cudaMemcpy2DAsync(host + a * OD, row, dev + a * OD, row, size_t(len[a]) * 8, B, cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s)); CK(cudaEventElapsedTime(&ms, e0, e1));
    if (pass) report("6 x memcpy2D padded dev -> padded host", useful, ms, reps);
    // agent-major compact device layout ([a][B][len_a]) -> padded host layout
    CK(cudaEventRecord(e0, s));
    for (int r = 0; r < reps; ++r) {
      size_t off = 0;
      for (int a = 0; a < NA; ++a) {
        CK(cudaMemcpy2DAsync(host + a * OD, row, devc + off, size_t(len[a]) * 8, size_t(len[a]) * 8, B, cudaMemcpyDeviceToHost, s));
        off += size_t(len[a]) * B;
      }
    }
    CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s)); CK(cudaEventElapsedTime(&ms, e0, e1));
    if (pass) report("6 x memcpy2D compact dev -> padded host", useful, ms, reps);
    // env-major compact device layout ([B][sum len]) -> compact host (what a compact wire format would cost)
    CK(cudaEventRecord(e0, s));
    for (int r = 0; r < reps; ++r) CK(cudaMemcpy2DAsync(host, row, devc, useful / B, useful / B, B, cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s)); CK(cudaEventElapsedTime(&ms, e0, e1));
    if (pass) report("1 x memcpy2D compact rows -> padded host rows", useful, ms, reps);
  }
  return 0;
}
