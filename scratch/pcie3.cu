// PCIe micro-benchmark 3 (not product): which direction's misalignment hurts a bidirectional transfer, and
// whether an SM-written (zero-copy) D2H with 128-byte aligned interior stores avoids it.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t err__ = (x); if (err__ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(err__)); exit(1); } } while (0)
// copies nrun runs of `len` doubles between a packed device block and host runs (run r at host + r*pitch doubles).
// The host address decides the lane layout: a warp always touches one 256-byte aligned window of the host run.
__global__ void zc_runs(double* __restrict__ host, double* __restrict__ dev, long long len, long long pitch, int nrun, int to_dev) {
  for (int r = blockIdx.y; r < nrun; r += gridDim.y) {
    double* h = host + (long long)r * pitch; double* d = dev + (long long)r * len;
    const long long lead = ((uintptr_t)h & 255) / 8;       // doubles between the previous 256-byte boundary and h
    const long long total = len + lead;                     // window index w <-> element w - lead
    for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (long long)gridDim.x * blockDim.x) {
      const long long q = w - lead;
      if (q >= 0) { if (to_dev) d[q] = h[q]; else h[q] = d[q]; }
    }
  }
}
int main() {
  const int nk = 41, nitot = 400, narr = 24, nslab = 8;
  cudaStream_t s1, s2; CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
  cudaEvent_t e[4]; for (auto& x : e) CK(cudaEventCreate(&x));
  for (int cfg = 3; cfg < 7; ++cfg) {
    // cfg 0: both aligned; 1: D2H unaligned, H2D aligned; 2: D2H aligned, H2D unaligned; 3: both unaligned; 4: D2H zero-copy unaligned + H2D engine unaligned
    const int njd = (cfg == 1 || cfg >= 3) ? 402 : 400, nju = (cfg == 2 || cfg >= 3) ? 402 : 400;
    const size_t pd = (size_t)njd * (nitot + 2 * (njd != 400)) * 8, pu = (size_t)nju * (nitot + 2 * (nju != 400)) * 8;
    char *d, *d2, *h, *h2;
    CK(cudaMalloc(&d, pd * nk * narr)); CK(cudaMalloc(&d2, pu * nk * narr));
    CK(cudaHostAlloc(&h, pd * nk * narr, cudaHostAllocMapped)); CK(cudaHostAlloc(&h2, pu * nk * narr, cudaHostAllocMapped));
    for (size_t i = 0; i < pd * nk * narr; i += 4096) h[i] = 1;
    for (size_t i = 0; i < pu * nk * narr; i += 4096) h2[i] = 2;
    const int ni = nitot / nslab;
    const size_t rund = (size_t)njd * ni * 8, runu = (size_t)nju * ni * 8;
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e[0], s1)); CK(cudaEventRecord(e[2], s2));
      for (int s = 0; s < nslab; ++s)
        for (int a = 0; a < narr; ++a) {
          char* hp = h + a * pd * nk + (size_t)s * rund; char* dp = d + a * pd * nk + (size_t)s * rund * nk;
          char* hp2 = h2 + a * pu * nk + (size_t)s * runu; char* dp2 = d2 + a * pu * nk + (size_t)s * runu * nk;
          // cfg 3: both engines; 4: D2H zero-copy + H2D engine; 5: D2H engine + H2D zero-copy; 6: H2D zero-copy alone
          if (cfg == 4) zc_runs<<<dim3(16, 41), 256, 0, s1>>>((double*)hp, (double*)dp, rund / 8, pd / 8, nk, 0);
          else if (cfg != 6) CK(cudaMemcpy2DAsync(hp, pd, dp, rund, rund, nk, cudaMemcpyDeviceToHost, s1));
          if (cfg >= 5) zc_runs<<<dim3(16, 41), 256, 0, s2>>>((double*)hp2, (double*)dp2, runu / 8, pu / 8, nk, 1);
          else CK(cudaMemcpy2DAsync(dp2, runu, hp2, pu, runu, nk, cudaMemcpyHostToDevice, s2));
        }
      CK(cudaEventRecord(e[1], s1)); CK(cudaEventRecord(e[3], s2));
      CK(cudaDeviceSynchronize());
      float t1, t2; CK(cudaEventElapsedTime(&t1, e[0], e[1])); CK(cudaEventElapsedTime(&t2, e[2], e[3]));
      printf("cfg %d: D2H %.2f ms (%.1f GB/s)  H2D %.2f ms (%.1f GB/s)\n", cfg, t1, rund * nk * nslab * narr / 1e6 / t1, t2, runu * nk * nslab * narr / 1e6 / t2);
    }
    cudaFree(d); cudaFree(d2); cudaFreeHost(h); cudaFreeHost(h2);
  }
  return 0;
}
