// PCIe micro-benchmark 4 (not product): bidirectional 2-D copies, H2D host source shifted by OFF bytes from a 4 KB boundary.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t err__ = (x); if (err__ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(err__)); exit(1); } } while (0)
int main() {
  const int nk = 41, narr = 24, nslab = 8;
  const size_t run = 163840, pitch = run * nslab;           // 160 KiB runs, aligned pitch
  char *d, *d2, *h, *h2;
  CK(cudaMalloc(&d, pitch * nk * narr)); CK(cudaMalloc(&d2, pitch * nk * narr));
  CK(cudaHostAlloc(&h, pitch * nk * narr + 8192, cudaHostAllocDefault)); CK(cudaHostAlloc(&h2, pitch * nk * narr + 8192, cudaHostAllocDefault));
  for (size_t i = 0; i < pitch * nk * narr; i += 4096) { h[i] = 1; h2[i] = 2; }
  cudaStream_t s1, s2; CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
  cudaEvent_t e[4]; for (auto& x : e) CK(cudaEventCreate(&x));
  for (int off : {0, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 0}) {
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e[0], s1)); CK(cudaEventRecord(e[2], s2));
    for (int s = 0; s < nslab; ++s)
      for (int a = 0; a < narr; ++a) {
        CK(cudaMemcpy2DAsync(h + a * pitch * nk + s * run, pitch, d + a * pitch * nk + s * run * nk, run, run, nk, cudaMemcpyDeviceToHost, s1));
        CK(cudaMemcpy2DAsync(d2 + a * pitch * nk + s * run * nk, run, h2 + off + a * pitch * nk + s * run, pitch, run, nk, cudaMemcpyHostToDevice, s2));
      }
    CK(cudaEventRecord(e[1], s1)); CK(cudaEventRecord(e[3], s2));
    CK(cudaDeviceSynchronize());
    float t1, t2; CK(cudaEventElapsedTime(&t1, e[0], e[1])); CK(cudaEventElapsedTime(&t2, e[2], e[3]));
    printf("H2D source offset %4d B: D2H %.1f GB/s  H2D %.1f GB/s\n", off, run * nk * nslab * narr / 1e6 / t1, run * nk * nslab * narr / 1e6 / t2);
  }
  return 0;
}
