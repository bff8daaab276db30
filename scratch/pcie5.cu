// PCIe micro-benchmark 5 (not product): per-copy overhead of stream-ordered 2-D copies and whether several
// streams per direction hide it.  Copies of SIZE_KB each, both directions, NS streams per direction.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t err__ = (x); if (err__ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(err__)); exit(1); } } while (0)
int main() {
  const size_t total = (size_t)1 << 30;
  char *d, *d2, *h, *h2;
  CK(cudaMalloc(&d, total)); CK(cudaMalloc(&d2, total));
  CK(cudaHostAlloc(&h, total, cudaHostAllocDefault)); CK(cudaHostAlloc(&h2, total, cudaHostAllocDefault));
  for (size_t i = 0; i < total; i += 4096) { h[i] = 1; h2[i] = 2; }
  cudaStream_t sd[8], su[8];
  for (int q = 0; q < 8; ++q) { CK(cudaStreamCreateWithFlags(&sd[q], cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&su[q], cudaStreamNonBlocking)); }
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int kb : {512, 1600, 6500}) for (int ns : {1, 2, 4}) for (int dir : {0, 2}) {
    const size_t sz = (size_t)kb * 1024, rows = 41, w = sz / rows / 128 * 128, n = total / (w * rows);
    CK(cudaDeviceSynchronize());
    auto t0 = clock();
    CK(cudaEventRecord(e0, 0));
    for (size_t c = 0; c < n; ++c) {
      CK(cudaMemcpy2DAsync(h + c * w * rows, w, d + c * w * rows, w, w, rows, cudaMemcpyDeviceToHost, sd[c % ns]));
      if (dir == 2) CK(cudaMemcpy2DAsync(d2 + c * w * rows, w, h2 + c * w * rows, w, w, rows, cudaMemcpyHostToDevice, su[c % ns]));
    }
    CK(cudaDeviceSynchronize());
    double ms = (double)(clock() - t0) / CLOCKS_PER_SEC * 1e3;
    printf("copies of %4d KB, %d stream(s)/direction, %s: %.1f GB/s per direction (%zu copies, %.2f ms)\n", kb, ns, dir ? "both" : "D2H only", n * w * rows / 1e6 / ms, n, ms);
  }
  return 0;
}
