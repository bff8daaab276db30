// PCIe micro-benchmark 2 (not product): the hand-off's copy shapes, each direction alone and both at once.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t err__ = (x); if (err__ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(err__)); exit(1); } } while (0)
int main() {
  const int nk = 41, nj = getenv("NJ") ? atoi(getenv("NJ")) : 400, nitot = 400, narr = 24;
  const size_t plane = (size_t)nj * nitot * 8, arrb = plane * nk;
  char *d, *d2, *h, *h2;
  CK(cudaMalloc(&d, arrb * narr)); CK(cudaMalloc(&d2, arrb * narr));
  CK(cudaHostAlloc(&h, arrb * narr, cudaHostAllocDefault)); CK(cudaHostAlloc(&h2, arrb * narr, cudaHostAllocDefault));
  for (size_t i = 0; i < arrb * narr; i += 4096) { h[i] = 1; h2[i] = 2; }
  cudaStream_t s1, s2; CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
  cudaEvent_t e[4]; for (auto& x : e) CK(cudaEventCreate(&x));
  for (int mode = 0; mode < 1; ++mode) {          // 0: 2-D slab copies (i-slabs), 1: linear copies of the same size, 2: one linear copy per array
    for (int nslab : {8}) {
      const int ni = nitot / nslab;
      const size_t run = (size_t)nj * ni * 8;
      for (int dir = 0; dir < 3; ++dir) {         // 0 down, 1 up, 2 both
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e[0], s1)); CK(cudaEventRecord(e[2], s2));
        for (int s = 0; s < nslab; ++s)
          for (int a = 0; a < narr; ++a) {
            char* hp = h + a * arrb + (size_t)s * run; char* hp2 = h2 + a * arrb + (size_t)s * run;
            char* dp = d + a * arrb + (size_t)s * run * nk; char* dp2 = d2 + a * arrb + (size_t)s * run * nk;
            if (mode == 0) {
              if (dir != 1) CK(cudaMemcpy2DAsync(hp, plane, dp, run, run, nk, cudaMemcpyDeviceToHost, s1));
              if (dir != 0) CK(cudaMemcpy2DAsync(dp2, run, hp2, plane, run, nk, cudaMemcpyHostToDevice, s2));
            } else if (mode == 1) {
              if (dir != 1) CK(cudaMemcpyAsync(h + a * arrb + (size_t)s * run * nk, dp, run * nk, cudaMemcpyDeviceToHost, s1));
              if (dir != 0) CK(cudaMemcpyAsync(dp2, h2 + a * arrb + (size_t)s * run * nk, run * nk, cudaMemcpyHostToDevice, s2));
            } else if (s == 0) {
              if (dir != 1) CK(cudaMemcpyAsync(h + a * arrb, d + a * arrb, arrb, cudaMemcpyDeviceToHost, s1));
              if (dir != 0) CK(cudaMemcpyAsync(d2 + a * arrb, h2 + a * arrb, arrb, cudaMemcpyHostToDevice, s2));
            }
          }
        CK(cudaEventRecord(e[1], s1)); CK(cudaEventRecord(e[3], s2));
        CK(cudaDeviceSynchronize());
        float t1, t2; CK(cudaEventElapsedTime(&t1, e[0], e[1])); CK(cudaEventElapsedTime(&t2, e[2], e[3]));
        const double gb = (double)arrb * narr / 1e9;
        printf("mode %d slabs %d dir %d: D2H %.2f ms (%.1f GB/s)  H2D %.2f ms (%.1f GB/s)\n", mode, nslab, dir, t1, dir != 1 ? gb / t1 * 1e3 : 0.0, t2, dir != 0 ? gb / t2 * 1e3 : 0.0);
      }
    }
  }
  return 0;
}
