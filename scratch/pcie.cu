// PCIe micro-benchmark (not product): copy engines vs SM-initiated zero-copy, each direction alone and both at once.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t err__ = (x); if (err__ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(err__)); exit(1); } } while (0)
__global__ void zc_copy(double2* __restrict__ dst, const double2* __restrict__ src, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
static float timeit(cudaStream_t s, cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b)); return ms; }
int main(int argc, char** argv) {
  const size_t bytes = (size_t)1 << 30, n2 = bytes / 16;
  double *d_a, *d_b, *h_a, *h_b;
  CK(cudaMalloc(&d_a, bytes)); CK(cudaMalloc(&d_b, bytes));
  CK(cudaHostAlloc(&h_a, bytes, cudaHostAllocMapped)); CK(cudaHostAlloc(&h_b, bytes, cudaHostAllocMapped));
  CK(cudaMemset(d_a, 1, bytes)); CK(cudaMemset(d_b, 2, bytes));
  for (size_t i = 0; i < bytes / 8; i += 512) { h_a[i] = 1.0; h_b[i] = 2.0; }
  cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
  cudaEvent_t e[4]; for (auto& x : e) CK(cudaEventCreate(&x));
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e[0], s1)); CK(cudaMemcpyAsync(h_a, d_a, bytes, cudaMemcpyDeviceToHost, s1)); CK(cudaEventRecord(e[1], s1));
    float t = timeit(s1, e[0], e[1]); printf("D2H copy engine alone      %.1f GB/s\n", bytes / t / 1e6);
    CK(cudaEventRecord(e[0], s1)); CK(cudaMemcpyAsync(d_b, h_b, bytes, cudaMemcpyHostToDevice, s1)); CK(cudaEventRecord(e[1], s1));
    t = timeit(s1, e[0], e[1]); printf("H2D copy engine alone      %.1f GB/s\n", bytes / t / 1e6);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e[0], s1)); CK(cudaEventRecord(e[2], s2));
    CK(cudaMemcpyAsync(h_a, d_a, bytes, cudaMemcpyDeviceToHost, s1)); CK(cudaMemcpyAsync(d_b, h_b, bytes, cudaMemcpyHostToDevice, s2));
    CK(cudaEventRecord(e[1], s1)); CK(cudaEventRecord(e[3], s2));
    float t1 = timeit(s1, e[0], e[1]), t2 = timeit(s2, e[2], e[3]);
    printf("both copy engines          D2H %.1f  H2D %.1f GB/s\n", bytes / t1 / 1e6, bytes / t2 / 1e6);
    for (int blocks : {64, 148, 296, 592}) {
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e[0], s1)); zc_copy<<<blocks, 256, 0, s1>>>((double2*)h_a, (const double2*)d_a, n2); CK(cudaEventRecord(e[1], s1));
      t = timeit(s1, e[0], e[1]); printf("D2H zero-copy %3d CTAs      %.1f GB/s\n", blocks, bytes / t / 1e6);
      CK(cudaEventRecord(e[0], s1)); zc_copy<<<blocks, 256, 0, s1>>>((double2*)d_b, (const double2*)h_b, n2); CK(cudaEventRecord(e[1], s1));
      t = timeit(s1, e[0], e[1]); printf("H2D zero-copy %3d CTAs      %.1f GB/s\n", blocks, bytes / t / 1e6);
    }
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e[0], s1)); CK(cudaEventRecord(e[2], s2));
    zc_copy<<<148, 256, 0, s1>>>((double2*)h_a, (const double2*)d_a, n2); zc_copy<<<296, 256, 0, s2>>>((double2*)d_b, (const double2*)h_b, n2);
    CK(cudaEventRecord(e[1], s1)); CK(cudaEventRecord(e[3], s2));
    t1 = timeit(s1, e[0], e[1]); t2 = timeit(s2, e[2], e[3]);
    printf("both zero-copy             D2H %.1f  H2D %.1f GB/s\n", bytes / t1 / 1e6, bytes / t2 / 1e6);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e[0], s1)); CK(cudaEventRecord(e[2], s2));
    zc_copy<<<148, 256, 0, s1>>>((double2*)h_a, (const double2*)d_a, n2); CK(cudaMemcpyAsync(d_b, h_b, bytes, cudaMemcpyHostToDevice, s2));
    CK(cudaEventRecord(e[1], s1)); CK(cudaEventRecord(e[3], s2));
    t1 = timeit(s1, e[0], e[1]); t2 = timeit(s2, e[2], e[3]);
    printf("D2H zero-copy + H2D engine D2H %.1f  H2D %.1f GB/s\n", bytes / t1 / 1e6, bytes / t2 / 1e6);
  }
  return 0;
}
