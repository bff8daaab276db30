// micro-benchmark: HBM throughput of column-wise (plane-strided) access vs streaming.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)
constexpr int KZ = 41;
// W doubles per lane per plane (chunk = 32*W doubles); NA arrays read, 1 written per plane
template <int W, int NA>
__global__ void colread(const double* __restrict__ a, double* __restrict__ o, long long plane, long long ncol, long long astride) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long c0 = warp * 32 * W + lane * W;
  if (c0 + W > ncol) return;
  double acc[W];
#pragma unroll
  for (int q = 0; q < W; ++q) acc[q] = 0.0;
#pragma unroll 1
  for (int k0 = 0; k0 < KZ; k0 += 8) {
    double v[8][NA][W];
#pragma unroll
    for (int kk = 0; kk < 8; ++kk)
#pragma unroll
      for (int n = 0; n < NA; ++n)
#pragma unroll
        for (int q = 0; q < W; ++q)
          v[kk][n][q] = (k0 + kk < KZ) ? a[n * astride + (k0 + kk) * plane + c0 + q] : 0.0;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
#pragma unroll
      for (int n = 0; n < NA; ++n)
#pragma unroll
        for (int q = 0; q < W; ++q) acc[q] += v[kk][n][q];
      if (k0 + kk < KZ)
#pragma unroll
        for (int q = 0; q < W; ++q) o[(k0 + kk) * plane + c0 + q] = acc[q];
    }
  }
}
__global__ void stream(const double* __restrict__ a, double* __restrict__ o, long long n, int na, long long astride) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double s = 0; for (int q = 0; q < na; ++q) s += a[q * astride + i];
    o[i] = s;
  }
}
int main() {
  const long long plane = 408LL * 406, ncol = plane / 256 * 256, astride = plane * KZ;
  const int NA = 4;
  double *a, *o;
  CK(cudaMalloc(&a, sizeof(double) * astride * NA)); CK(cudaMalloc(&o, sizeof(double) * astride));
  CK(cudaMemset(a, 0, sizeof(double) * astride * NA));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double bytes = (double)ncol * KZ * 8 * (NA + 1);
  auto run = [&](const char* name, auto launch) {
    for (int w = 0; w < 3; ++w) launch();
    cudaEventRecord(e0);
    for (int r = 0; r < 10; ++r) launch();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-28s %8.3f us  %7.1f GB/s  (%s)\n", name, ms * 100, bytes * 10 / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
  };
  run("stream", [&] { stream<<<148 * 16, 256>>>(a, o, ncol * KZ, NA, astride); });
  for (int tb : {32, 64, 128, 256}) {
    char nm[64];
    snprintf(nm, 64, "col W=1 (256B) tb=%d", tb);
    run(nm, [&] { colread<1, NA><<<(unsigned)((ncol / 32 * 32 + tb - 1) / tb), tb>>>(a, o, plane, ncol, astride); });
    snprintf(nm, 64, "col W=2 (512B) tb=%d", tb);
    run(nm, [&] { colread<2, NA><<<(unsigned)((ncol / 2 + tb - 1) / tb), tb>>>(a, o, plane, ncol, astride); });
    snprintf(nm, 64, "col W=4 (1KB) tb=%d", tb);
    run(nm, [&] { colread<4, NA><<<(unsigned)((ncol / 4 + tb - 1) / tb), tb>>>(a, o, plane, ncol, astride); });
  }
  return 0;
}
