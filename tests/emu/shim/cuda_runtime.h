// cuda_runtime.h (shim) -- TEST INFRASTRUCTURE, never part of the product.
//
// A host stand-in for the CUDA runtime and the SIMT execution model, just large
// enough to compile regcm_b200/csrc/*.cu with g++ (after tests/emu_lib.py has
// rewritten the <<<...>>> launches) and to run the *same kernel source* on the
// CPU: every CUDA thread is a fiber, __syncthreads / warp shuffles / votes are
// real rendez-vous points between the fibers of a CTA, cp.async is a deferred
// copy that only lands at the matching wait_group, "device" memory is host
// memory poisoned with NaNs, peer mappings are plain pointers between the rank
// threads of one process.  It exists so that kernels written while no GPU is
// available can still be checked bit for bit against the oracle
// (tests/test_emu_full.py); nothing under regcm_b200/ builds, loads or needs it.
#pragma once
#define MB_HOST_EMU 1
#ifndef __CUDACC__
#define __CUDACC__ 1
#endif
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <cmath>
#include <type_traits>
// every standard header the sources use comes in BEFORE __noinline__ is defined
// (libstdc++ spells its attributes __attribute__((__noinline__)))
#include <map>
#include <string>
#include <utility>
#include <vector>
#include <cstring>
#include <cstdio>

#define __global__
#define __device__
#define __host__
#define __constant__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

namespace emu {
extern thread_local uint3 tIdx, bIdx;
extern thread_local dim3 bDim, gDim;
void* dyn_smem();
void sync_threads();
void warp_sync();                       // rendez-vous of the live lanes of the calling fiber's warp
uint64_t* warp_slot(int lane);          // exchange slot of `lane` for the current warp collective
int lane_id();
long long clock_now();
// cp.async model: copies are queued per thread and performed by wait_group
void cp_async_enqueue(void* dst, const void* src, int bytes);
void cp_async_commit();
void cp_async_wait(int keep_groups);
struct Kernel { void (*fn)(void*); void* arg; };
void launch(dim3 grid, dim3 block, size_t smem, void (*fn)(void*), void* arg);
template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem, F&& f) {
  using Fn = typename std::remove_reference<F>::type;
  launch(grid, block, smem, [](void* p) { (*static_cast<Fn*>(p))(); }, (void*)&f);
}
}  // namespace emu

#define threadIdx (emu::tIdx)
#define blockIdx (emu::bIdx)
#define blockDim (emu::bDim)
#define gridDim (emu::gDim)

// ---- device intrinsics -------------------------------------------------------
template <class A, class B>
inline typename std::common_type<A, B>::type min(A a, B b) { return (b < a) ? b : a; }
template <class A, class B>
inline typename std::common_type<A, B>::type max(A a, B b) { return (a < b) ? b : a; }

inline void __syncthreads() { emu::sync_threads(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_sync(); }
#ifdef __SANITIZE_THREAD__
// ThreadSanitizer models a seq_cst fence as an acquire+release on one global object: every fence of every rank
// would then order everything before it with everything after any later fence of any other rank, and hide
// exactly the races the TSan run looks for.  The protocol's ordering comes from the release stores / acquire
// loads of the arrival flags (on the host's TSO the fences add nothing): compiler barriers here.
inline void __threadfence_system() { __asm__ __volatile__("" ::: "memory"); }
inline void __threadfence() { __asm__ __volatile__("" ::: "memory"); }
#else
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
#endif
inline long long clock64() { return emu::clock_now(); }
template <class T> inline T __ldg(const T* p) { return *p; }

template <class T>
inline T emu_shfl(T v, int src_lane_or_neg) {   // all lanes publish, then read lane `src` (or keep v when < 0)
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  const int lane = emu::lane_id();
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  *emu::warp_slot(lane) = raw;
  // the slot pointer must be taken before the rendez-vous (it names this collective's buffer)
  uint64_t* src = (src_lane_or_neg >= 0 && src_lane_or_neg < 32) ? emu::warp_slot(src_lane_or_neg) : nullptr;
  emu::warp_sync();
  T out = v;
  if (src) memcpy(&out, src, sizeof(T));
  return out;
}
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
  const int lane = emu::lane_id();
  return emu_shfl(v, lane - (int)d >= 0 ? lane - (int)d : -1);
}
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
  const int lane = emu::lane_id();
  return emu_shfl(v, lane + (int)d < 32 ? lane + (int)d : -1);
}
template <class T> inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emu_shfl(v, src & 31); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return emu_shfl(v, emu::lane_id() ^ m); }
unsigned emu_ballot(int pred);
inline unsigned __ballot_sync(unsigned, int pred) { return emu_ballot(pred); }
inline int __any_sync(unsigned, int pred) { return emu_ballot(pred) != 0; }
inline int __all_sync(unsigned, int pred) { return emu_ballot(!pred) == 0; }

inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline int __double2hiint(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(uint32_t)(u >> 32); }
inline int __double2loint(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(uint32_t)u; }
inline double __hiloint2double(int hi, int lo) {
  const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double d; memcpy(&d, &u, 8); return d;
}
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline long long __double_as_longlong(double d) { long long u; memcpy(&u, &d, 8); return u; }
inline double __longlong_as_double(long long u) { double d; memcpy(&d, &u, 8); return d; }

template <class T> inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline double atomicAdd(double* p, double v) {   // CTAs run one after another inside a rank thread
  double o = *p; *p = o + v; return o;
}
template <class T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T atomicExch(T* p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <class T> inline T atomicCAS(T* p, T c, T v) {
  __atomic_compare_exchange_n(p, &c, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return c;
}

// ---- runtime API ---------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1, cudaErrorPeerAccessAlreadyEnabled = 704 };
typedef struct emuStream* cudaStream_t;
typedef struct emuEvent* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaEventDefault = 0, cudaEventDisableTiming = 2, cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaIpcMemHandle_t { char reserved[64]; };
struct cudaPitchedPtr { void* ptr; size_t pitch, xsize, ysize; };
struct cudaPos { size_t x, y, z; };
struct cudaExtent { size_t width, height, depth; };
struct cudaMemcpy3DParms {
  void* srcArray; cudaPos srcPos; cudaPitchedPtr srcPtr;
  void* dstArray; cudaPos dstPos; cudaPitchedPtr dstPtr;
  cudaExtent extent; cudaMemcpyKind kind;
};
inline cudaPitchedPtr make_cudaPitchedPtr(void* d, size_t p, size_t xsz, size_t ysz) { return cudaPitchedPtr{d, p, xsz, ysz}; }
inline cudaPos make_cudaPos(size_t x, size_t y, size_t z) { return cudaPos{x, y, z}; }
inline cudaExtent make_cudaExtent(size_t w, size_t h, size_t d) { return cudaExtent{w, h, d}; }

cudaError_t cudaGetLastError();
const char* cudaGetErrorString(cudaError_t);
cudaError_t cudaGetDeviceCount(int*);
cudaError_t cudaSetDevice(int);
cudaError_t cudaMalloc(void**, size_t);
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
cudaError_t cudaFree(void*);
cudaError_t cudaHostAlloc(void**, size_t, unsigned);
cudaError_t cudaFreeHost(void*);
enum { cudaHostRegisterDefault = 0 };
inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }
cudaError_t cudaMemset(void*, int, size_t);
cudaError_t cudaMemsetAsync(void*, int, size_t, cudaStream_t = nullptr);
cudaError_t cudaMemcpy(void*, const void*, size_t, cudaMemcpyKind);
cudaError_t cudaMemcpyAsync(void*, const void*, size_t, cudaMemcpyKind, cudaStream_t = nullptr);
cudaError_t cudaMemcpy2DAsync(void*, size_t, const void*, size_t, size_t, size_t, cudaMemcpyKind, cudaStream_t = nullptr);
cudaError_t cudaMemcpy3D(const cudaMemcpy3DParms*);
cudaError_t cudaMemcpy3DAsync(const cudaMemcpy3DParms*, cudaStream_t = nullptr);
cudaError_t cudaStreamCreate(cudaStream_t*);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t*, unsigned);
cudaError_t cudaStreamDestroy(cudaStream_t);
cudaError_t cudaStreamSynchronize(cudaStream_t);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaEventCreate(cudaEvent_t*);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t*, unsigned);
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0);
cudaError_t cudaEventDestroy(cudaEvent_t);
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t);
cudaError_t cudaEventElapsedTime(float*, cudaEvent_t, cudaEvent_t);
cudaError_t cudaDeviceCanAccessPeer(int*, int, int);
cudaError_t cudaDeviceEnablePeerAccess(int, unsigned);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*);
cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned);
cudaError_t cudaIpcCloseMemHandle(void*);
template <class F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
