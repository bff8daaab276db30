// emu_runtime.cpp -- TEST INFRASTRUCTURE (see shim/cuda_runtime.h).
//
// SIMT on the CPU: a launch runs its CTAs one after another on the calling
// (rank) thread; the threads of a CTA are fibers with their own stacks that
// run until they reach a barrier, a warp collective or the end of the kernel.
//   * __syncthreads: generation barrier over the CTA's live fibers
//   * __shfl_*_sync / __any_sync / __ballot_sync: generation barrier over the
//     live lanes of a warp + double-buffered exchange slots
//   * cp.async: queued per thread, performed at wait_group (the latest moment
//     the hardware allows), so data used before its wait reads NaN poison
//   * device memory / dynamic shared memory: host memory filled with 0xFF
//   * EMU_ORDER=reverse runs CTAs and fibers in the opposite order, which
//     exposes dependences between threads that are not ordered by a barrier
// Ranks of a multi-GPU run are threads of one process: peer mappings are plain
// pointers, system-scope release/acquire flags are C++ atomics, NCCL send/recv
// is a mailbox.
#include <sys/mman.h>
#include <time.h>

#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <deque>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "shim/cuda_runtime.h"
#include "shim/nccl.h"

// ThreadSanitizer build (tests/emu_lib.py sanitize="thread", scripts/emu_tsan.sh): compiled without function
// entry/exit instrumentation, so TSan keeps no shadow call stack that the fiber switches would corrupt and every
// emulated thread of a rank simply is the rank's OS thread.  What TSan then reports are accesses of DIFFERENT
// rank threads to one address that no release/acquire flag orders: the races of the halo protocol.
extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl emu_switch
.hidden emu_switch
.type emu_switch,@function
emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size emu_switch,.-emu_switch
)");

namespace emu {

thread_local uint3 tIdx, bIdx;
thread_local dim3 bDim, gDim;

namespace {
constexpr size_t STACK_BYTES = 256 * 1024;

struct Copy { void* dst; const void* src; int bytes; };
struct Fiber {
  void* sp = nullptr;
  bool done = false;
  uint3 tid{0, 0, 0};
  int warp = 0, lane = 0;
  unsigned long long warp_ops = 0;
  std::vector<Copy> pending;       // cp.async copies not yet performed
  std::vector<size_t> group_end;   // pending.size() at each commit
};
struct Barrier {
  int n = 0, exited = 0, arrived = 0;
  unsigned long long gen = 0;
};
struct Warp {
  Barrier b;
  uint64_t slot[2][32];
  bool live[32];
};
struct Sched {
  void* main_sp = nullptr;
  Fiber* cur = nullptr;
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  Barrier cta;
  void (*fn)(void*) = nullptr;
  void* arg = nullptr;
  std::vector<char*> stacks;
  char* smem = nullptr;
  size_t smem_cap = 0;
  unsigned long long progress = 0;
  ~Sched() {
    for (char* s : stacks) munmap(s, STACK_BYTES);
    free(smem);
  }
};
thread_local Sched sched;

int g_order = -1;   // -1: from EMU_ORDER, 0 forward, 1 reverse (emu_set_order)
bool reverse_order() {
  if (g_order >= 0) return g_order == 1;
  static const bool r = [] { const char* e = getenv("EMU_ORDER"); return e && e[0] == 'r'; }();
  return r;
}

void yield() {
  Sched& s = sched;
  emu_switch(&s.cur->sp, s.main_sp);
}
void arrive(Barrier& b) {
  Sched& s = sched;
  const unsigned long long g = b.gen;
  if (++b.arrived >= b.n - b.exited) {
    b.arrived = 0; ++b.gen; ++s.progress;
  } else {
    while (b.gen == g) yield();
  }
}
void leave(Barrier& b) {
  ++b.exited;
  if (b.arrived > 0 && b.arrived >= b.n - b.exited) { b.arrived = 0; ++b.gen; }
}
void flush_copies(Fiber& f, size_t upto) {
  for (size_t q = 0; q < upto; ++q) memcpy(f.pending[q].dst, f.pending[q].src, (size_t)f.pending[q].bytes);
  f.pending.erase(f.pending.begin(), f.pending.begin() + (long)upto);
  for (size_t& e : f.group_end) e -= upto;
}
void fiber_entry() {
  Sched& s = sched;
  s.fn(s.arg);
  Fiber& f = *s.cur;
  flush_copies(f, f.pending.size());
  f.group_end.clear();
  f.done = true;
  ++s.progress;
  s.warps[(size_t)f.warp].live[f.lane] = false;
  leave(s.warps[(size_t)f.warp].b);
  leave(s.cta);
  emu_switch(&f.sp, s.main_sp);
  fprintf(stderr, "emu: finished fiber resumed\n");
  abort();
}
void prepare(Fiber& f, char* stack) {
  // top of the stack, 16-byte aligned; a fake return address keeps the ABI alignment of fiber_entry
  uintptr_t top = ((uintptr_t)stack + STACK_BYTES) & ~(uintptr_t)15;
  void** sp = (void**)top;
  *--sp = nullptr;                 // return address of fiber_entry (never used)
  *--sp = (void*)&fiber_entry;     // popped by emu_switch's ret
  for (int q = 0; q < 6; ++q) *--sp = nullptr;   // rbp rbx r12 r13 r14 r15
  f.sp = (void*)sp;
}

void run_cta(dim3 block, size_t smem) {
  Sched& s = sched;
  const int n = (int)(block.x * block.y * block.z);
  while ((int)s.stacks.size() < n) {
    void* p = mmap(nullptr, STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) { perror("emu: mmap"); abort(); }
    s.stacks.push_back((char*)p);
  }

  if (smem + 64 > s.smem_cap) {
    free(s.smem);
    s.smem_cap = smem + 64;
    if (posix_memalign((void**)&s.smem, 128, s.smem_cap)) abort();
  }
  memset(s.smem, 0xFF, smem + 64);
  const int nw = (n + 31) / 32;
  s.fibers.assign((size_t)n, Fiber());
  s.warps.assign((size_t)nw, Warp());
  s.cta = Barrier();
  s.cta.n = n;
  for (int t = 0; t < n; ++t) {
    Fiber& f = s.fibers[(size_t)t];
    f.tid.x = (unsigned)t % block.x;
    f.tid.y = ((unsigned)t / block.x) % block.y;
    f.tid.z = (unsigned)t / (block.x * block.y);
    f.warp = t / 32; f.lane = t % 32;
    Warp& w = s.warps[(size_t)f.warp];
    w.b.n++; w.live[f.lane] = true;
    prepare(f, s.stacks[(size_t)t]);
  }
  for (Warp& w : s.warps)
    for (int l = w.b.n; l < 32; ++l) w.live[l] = false;
  const bool rev = reverse_order();
  int alive = n;
  while (alive > 0) {
    const unsigned long long before = s.progress;
    alive = 0;
    for (int q = 0; q < n; ++q) {
      Fiber& f = s.fibers[(size_t)(rev ? n - 1 - q : q)];
      if (f.done) continue;
      s.cur = &f;
      tIdx = f.tid;
      emu_switch(&s.main_sp, f.sp);
      if (!f.done) ++alive;
    }
    if (alive > 0 && s.progress == before) {
      fprintf(stderr, "emu: deadlock in CTA (%u,%u,%u): %d threads wait at a barrier or warp collective that the "
                      "others never reach\n", bIdx.x, bIdx.y, bIdx.z, alive);
      abort();
    }
  }
  s.cur = nullptr;
}
}  // namespace

// EMU_JITTER=<usec> (or emu_set_jitter): every launch first sleeps a pseudo-random time below that bound, so that
// the rank threads of a multi-GPU run drift against each other and the peer-store protocol meets other
// interleavings than the lock-step one (a rank far ahead of / behind its neighbours)
int g_jitter_us = -1;
void jitter() {
  int bound = __atomic_load_n(&g_jitter_us, __ATOMIC_RELAXED);
  if (bound < 0) {
    const char* e = getenv("EMU_JITTER");
    bound = e ? atoi(e) : 0;
    __atomic_store_n(&g_jitter_us, bound, __ATOMIC_RELAXED);
  }
  if (bound <= 0) return;
  static thread_local unsigned long long x = 0x9E3779B97F4A7C15ULL ^ (unsigned long long)(uintptr_t)&x;
  x ^= x << 13; x ^= x >> 7; x ^= x << 17;
  const long us = (long)(x % (unsigned long long)bound);
  // one launch in 16 sleeps (emulated kernels take a fraction of a millisecond: the bound should be several
  // milliseconds for a rank to fall whole kernels behind), the others start at once
  if (((x >> 32) & 15) != 0) return;
  struct timespec ts = {us / 1000000L, (us % 1000000L) * 1000L};
  nanosleep(&ts, nullptr);
}
void set_jitter(int us) { __atomic_store_n(&g_jitter_us, us, __ATOMIC_RELAXED); }

void launch(dim3 grid, dim3 block, size_t smem, void (*fn)(void*), void* arg) {
  Sched& s = sched;
  if (s.cur) { fprintf(stderr, "emu: nested launch\n"); abort(); }
  jitter();
  s.fn = fn; s.arg = arg;
  bDim = block; gDim = grid;
  const long long nb = (long long)grid.x * grid.y * grid.z;
  const bool rev = reverse_order();
  for (long long q = 0; q < nb; ++q) {
    const long long b = rev ? nb - 1 - q : q;
    bIdx.x = (unsigned)(b % grid.x);
    bIdx.y = (unsigned)((b / grid.x) % grid.y);
    bIdx.z = (unsigned)(b / ((long long)grid.x * grid.y));
    run_cta(block, smem);
  }
}

void set_order(int reverse) { g_order = reverse; }
void* dyn_smem() { return sched.smem; }
void sync_threads() { arrive(sched.cta); }
void warp_sync() {
  Sched& s = sched;
  Fiber& f = *s.cur;
  arrive(s.warps[(size_t)f.warp].b);
  ++f.warp_ops;
}
uint64_t* warp_slot(int lane) {
  Sched& s = sched;
  Fiber& f = *s.cur;
  return &s.warps[(size_t)f.warp].slot[f.warp_ops & 1][lane & 31];
}
int lane_id() { return sched.cur->lane; }
long long clock_now() {
  return (long long)std::chrono::duration_cast<std::chrono::nanoseconds>(
             std::chrono::steady_clock::now().time_since_epoch()).count();
}
void cp_async_enqueue(void* dst, const void* src, int bytes) { sched.cur->pending.push_back(Copy{dst, src, bytes}); }
void cp_async_commit() { Fiber& f = *sched.cur; f.group_end.push_back(f.pending.size()); }
void cp_async_wait(int keep) {
  Fiber& f = *sched.cur;
  const long ng = (long)f.group_end.size();
  if (ng <= keep) return;
  const size_t upto = f.group_end[(size_t)(ng - keep - 1)];
  flush_copies(f, upto);
  f.group_end.erase(f.group_end.begin(), f.group_end.begin() + (ng - keep));
}
}  // namespace emu

extern "C" void emu_set_order(int reverse) { emu::set_order(reverse); }
extern "C" void emu_set_jitter(int usec) { emu::set_jitter(usec); }

unsigned emu_ballot(int pred) {
  using namespace emu;
  Sched& s = sched;
  Fiber& f = *s.cur;
  Warp& w = s.warps[(size_t)f.warp];
  uint64_t* mine = warp_slot(f.lane);
  uint64_t* base = warp_slot(0);
  // tag the vote with the number of this collective: a lane that has left the
  // kernel (or never reaches this vote) leaves an older tag behind
  const uint64_t tag = (f.warp_ops + 1) << 1;
  *mine = tag | (pred ? 1u : 0u);
  warp_sync();
  unsigned r = 0;
  for (int l = 0; l < 32; ++l)
    if (base[l] == (tag | 1u)) r |= 1u << l;
  (void)w;
  return r;
}

// ---- runtime API -------------------------------------------------------------------
// Streams: kernels and H2D copies run at once (in program order, which is a legal
// stream order), but a device-to-host cudaMemcpy*Async only LANDS when the host
// synchronises with it -- cudaStreamSynchronize of its stream, cudaEventSynchronize
// of an event recorded behind it, cudaDeviceSynchronize -- like on a real copy
// engine.  The source is snapshotted at enqueue time (stream order: later work
// of the stream cannot change what the copy reads).  Host code that looks at a
// buffer before synchronising sees the old bytes, as it may on the GPU.
struct PendingCopy { std::vector<char> data; char* dst; size_t dpitch, width, height; };
struct emuStream { std::deque<PendingCopy> q; long long enq = 0, done = 0; };
struct emuEvent { long long t; emuStream* s = nullptr; long long seq = 0; };
namespace {
thread_local emuStream g_stream0;
emuStream* sq(cudaStream_t s) { return s ? s : &g_stream0; }
void flush_to(emuStream* s, long long seq) {
  while (!s->q.empty() && s->done < seq) {
    PendingCopy& c = s->q.front();
    for (size_t y = 0; y < c.height; ++y) memcpy(c.dst + y * c.dpitch, c.data.data() + y * c.width, c.width);
    s->q.pop_front();
    ++s->done;
  }
}
void defer_d2h(cudaStream_t st, void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h) {
  PendingCopy c;
  c.data.resize(w * h);
  for (size_t y = 0; y < h; ++y) memcpy(c.data.data() + y * w, (const char*)s + y * sp, w);
  c.dst = (char*)d; c.dpitch = dp; c.width = w; c.height = h;
  emuStream* q = sq(st);
  q->q.push_back(std::move(c));
  ++q->enq;
}
}  // namespace

cudaError_t cudaGetLastError() { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
cudaError_t cudaGetDeviceCount(int* n) {
  const char* e = getenv("EMU_NDEV");
  *n = e ? atoi(e) : 8;
  return cudaSuccess;
}
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaMalloc(void** p, size_t n) {
  if (posix_memalign(p, 256, n ? n : 256)) return cudaErrorMemoryAllocation;
  memset(*p, 0xFF, n);
  return cudaSuccess;
}
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) {
  return posix_memalign(p, 256, n ? n : 256) ? cudaErrorMemoryAllocation : cudaSuccess;
}
cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMemset(void* p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind kind, cudaStream_t st) {
  if (kind == cudaMemcpyDeviceToHost) { defer_d2h(st, d, n, s, n, n, 1); return cudaSuccess; }
  memmove(d, s, n);
  return cudaSuccess;
}
cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind kind,
                              cudaStream_t st) {
  if (w > dp || w > sp) return cudaErrorInvalidValue;
  if (kind == cudaMemcpyDeviceToHost) { defer_d2h(st, d, dp, s, sp, w, h); return cudaSuccess; }
  for (size_t y = 0; y < h; ++y) memcpy((char*)d + y * dp, (const char*)s + y * sp, w);
  return cudaSuccess;
}
cudaError_t cudaMemcpy3D(const cudaMemcpy3DParms* p) {
  const cudaPitchedPtr &sp = p->srcPtr, &dp = p->dstPtr;
  for (size_t z = 0; z < p->extent.depth; ++z)
    for (size_t y = 0; y < p->extent.height; ++y) {
      const char* s = (const char*)sp.ptr + ((p->srcPos.z + z) * sp.ysize + (p->srcPos.y + y)) * sp.pitch + p->srcPos.x;
      char* d = (char*)dp.ptr + ((p->dstPos.z + z) * dp.ysize + (p->dstPos.y + y)) * dp.pitch + p->dstPos.x;
      memcpy(d, s, p->extent.width);
    }
  return cudaSuccess;
}
cudaError_t cudaMemcpy3DAsync(const cudaMemcpy3DParms* p, cudaStream_t st) {
  if (p->kind != cudaMemcpyDeviceToHost) return cudaMemcpy3D(p);
  const cudaPitchedPtr &sp = p->srcPtr, &dp = p->dstPtr;
  for (size_t z = 0; z < p->extent.depth; ++z) {   // one deferred 2-D copy per z-plane
    const char* s = (const char*)sp.ptr + ((p->srcPos.z + z) * sp.ysize + p->srcPos.y) * sp.pitch + p->srcPos.x;
    char* d = (char*)dp.ptr + ((p->dstPos.z + z) * dp.ysize + p->dstPos.y) * dp.pitch + p->dstPos.x;
    defer_d2h(st, d, dp.pitch, s, sp.pitch, p->extent.width, p->extent.height);
  }
  return cudaSuccess;
}
cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = new emuStream(); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { return cudaStreamCreate(s); }
cudaError_t cudaStreamDestroy(cudaStream_t s) { flush_to(s, s->enq); delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t s) { emuStream* q = sq(s); flush_to(q, q->enq); return cudaSuccess; }
// (a rank only owns the streams it created: the context's copy streams are synchronised by name in the library)
cudaError_t cudaDeviceSynchronize() { flush_to(&g_stream0, g_stream0.enq); return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emuEvent{0}; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) {
  e->t = emu::clock_now(); e->s = sq(s); e->seq = e->s->enq;
  return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t e) { if (e->s) flush_to(e->s, e->seq); return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = (float)((double)(b->t - a->t) * 1e-6);
  return cudaSuccess;
}
cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) { *can = 1; return cudaSuccess; }
cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
  memset(h, 0, sizeof(*h));
  memcpy(h->reserved, &p, sizeof(p));
  return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) {
  memcpy(p, h.reserved, sizeof(*p));
  return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }

// ---- NCCL mailbox ---------------------------------------------------------------------
struct emuNcclComm { long long id; int rank, nranks; };
namespace {
std::mutex g_mu;
std::condition_variable g_cv;
std::map<std::tuple<long long, int, int>, std::deque<std::vector<char>>> g_box;   // (comm, src, dst)
long long g_next_id = 1;
struct Op { bool send; void* p; size_t bytes; int peer; emuNcclComm* c; };
thread_local int g_depth = 0;
thread_local std::vector<Op> g_ops;
void do_send(const Op& o) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_box[std::make_tuple(o.c->id, o.c->rank, o.peer)].emplace_back((const char*)o.p, (const char*)o.p + o.bytes);
  g_cv.notify_all();
}
ncclResult_t do_recv(const Op& o) {
  std::unique_lock<std::mutex> lk(g_mu);
  auto key = std::make_tuple(o.c->id, o.peer, o.c->rank);
  if (!g_cv.wait_for(lk, std::chrono::seconds(60), [&] { return !g_box[key].empty(); })) return ncclInternalError;
  std::vector<char>& m = g_box[key].front();
  if (m.size() != o.bytes) return ncclInvalidArgument;
  memcpy(o.p, m.data(), o.bytes);
  g_box[key].pop_front();
  return ncclSuccess;
}
ncclResult_t flush_ops() {
  ncclResult_t r = ncclSuccess;
  for (const Op& o : g_ops) if (o.send) do_send(o);
  for (const Op& o : g_ops) if (!o.send) { ncclResult_t q = do_recv(o); if (q != ncclSuccess) r = q; }
  g_ops.clear();
  return r;
}
}  // namespace
extern "C" {
ncclResult_t ncclGetUniqueId(ncclUniqueId* id) {
  std::lock_guard<std::mutex> lk(g_mu);
  memset(id, 0, sizeof(*id));
  const long long v = g_next_id++;
  memcpy(id->internal, &v, sizeof(v));
  return ncclSuccess;
}
ncclResult_t ncclCommInitRank(ncclComm_t* c, int n, ncclUniqueId id, int rank) {
  long long v;
  memcpy(&v, id.internal, sizeof(v));
  *c = new emuNcclComm{v, rank, n};
  return ncclSuccess;
}
ncclResult_t ncclCommDestroy(ncclComm_t c) { delete c; return ncclSuccess; }
ncclResult_t ncclSend(const void* p, size_t n, ncclDataType_t, int peer, ncclComm_t c, cudaStream_t) {
  g_ops.push_back(Op{true, (void*)p, n * 8, peer, c});
  return g_depth ? ncclSuccess : flush_ops();
}
ncclResult_t ncclRecv(void* p, size_t n, ncclDataType_t, int peer, ncclComm_t c, cudaStream_t) {
  g_ops.push_back(Op{false, p, n * 8, peer, c});
  return g_depth ? ncclSuccess : flush_ops();
}
ncclResult_t ncclGroupStart() { ++g_depth; return ncclSuccess; }
ncclResult_t ncclGroupEnd() { return (--g_depth == 0) ? flush_ops() : ncclSuccess; }
const char* ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "emulated NCCL error (timeout or size mismatch)"; }
}
