// common.cuh -- internal types of the B200 MOLOCH library (not part of the ABI).
//
// Data layout in HBM.  Every 2-D/3-D array of a rank -- cross, U, V or dot
// staggered alike -- lives in the SAME padded box
//     j in [j0, j0+NJ)   i in [i0, i0+NI)   k in [1, nk]
// with j0 = jde1-HJ, i0 = ide1-HI, j fastest (RegCM's (j,i,k) order,
// Main/mod_atm_interface.F90:579-624).  One linear offset therefore addresses
// the same (j,i,k) in every array, ghost cells always exist (HJ=4, HI=3 >= the
// reference's widest ghost of 2) and the first owned point of a row sits on a
// 32-byte boundary (NJ is a multiple of 4 doubles, the arena is 256-B aligned).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/moloch_b200.h"
#include "geo.h"

#ifdef MB_HOST_EMU   // tests/emu: the host build has no vector types
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
#endif

namespace mb {

enum KernelId {
  KID_RESET = 0, KID_TETAVF, KID_SOUND_PRE, KID_DIVDAMP, KID_WSOLVE, KID_UVUPDATE, KID_SFINISH,
  KID_DESTAG, KID_WAF_Z, KID_WAF_Y, KID_WAF_X, KID_CURV, KID_RESTAG, KID_TVIRT, KID_DIAG, KID_PS,
  KID_STATUS, KID_HALO, KID_HALO_PACK, KID_HALO_UNPACK, KID_INIT, KID_WAF_H, KID_BOX,
  KID_BDYVAL, KID_BDYRELAX, KID_BDYFINISH, KID_MKSLICE, KID_TKE, KID_SPECTRAL, KID_MASSCK, KID_DIAGTEN, KID_COUNT
};

struct ProfEvent { cudaEvent_t a, b; int kid; };

// arena slots: 0..MB_NFIELDS-1 are the ABI fields, then the internal arrays
enum Slot { SL_UD = MB_NFIELDS, SL_VD, SL_ZB, SL_2D, SL_ZR, SL_WZ, SL_P0, SL_PROF, SL_TAB, SL_FLAGS, SL_COUNT };
struct Layout {
  std::vector<size_t> off, size;  // bytes, per slot
  size_t stride2d = 0, stridezr = 0, strideprof = 0, total = 0;
};
static_assert(MB_UTEN == MB_TTEN + 1 && MB_VTEN == MB_TTEN + 2 && MB_QXTEN == MB_TTEN + 3 &&
              MB_CHITEN == MB_TTEN + 4 && MB_S == MB_TTEN + 5 && MB_ZDIV2 == MB_TTEN + 6,
              "reset_tendencies clears tten..zdiv2 as one range");
Layout make_layout(const moloch_b200_config& f);

// one directly addressable neighbour (NVLink peer mapping of its arena)
struct Peer {
  bool mapped = false;
  bool ipc = false;           // opened with cudaIpcOpenMemHandle (else same process)
  bool owner = false;         // this side opened the mapping (a second side towards the same rank shares it)
  char* arena = nullptr;      // peer-visible base of the neighbour's arena
  moloch_b200_config cfg;
  Layout layout;
  int NJ = 0, j0 = 0, i0 = 0;
  long long plane = 0;
};

// control blocks of the fused halo rounds (see "fused halo rounds" below)
struct EdgePush {            // one array, edges nexj columns (left/right) and nexi rows (bottom/top) wide
  int j1, j2, i1, i2;        // owned box of the array's staggering
  int nexj, nexi;
  double* q[4];              // the array inside each neighbour's arena (null: side not exchanged)
  int dj[4], di[4];          // index shift into the neighbour's numbering (periodic wrap)
};
struct PushCtl {
  int mask;                  // remote sides (0: nothing to push)
  int pNJ[4], pj0[4], pi0[4];
  long long pplane[4];
  // producer-side signalling (halo_producer_done): the last CTA of the producer grid publishes the round to the
  // neighbours, so the word travels while the producer drains and the consumer is launched
  int sig;                        // 0: the consumer's first CTA signals (halo_sync)
  unsigned long long seq;         // round number, relative to flags[6]
  unsigned long long* pflag[4];   // neighbour's arrival counter for the side it sees me on
  unsigned long long* flags;      // own flag block ([4]: CTA counter, [6]: round base)
};
struct WaitCtl {
  int mask;                  // remote sides to signal and to wait for
  int nosig;                 // the round's producer signals (PushCtl::sig): halo_sync only waits
  unsigned long long seq;         // round number, relative to the base in flags[6] (see Ctx::seq_base)
  unsigned long long* pflag[4];   // neighbour's arrival counter for the side it sees me on
  unsigned long long* flags;      // own flag block: [0..3] arrival counters, [5] timeout marker, [6] round base
  long long timeout_cycles;
};
struct Ctx;
void halo_free(Ctx& c);

struct Ctx {
  moloch_b200_config cfg;
  Geo g;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  // arena
  char* arena = nullptr;
  size_t arena_bytes = 0;
  // field table: device pointer, number of levels, first level, species count
  struct Field { double* p = nullptr; int nk = 0; int klo = 1; int nspec = 1; bool is2d = false; };
  Field f[MB_NFIELDS];
  // extra device arrays
  double *zdiv2b, *mx2, *rmx, *rmu, *rmv;
  double *zru, *zrd;      // static ratios of the vertical WAF pass
  int waf_impl = 2;       // 1: per-loop kernels, 2: field-batched fused kernels
  int waf_zero_skip = 1;  // fused kernels: a field that is exactly +0 in a CTA's / warp's window is not advected
                          // (bit-identical; MOLOCH_B200_WAF_ZEROSKIP=0 / set_option("waf_zero_skip", 0) computes it)
  int wsolve_impl = 12;   // 12: thread per column, sweep arrays in tensor memory, 8 warps/SM (kz <= 41; else 8);
                          // 8..10: row tiles + cp.async.cg ring, sweeps in shared memory; 5..7: round 1's variants;
                          // 2: CTA-parallel coefficients + one-warp sweeps (MOLOCH_B200_WSOLVE / set_option)
  double *wzall, *p0all;  // per-field scratch of the batched wafone
  double* prof[MB_NPROFILES];
  int prof_n[MB_NPROFILES];
  double** d_ptrtab = nullptr;  // device table of field pointers (batched wafone)
  std::vector<double*> h_ptrtab;
  int nadv_fields = 0;
  // scalars (Main/mod_moloch.F90:306-307, :275)
  double dtstepa, dtsound, rdx, rdzita;
  bool initialised = false;
  // halo transport
  Layout layout;
  unsigned long long* flags = nullptr;   // [0..3] arrival counters per side, [4] CTA counter, [5] timeout flag
  unsigned long long halo_seq = 0;       // rounds opened so far (every rank counts them identically)
  // Round numbers reach the kernels RELATIVE to a base that lives in device memory (flags[6]): a captured CUDA
  // graph of the step carries constant offsets and its last node adds the step's number of rounds to the base,
  // so the same graph can be replayed step after step.  seq_base mirrors flags[6] on the host.
  unsigned long long seq_base = 0;
  long long halo_timeout_cycles = 60000000000LL;   // ~30 s at 1.97 GHz (MOLOCH_B200_HALO_TIMEOUT_MS, set_option)
  bool p2p = false;
  bool fuse_halo = true;                 // sound-loop exchanges fused into producer/consumer kernels (p2p only)
  int psignal = 0;                       // 1: fused rounds are signalled by the producer's last edge CTA (halo_producer_done)
                                         // instead of the consumer's first CTA (MOLOCH_B200_PSIGNAL=1 / set_option(
                                         // "halo_psignal", 1)).  Measured and rejected as the default: cordex25 on 8 GPUs
                                         // 1.995 vs 1.873 ms/step (r2s8c) -- the fence + flag store at the producer's tail
                                         // delays the next launch, whereas the consumer-side signal overlaps with the
                                         // consumer's interior CTAs
  int fuse_status = 1;                   // fuse level 2: status_update's fence and ux, vx round folded into status_update /
                                         // uvxtouvstag (MOLOCH_B200_FUSE_STATUS=0 / set_option("fuse_status", 0))
  int fuse_wz = 1;                       // fuse level 2, decomposition along i only: the vertical WAF kernel stores wz's
                                         // edge rows into the neighbours' ghost rows, the horizontal kernel waits
                                         // (MOLOCH_B200_FUSE_WZ=0 / set_option("fuse_wz", 0): stand-alone round)
  int fuse_level = 1;                    // 1: sub-steps 2.. of the sound loop only (the GPU-measured configuration);
                                         // 2: also the first sub-step and advection's u,v / ux,vx exchanges
                                         // (MOLOCH_B200_FUSE_HALO=0|1|2, set_option("fuse_halo"); bench.py times 2 vs 1)
  Peer peer[4];                          // left, right, bottom, top
  void* nccl_comm = nullptr;
  double *sendbuf = nullptr, *recvbuf = nullptr;
  size_t halo_buf_doubles = 0;
  WaitCtl adv_wait = {};                 // u, v pushed by the last sub-step's uvupdate: destagger waits (dynamical_core)
  bool adv_wait_valid = false;
  double* gather_buf = nullptr;          // row/column reductions: the other members' partial sums
  size_t gather_doubles = 0;
  // CUDA graphs of the launch-bound call sequences (capi.cu: run_graphed)
  struct GraphSlot {
    void* exec = nullptr;                 // cudaGraphExec_t
    unsigned long long rounds = 0;        // halo rounds one replay opens
    long long launches = 0;               // kernel launches one replay stands for
    int calls = 0;                        // eager calls so far (the first one warms up lazy allocations)
  };
  enum { G_STEP = 0, G_DYNCORE, G_STATUS, G_COUNT };
  GraphSlot graph[G_COUNT];
  int use_graph = 1;                      // MOLOCH_B200_GRAPH=0 / set_option("graph", 0): eager launches
  // profiling
  bool profiling = false;
  std::vector<ProfEvent> events;
  double prof_ms[KID_COUNT] = {0};
  long long prof_n_launch[KID_COUNT] = {0};
  long long launches = 0;
  // lateral boundary: ba%ibnd planes, tables, RegCM's xbctime
  int* ibnd[3] = {nullptr, nullptr, nullptr};
  bool ibnd_set[3] = {false, false, false};
  double* tab[MB_NTABLES] = {nullptr};
  int tab_n[MB_NTABLES] = {0};
  double xbctime = 0.0;     // Main/mpplib/mod_runparams.F90:102
  double tspectral = 0.0;   // Main/mod_moloch.F90:452
  double calday = 1.0, dayspy = 365.2422;   // Main/mod_sun.F90:316, Share/mod_constants.F90
  double* spec_work = nullptr;   // mospectral_nudge scratch (sx, sxg, sy, syg, stale tails)
  size_t spec_work_doubles = 0;
  double* mass_work = nullptr;   // massck / ps guard partial sums
  size_t mass_work_doubles = 0;
  // staging for set/get
  double* stage = nullptr;
  size_t stage_doubles = 0;
  bool async_xfer = false;
  // pipelined physics hand-off (moloch_b200_handoff): one copy stream and one
  // staging buffer per direction, one event per slab
  cudaStream_t xs_down = nullptr, xs_up = nullptr;       // copy streams (D2H, H2D)
  cudaStream_t xs_pack = nullptr, xs_unpack = nullptr;   // gather / scatter kernels of the slabs
  cudaEvent_t ev_ready = nullptr;
  bool ev_traced = false;
  std::vector<cudaEvent_t> ev_slab;                      // 4 per slab: gathered, arrived, uploaded, scattered
  double *stage_down = nullptr, *stage_up = nullptr;     // two slab-sized staging blocks per direction
  size_t stage_down_doubles = 0, stage_up_doubles = 0;   // size of ONE block
};

extern thread_local std::string g_err;
int fail(const std::string& msg);
// cudaStreamSynchronize(c.stream) + the peer-store transport's timeout marker: every point where the host
// looks at results of the context's stream goes through this, so stale ghost cells never reach it unreported
int sync_stream(Ctx& c);
int halo_timeout_check(Ctx& c);
#define MB_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess)                                                             \
      return mb::fail(std::string(#call) + ": " + cudaGetErrorString(e__));             \
  } while (0)

const char* kernel_name(int kid);

// NVTX ranges with the reference's own names ("moloch", "reset_tendencies", "dynamical_core", "sound",
// "advection", "boundary", "mkslice", "status_update", "real8_3d_exchange_left_right_bottom_top": the !@acc
// nvtxStartRange calls of Main/mod_moloch.F90:323,356,454,556,775,1048,1090,1408 and
// Main/mpplib/mod_mppparam.F90:3816), so that an nsys
// timeline of the library lines up with one of the reference's OpenACC build.  NVTX3 is header-only; without
// a profiler attached a range costs a few nanoseconds on the host.
#ifdef MB_HOST_EMU
struct NvtxRange { explicit NvtxRange(const char*) {} };
#else
struct NvtxRange { explicit NvtxRange(const char* name); ~NvtxRange(); };
#endif

// RAII-less launch bracket used by every launcher: counts the launch and, when
// profiling, records CUDA events on the launching stream around it.
struct LaunchScope {
  Ctx& c; int kid; cudaEvent_t a = nullptr, b = nullptr;
  LaunchScope(Ctx& c_, int kid_);
  ~LaunchScope();
};

// ---- fused halo rounds (peer-store transport only) ----------------------------
// Inside the sound loop the producer of an exchanged array stores its edge cells
// straight into the neighbours' ghost cells while it computes them (plain peer
// stores, no fences, no counters).  The consumer kernel -- the next kernel in
// the stream, so the producer's grid has completed and its stores have been
// flushed -- starts with halo_sync(): its first CTA tells the neighbours "my
// edges are in your ghost cells", every CTA waits for the neighbours' word.
// No separate exchange launch, no per-CTA synchronisation in the producers.
#ifdef __CUDACC__
// system-scope release store / acquire load of a flag word (arrival counters of
// the peer-store halo transport).  MB_HOST_EMU: the tests' host build of this
// source (tests/emu) -- ranks are threads there and the flags C++ atomics.
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
#ifdef MB_HOST_EMU
  __atomic_store_n(p, v, __ATOMIC_RELEASE);
#else
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#endif
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
#ifdef MB_HOST_EMU
  return __atomic_load_n(p, __ATOMIC_ACQUIRE);
#else
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
#endif
}
__device__ __forceinline__ void edge_push(const PushCtl& pc, const EdgePush& e, int j, int i, int k, double v) {
  if (e.q[0] && j >= e.j1 && j < e.j1 + e.nexj && i >= e.i1 && i <= e.i2)
    e.q[0][(long long)(k - 1) * pc.pplane[0] + (long long)(i + e.di[0] - pc.pi0[0]) * pc.pNJ[0] + (j + e.dj[0] - pc.pj0[0])] = v;
  if (e.q[1] && j <= e.j2 && j > e.j2 - e.nexj && i >= e.i1 && i <= e.i2)
    e.q[1][(long long)(k - 1) * pc.pplane[1] + (long long)(i + e.di[1] - pc.pi0[1]) * pc.pNJ[1] + (j + e.dj[1] - pc.pj0[1])] = v;
  if (e.q[2] && i >= e.i1 && i < e.i1 + e.nexi && j >= e.j1 && j <= e.j2)
    e.q[2][(long long)(k - 1) * pc.pplane[2] + (long long)(i + e.di[2] - pc.pi0[2]) * pc.pNJ[2] + (j + e.dj[2] - pc.pj0[2])] = v;
  if (e.q[3] && i <= e.i2 && i > e.i2 - e.nexi && j >= e.j1 && j <= e.j2)
    e.q[3][(long long)(k - 1) * pc.pplane[3] + (long long)(i + e.di[3] - pc.pi0[3]) * pc.pNJ[3] + (j + e.dj[3] - pc.pj0[3])] = v;
}
// The same for a thread that owns one column (j, i) and stores many levels of it (column solver, vertical WAF
// pass): which sides the column's cells go to, and where, is decided once; a store is then one predicated
// remote store per side.  Bottom/top sides only (rows-only decompositions; with left/right neighbours the
// callers use edge_push).
struct ColPush { double* t2; double* t3; };
__device__ __forceinline__ ColPush col_push_init(const PushCtl& pc, const EdgePush& e, int j, int i, bool valid) {
  ColPush c;
  c.t2 = nullptr; c.t3 = nullptr;
  if (pc.mask && valid) {
    if (e.q[2] && i >= e.i1 && i < e.i1 + e.nexi && j >= e.j1 && j <= e.j2) {
      c.t2 = e.q[2] + (long long)(i + e.di[2] - pc.pi0[2]) * pc.pNJ[2] + (j + e.dj[2] - pc.pj0[2]);
    }
    if (e.q[3] && i <= e.i2 && i > e.i2 - e.nexi && j >= e.j1 && j <= e.j2) {
      c.t3 = e.q[3] + (long long)(i + e.di[3] - pc.pi0[3]) * pc.pNJ[3] + (j + e.dj[3] - pc.pj0[3]);
    }
  }
  return c;
}
__device__ __forceinline__ void col_push(const PushCtl& pc, const ColPush& c, long long k, double v) {   // k: 1-based level
  if (c.t2) c.t2[(k - 1) * pc.pplane[2]] = v;
  if (c.t3) c.t3[(k - 1) * pc.pplane[3]] = v;
}
// First statement of a consumer kernel whose grid tiles the rank's box with
// blockIdx.x along j and blockIdx.y along i.  Only the CTAs on an edge of the
// grid read ghost cells of that side, so only they wait for that neighbour; the
// interior CTAs start at once and overlap the signal's flight.
// `reach` > 1 (with the number of columns / rows the grid covers): the kernel's stencil reads `reach` points
// beyond a cell, so a CTA whose cells come within reach-1 of the last column / row waits as well (the last
// CTA may be narrower than the reach).  `bx`, `by`: CTA extent in j and i.
__device__ __forceinline__ void halo_sync(const WaitCtl& w, int reach = 1, int ncols = 0, int nrows = 0,
                                          int bx = 0, int by = 0) {
  if (w.mask == 0) return;
  int need = 0;
  if (blockIdx.x == 0) need |= 1;
  if (blockIdx.x == gridDim.x - 1) need |= 2;
  if (blockIdx.y == 0) need |= 4;
  if (blockIdx.y == gridDim.y - 1) need |= 8;
  if (reach > 1) {
    if ((int)(blockIdx.x + 1) * bx - 1 + reach >= ncols) need |= 2;
    if ((int)(blockIdx.y + 1) * by - 1 + reach >= nrows) need |= 8;
  }
  need &= w.mask;
  const bool first = (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) && !w.nosig;
  if (need == 0 && !first) return;
  if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) {
    const unsigned long long seq = w.seq + w.flags[6];
    if (first) {
      __threadfence_system();
      for (int sd = 0; sd < 4; ++sd)
        if ((w.mask >> sd) & 1)
          st_release_sys(w.pflag[sd], seq);
    }
    const long long t0 = clock64();
    for (int sd = 0; sd < 4; ++sd) {
      if (!((need >> sd) & 1)) continue;
      for (;;) {
        if (ld_acquire_sys(w.flags + sd) >= seq) break;
        // neighbour never arrived: mark the round (moloch_b200_sync reports it); later waits give up at once
        if (ld_acquire_sys(w.flags + 5) != 0ULL) break;
        if (clock64() - t0 > w.timeout_cycles) { w.flags[5] = seq; break; }
      }
    }
  }
  __syncthreads();
}
// Last statement of a producer kernel of a fused round (every thread of the CTAs concerned gets here: no early
// returns in the kernel).  The CTAs that hold cells within three rows of the rank's first / last row -- the only
// ones that store into a neighbour's ghost rows or read the rank's own -- count themselves; the last of them
// publishes the round number to the neighbours: "my edges are in your ghost cells, and every kernel before this
// one has completed".  Same meaning as the consumer-side signal of halo_sync (producer and consumer are
// adjacent in the stream), but the word is on its way while the grid drains and the consumer is launched,
// instead of after the consumer's first CTA has started.  Interior CTAs do nothing here (a fence and a counter
// increment in every CTA of a streaming kernel cost more than the signal gains: measured, r2n2b).
// Rows-only decompositions (no left/right neighbour): halo_fused_begin sets PushCtl::sig only there.
// A kernel describes its CTAs as runs of `c` consecutive cells of a row-major sequence with rows of `L` cells and
// `R` rows (2-D grids: c rows per CTA row, L = 1); `b`, `nb`: the CTA's position / count along that sequence,
// `mult`: CTAs per position (the other grid dimensions).
// Ordering: a CTA's peer stores -> barrier -> device-scope fence + counter increment (thread 0) -> the last
// increment -> system-scope fence -> release store of the flag (causality order is transitive over the two
// scopes, as it is over the kernel boundary + fence of the consumer-side signal).
__device__ __forceinline__ void halo_producer_done(const PushCtl& pc, unsigned b, unsigned nb, int c, long long L,
                                                   long long R, unsigned mult) {
  if (!pc.sig) return;
  constexpr long long reach = 3;    // two ghost rows + one for the staggered row ranges (a superset costs nothing)
  const long long lo_n = (reach * L + c - 1) / c;                       // b < lo_n: holds cells of the first rows
  long long hi = ((R - reach) * L - (c - 1) + (c - 1)) / c;             // b >= hi: holds cells of the last rows
  if ((R - reach) * L - (c - 1) <= 0) hi = 0;
  if (!((long long)b < lo_n || (long long)b >= hi)) return;             // (CTA-uniform)
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) {
    const long long lo_c = lo_n < (long long)nb ? lo_n : (long long)nb;
    const long long hs = hi > lo_n ? hi : lo_n;
    const long long n_pos = lo_c + ((long long)nb > hs ? (long long)nb - hs : 0);
    const unsigned long long n = (unsigned long long)n_pos * mult;
    __threadfence();
    if (atomicAdd(pc.flags + 4, 1ULL) == n - 1ULL) {
      pc.flags[4] = 0ULL;
      const unsigned long long seq = pc.seq + pc.flags[6];
      __threadfence_system();
      for (int sd = 0; sd < 4; ++sd)
        if ((pc.mask >> sd) & 1)
          st_release_sys(pc.pflag[sd], seq);
    }
  }
}
#endif

// ---- launchers (kernels.cu) ------------------------------------------------
int k_seq_bump(Ctx& c, unsigned long long rounds);
int k_reset_tendencies(Ctx& c);
int k_tetavf_init(Ctx& c);
// kernels_sound.cu
int k_sound_div(Ctx& c, double dts, const WaitCtl* wc = nullptr);
int k_uvupdate2(Ctx& c, double dts, const WaitCtl* wc = nullptr, const PushCtl* pc = nullptr, const EdgePush* eu = nullptr,
                const EdgePush* ev = nullptr);
int k_wsolve(Ctx& c, double dts, bool last, const PushCtl* pc = nullptr, const EdgePush* ep = nullptr);
int k_sfinish(Ctx& c);
int k_destagger(Ctx& c, const WaitCtl* wc = nullptr, const PushCtl* pc = nullptr, const EdgePush* eux = nullptr,
                const EdgePush* evx = nullptr);
int k_waf_z(Ctx& c, int first, int count, double dta);
int k_waf_y(Ctx& c, int first, int count, double dta);
int k_waf_x(Ctx& c, int first, int count, double dta);
int k_curvature(Ctx& c, double dta, const PushCtl* pc = nullptr, const EdgePush* eux = nullptr,
                const EdgePush* evx = nullptr);
int k_restagger(Ctx& c, bool with_w, const WaitCtl* wc = nullptr);
int k_tvirt_temp(Ctx& c);
int k_diagnostics(Ctx& c);
int k_status_update(Ctx& c, double dtinc, const WaitCtl* wf = nullptr, const PushCtl* pc = nullptr,
                    const EdgePush* eux = nullptr, const EdgePush* evx = nullptr);
int k_init_static(Ctx& c);
int k_box_copy(Ctx& c, double* dev, double* stage, int ja, int ia, int ka, int nj, int ni, int nk, bool pack,
               cudaStream_t on = nullptr);   // on: another stream than the context's (hand-off copy streams)
// all arrays of one slab of the physics hand-off in one launch: padded device boxes <-> packed staging block
constexpr int SLAB_MAX_ARRAYS = 60;
// off: doubles into the staging block; plane k of the array sits at off + k*w + lead(k), lead(k) = ((a0 + k*p8) & 127) / 8
// doubles (upward direction: every host read starts on a 128-byte boundary, see moloch_b200_handoff; downward: w = nj*ni,
// a0 = p8 = 0); plan: host-side index of the array
struct SlabArray { double* dev; long long off; int ja, nj, ia, ni, ka, nk, w, a0, p8, plan; };
struct SlabTable { SlabArray a[SLAB_MAX_ARRAYS]; int n; };
int k_slab_copy(Ctx& c, const SlabTable& t, double* stage, bool pack, cudaStream_t on);
// kernels_bdy.cu
int k_bdyval(Ctx& c, double xbctime);
int k_bdy_relax(Ctx& c, double xbctime);
int k_bdy_finish(Ctx& c);
int k_mkslice(Ctx& c);
int k_tke_destagger(Ctx& c);
int k_tke_restagger(Ctx& c);
int k_tke_update(Ctx& c, double dtinc);
int k_ibnd_fill(Ctx& c, int* dst, const int* src, int jlo, int jhi, int ilo, int ihi);
int k_spectral_nudge(Ctx& c, double xbctime);
int k_massck(Ctx& c, int what, double* out7);
int k_diag(Ctx& c, int which, bool diff);
// kernels_waf.cu
int k_waf_ratios(Ctx& c);
int k_waf_z2(Ctx& c, int first, int count, double dta, const PushCtl* pc = nullptr, const EdgePush* ewz = nullptr);
int k_waf_yx(Ctx& c, int first, int count, double dta, const WaitCtl* wc = nullptr);

// ---- halo exchange (halo.cu) ------------------------------------------------
enum HaloStag { HS_CROSS = 0, HS_U, HS_V, HS_DOT, HS_P0 };
struct HaloItem { double* p; int nk; };
int halo_exchange(Ctx& c, const HaloItem* items, int nitems, int stag, int nex, bool lr, bool bt, int ext = 0);
struct HaloSpec { const HaloItem* items; int n; int stag; int nex; bool lr, bt; int ext; };
int halo_exchange_multi(Ctx& c, const HaloSpec* specs, int nspecs);
int halo_fence(Ctx& c);
// fused rounds: allocate the round, describe one array's edges
bool halo_fused_available(const Ctx& c);
int halo_fused_begin(Ctx& c, PushCtl* pc, WaitCtl* wc);
int halo_fused_edge(Ctx& c, double* array, int stag, bool lr, bool bt, EdgePush* ep, int nexj = 1, int nexi = -1);
// geometry of the neighbours' boxes only (no round number): for a producer that pushes ahead of the round
// in which the last producer of the same array signals
int halo_push_ctl(Ctx& c, PushCtl* pc);
// row_reduce / column_reduce (Main/mpplib/mod_mppparam.F90:20618-20664): in-place sum of `count` doubles
// over the ranks `members` (ascending, this rank included), added in rank order on every member
int halo_group_sum(Ctx& c, double* data, size_t count, const int* members, int nmem);
int halo_comm_init(Ctx& c, const void* id128);
int halo_comm_id(void* id128);
int halo_p2p_export(Ctx& c, void* blob);
int halo_p2p_connect(Ctx& c, const void* blobs, int nranks);
size_t halo_p2p_blob_size();
void halo_boxes(const moloch_b200_config& cfg, int stag, int nex, bool lr, bool bt,
                int32_t send_box[4][4], int32_t recv_box[4][4]);

}  // namespace mb
