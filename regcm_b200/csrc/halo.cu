// halo.cu -- ghost-cell exchange of the 2-D block decomposition.
//
// Replaces real8_3d_exchange_left_right_bottom_top / _left_right / _bottom_top
// (Main/mpplib/mod_mppparam.F90:3809-3878, 4257-4309, 4661-4712): ghost column
// (j1-iex) receives the left neighbour's (j2-(iex-1)), and so on; no corners;
// a side without neighbour (mpi_proc_null) is skipped; a rank that is its own
// neighbour (single rank in a periodic direction, :1304-1314) copies locally.
// Several arrays are exchanged in ONE message per side (the reference sends
// one message per array).  Remote sides go through NCCL send/recv over NVLink.
#include <dlfcn.h>
#include <nccl.h>
#include <unistd.h>
#include <cstring>
#include "common.cuh"

namespace mb {

// ---- box algebra shared by the kernels and the host-only plan --------------
static void owned_box(const moloch_b200_config& c, int stag, int& j1, int& j2, int& i1, int& i2) {
  const int ici1 = c.ice1 + (c.has_bdy_bottom ? 1 : 0), ici2 = c.ice2 - (c.has_bdy_top ? 1 : 0);
  switch (stag) {
    case HS_CROSS: j1 = c.jce1; j2 = c.jce2; i1 = c.ice1; i2 = c.ice2; break;
    case HS_U: j1 = c.jde1; j2 = c.jde2; i1 = c.ice1; i2 = c.ice2; break;
    case HS_V: j1 = c.jce1; j2 = c.jce2; i1 = c.ide1; i2 = c.ide2; break;
    case HS_P0: j1 = c.jce1; j2 = c.jce2; i1 = ici1; i2 = ici2; break;  // exchange_lr(p0,..,ici1,ici2) :955
    default: j1 = c.jde1; j2 = c.jde2; i1 = c.ide1; i2 = c.ide2; break;
  }
}

void halo_boxes(const moloch_b200_config& c, int stag, int nex, bool lr, bool bt, int32_t send_box[4][4],
                int32_t recv_box[4][4]) {
  int j1, j2, i1, i2;
  owned_box(c, stag, j1, j2, i1, i2);
  const int nbr[4] = {c.nbr_left, c.nbr_right, c.nbr_bottom, c.nbr_top};
  for (int sd = 0; sd < 4; ++sd) {
    const bool on = (nbr[sd] >= 0) && ((sd < 2) ? lr : bt);
    int32_t* s = send_box[sd];
    int32_t* r = recv_box[sd];
    if (!on) { s[0] = r[0] = 1; s[1] = r[1] = 0; s[2] = r[2] = 1; s[3] = r[3] = 0; continue; }
    switch (sd) {
      case 0: s[0] = j1; s[1] = j1 + nex - 1; s[2] = i1; s[3] = i2;
              r[0] = j1 - nex; r[1] = j1 - 1; r[2] = i1; r[3] = i2; break;
      case 1: s[0] = j2 - nex + 1; s[1] = j2; s[2] = i1; s[3] = i2;
              r[0] = j2 + 1; r[1] = j2 + nex; r[2] = i1; r[3] = i2; break;
      case 2: s[0] = j1; s[1] = j2; s[2] = i1; s[3] = i1 + nex - 1;
              r[0] = j1; r[1] = j2; r[2] = i1 - nex; r[3] = i1 - 1; break;
      default: s[0] = j1; s[1] = j2; s[2] = i2 - nex + 1; s[3] = i2;
               r[0] = j1; r[1] = j2; r[2] = i2 + 1; r[3] = i2 + nex; break;
    }
  }
}

// ---- kernels ---------------------------------------------------------------
constexpr int HALO_MAX_ITEMS = 64;
struct HaloParams {
  double* p[HALO_MAX_ITEMS];
  int nitems, nk, nex;
  int j1, j2, i1, i2;
  int elo[4], elen[4];   // first index and length of the edge run of each side
  int mode[4];           // 0 none, 1 local copy, 2 remote
  long long seg[4];      // offset (doubles) of each side's segment
  long long count[4];    // doubles per side
};
enum { HM_LOCAL = 0, HM_PACK = 1, HM_UNPACK = 2 };

// element e of side sd -> (send cell, ghost cell); order (item, k, r, iex)
__device__ __forceinline__ void halo_cell(const HaloParams& h, int sd, long long e, int& item, int& k, int& js,
                                          int& is, int& jg, int& ig) {
  const int len = h.elen[sd];
  const int iex = (int)(e % h.nex) + 1; e /= h.nex;
  const int r = (int)(e % len); e /= len;
  k = (int)(e % h.nk) + 1; item = (int)(e / h.nk);
  switch (sd) {
    case 0: js = h.j1 + iex - 1; jg = h.j1 - iex; is = ig = h.elo[sd] + r; break;
    case 1: js = h.j2 - (iex - 1); jg = h.j2 + iex; is = ig = h.elo[sd] + r; break;
    case 2: is = h.i1 + iex - 1; ig = h.i1 - iex; js = jg = h.elo[sd] + r; break;
    default: is = h.i2 - (iex - 1); ig = h.i2 + iex; js = jg = h.elo[sd] + r; break;
  }
}

template <int MODE>
__global__ void moloch_halo(Geo g, HaloParams h, double* __restrict__ buf) {
  const long long tot = h.count[0] + h.count[1] + h.count[2] + h.count[3];
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < tot;
       t += (long long)gridDim.x * blockDim.x) {
    int sd = 0; long long e = t;
    while (e >= h.count[sd]) { e -= h.count[sd]; ++sd; }
    const int want = (MODE == HM_LOCAL) ? 1 : 2;
    if (h.mode[sd] != want) continue;
    int item, k, js, is, jg, ig;
    if (MODE == HM_LOCAL) {
      // my ghost on side sd comes from my own send cells of the opposite side
      halo_cell(h, sd, e, item, k, js, is, jg, ig);
      int item2, k2, js2, is2, jg2, ig2;
      halo_cell(h, sd ^ 1, e, item2, k2, js2, is2, jg2, ig2);
      h.p[item][gidx(g, jg, ig, k)] = h.p[item][gidx(g, js2, is2, k)];
    } else if (MODE == HM_PACK) {
      halo_cell(h, sd, e, item, k, js, is, jg, ig);
      buf[h.seg[sd] + e] = h.p[item][gidx(g, js, is, k)];
    } else {
      halo_cell(h, sd, e, item, k, js, is, jg, ig);
      h.p[item][gidx(g, jg, ig, k)] = buf[h.seg[sd] + e];
    }
  }
}

// ---- NCCL through dlopen (no link-time dependency for single-GPU runs) -----
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static int nccl_load() {
  if (g_nccl.lib) return 0;
#ifdef MB_HOST_EMU   // tests/emu: the in-process mailbox that stands in for NCCL
  g_nccl.lib = (void*)&g_nccl;
  g_nccl.GetUniqueId = ncclGetUniqueId; g_nccl.CommInitRank = ncclCommInitRank; g_nccl.CommDestroy = ncclCommDestroy;
  g_nccl.Send = ncclSend; g_nccl.Recv = ncclRecv; g_nccl.GroupStart = ncclGroupStart; g_nccl.GroupEnd = ncclGroupEnd;
  g_nccl.GetErrorString = ncclGetErrorString;
  return 0;
#endif
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) { g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.lib) break; }
  if (!g_nccl.lib) return fail(std::string("cannot dlopen libnccl.so.2: ") + dlerror());
#define LOADSYM(field, sym)                                                   \
  *(void**)(&g_nccl.field) = dlsym(g_nccl.lib, sym);                          \
  if (!g_nccl.field) return fail(std::string("NCCL symbol missing: ") + sym);
  LOADSYM(GetUniqueId, "ncclGetUniqueId") LOADSYM(CommInitRank, "ncclCommInitRank")
  LOADSYM(CommDestroy, "ncclCommDestroy") LOADSYM(Send, "ncclSend") LOADSYM(Recv, "ncclRecv")
  LOADSYM(GroupStart, "ncclGroupStart") LOADSYM(GroupEnd, "ncclGroupEnd")
  LOADSYM(GetErrorString, "ncclGetErrorString")
#undef LOADSYM
  return 0;
}
#define MB_NCCL(call)                                                                       \
  do {                                                                                      \
    ncclResult_t r__ = (call);                                                              \
    if (r__ != ncclSuccess) return fail(std::string(#call) + ": " + g_nccl.GetErrorString(r__)); \
  } while (0)

// ---- row_reduce / column_reduce ------------------------------------------------------
// MPI_Allreduce(SUM) over a row or column of ranks (cartesian_row/_column_communicator,
// Main/mpplib/mod_mppparam.F90:1459-1469, 20618-20664) as an all-gather of the partial
// sums over NCCL (grouped send/recv between the members) followed by one kernel that adds
// them in rank order on every member: deterministic and identical on all of them.
__global__ void moloch_ordered_sum(double* __restrict__ data, const double* __restrict__ others, long long count,
                                   int nmem, int mypos) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < count;
       e += (long long)gridDim.x * blockDim.x) {
    double acc = 0.0;
    int slot = 0;
    for (int b = 0; b < nmem; ++b) {
      const double v = (b == mypos) ? data[e] : others[(long long)(slot++) * count + e];
      acc = (b == 0) ? v : acc + v;
    }
    data[e] = acc;
  }
}
int halo_group_sum(Ctx& c, double* data, size_t count, const int* members, int nmem) {
  if (nmem <= 1 || count == 0) return 0;
  if (!c.nccl_comm)
    return fail("row/column reduction on more than one rank needs the NCCL communicator (moloch_b200_comm_init)");
  int mypos = -1;
  for (int b = 0; b < nmem; ++b) if (members[b] == c.cfg.rank) mypos = b;
  if (mypos < 0) return fail("halo_group_sum: this rank is not a member of the group");
  const size_t need = (size_t)(nmem - 1) * count;
  if (need > c.gather_doubles) {
    if (sync_stream(c)) return 1;
    if (c.gather_buf) cudaFree(c.gather_buf);
    c.gather_buf = nullptr;
    MB_CUDA(cudaMalloc(&c.gather_buf, need * sizeof(double)));
    c.gather_doubles = need;
  }
  ncclComm_t comm = (ncclComm_t)c.nccl_comm;
  MB_NCCL(g_nccl.GroupStart());
  int slot = 0;
  for (int b = 0; b < nmem; ++b) {
    if (b == mypos) continue;
    MB_NCCL(g_nccl.Send(data, count, ncclDouble, members[b], comm, c.stream));
    MB_NCCL(g_nccl.Recv(c.gather_buf + (size_t)(slot++) * count, count, ncclDouble, members[b], comm, c.stream));
  }
  MB_NCCL(g_nccl.GroupEnd());
  LaunchScope ls(c, KID_SPECTRAL);
  const long long nb = ((long long)count + 255) / 256;
  moloch_ordered_sum<<<(unsigned)(nb < 148 * 8 ? nb : 148 * 8), 256, 0, c.stream>>>(data, c.gather_buf, (long long)count,
                                                                                     nmem, mypos);
  MB_CUDA(cudaGetLastError());
  return 0;
}

int halo_comm_id(void* id128) {
  if (nccl_load()) return 1;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  MB_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return 0;
}

int halo_comm_init(Ctx& c, const void* id128) {
  if (c.cfg.nranks <= 1) return 0;
  if (nccl_load()) return 1;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm;
  MB_CUDA(cudaSetDevice(c.device));
  MB_NCCL(g_nccl.CommInitRank(&comm, c.cfg.nranks, id, c.cfg.rank));
  c.nccl_comm = comm;
  return 0;
}

static void halo_p2p_close(Ctx& c);
void halo_free(Ctx& c) {
  halo_p2p_close(c);
  if (c.nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)c.nccl_comm);
  c.nccl_comm = nullptr;
  if (c.sendbuf) cudaFree(c.sendbuf);
  if (c.recvbuf) cudaFree(c.recvbuf);
  c.sendbuf = c.recvbuf = nullptr;
  if (c.gather_buf) cudaFree(c.gather_buf);
  c.gather_buf = nullptr; c.gather_doubles = 0;
}

static int launch_halo(Ctx& c, int mode, const HaloParams& h, double* buf, long long tot) {
  const int tb = 256;
  long long nb = (tot + tb - 1) / tb;
  if (nb > 148 * 16) nb = 148 * 16;
  if (nb < 1) nb = 1;
  if (mode == HM_LOCAL) {
    LaunchScope ls(c, KID_HALO);
    moloch_halo<HM_LOCAL><<<(unsigned)nb, tb, 0, c.stream>>>(c.g, h, buf);
  } else if (mode == HM_PACK) {
    LaunchScope ls(c, KID_HALO_PACK);
    moloch_halo<HM_PACK><<<(unsigned)nb, tb, 0, c.stream>>>(c.g, h, buf);
  } else {
    LaunchScope ls(c, KID_HALO_UNPACK);
    moloch_halo<HM_UNPACK><<<(unsigned)nb, tb, 0, c.stream>>>(c.g, h, buf);
  }
  MB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// Direct peer transport (NVLink/NVSwitch): the sender's kernel stores its edge
// cells straight into the neighbour's ghost cells (peer-mapped arena) and then
// publishes an arrival counter; the receiver's stream waits on its counters.
// No send/receive buffers, no unpack pass, no host involvement per exchange.
// ---------------------------------------------------------------------------
struct P2PBlob {
  unsigned long long pid;
  int device, rank;
  void* arena;
  cudaIpcMemHandle_t handle;
  moloch_b200_config cfg;
};
size_t halo_p2p_blob_size() { return sizeof(P2PBlob); }

int halo_p2p_export(Ctx& c, void* blob) {
  P2PBlob b;
  memset(&b, 0, sizeof(b));
  b.pid = (unsigned long long)getpid();
  b.device = c.device; b.rank = c.cfg.rank; b.arena = c.arena; b.cfg = c.cfg;
  MB_CUDA(cudaSetDevice(c.device));
  MB_CUDA(cudaIpcGetMemHandle(&b.handle, c.arena));
  memcpy(blob, &b, sizeof(b));
  return 0;
}

int halo_p2p_connect(Ctx& c, const void* blobs, int nranks) {
  if (nranks != c.cfg.nranks) return fail("p2p_connect: nranks mismatch");
  const P2PBlob* all = (const P2PBlob*)blobs;
  const int nbr[4] = {c.cfg.nbr_left, c.cfg.nbr_right, c.cfg.nbr_bottom, c.cfg.nbr_top};
  MB_CUDA(cudaSetDevice(c.device));
  for (int sd = 0; sd < 4; ++sd) {
    Peer& pr = c.peer[sd];
    pr.mapped = false;
    if (nbr[sd] < 0 || nbr[sd] == c.cfg.rank) continue;
    const P2PBlob& b = all[nbr[sd]];
    if (b.rank != nbr[sd]) return fail("p2p_connect: blob table is not indexed by rank");
    bool shared = false;
    for (int q = 0; q < sd; ++q)
      if (c.peer[q].mapped && nbr[q] == nbr[sd]) { pr = c.peer[q]; pr.owner = false; shared = true; break; }
    if (shared) continue;
    if (b.pid == (unsigned long long)getpid()) {
      if (b.device != c.device) {
        int can = 0;
        MB_CUDA(cudaDeviceCanAccessPeer(&can, c.device, b.device));
        if (!can) return fail("p2p_connect: no peer access between the two GPUs");
        cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
          return fail(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        cudaGetLastError();
      }
      pr.arena = (char*)b.arena; pr.ipc = false;
    } else {
      void* ptr = nullptr;
      MB_CUDA(cudaIpcOpenMemHandle(&ptr, b.handle, cudaIpcMemLazyEnablePeerAccess));
      pr.arena = (char*)ptr; pr.ipc = true;
    }
    pr.cfg = b.cfg;
    pr.layout = make_layout(b.cfg);
    const Geo pg = geo_from_cfg(b.cfg);
    pr.NJ = pg.NJ; pr.j0 = pg.j0; pr.i0 = pg.i0; pr.plane = pg.plane;
    pr.mapped = true; pr.owner = true;
  }
  c.p2p = true;
  return 0;
}

static void halo_p2p_close(Ctx& c) {
  for (int sd = 0; sd < 4; ++sd) {
    Peer& pr = c.peer[sd];
    // two sides may share one neighbour (2 ranks in a periodic direction): only the side that
    // opened the IPC handle closes it
    if (pr.mapped && pr.ipc && pr.owner) cudaIpcCloseMemHandle(pr.arena);
    pr.mapped = false; pr.owner = false;
  }
  c.p2p = false;
}

int halo_exchange(Ctx& c, const HaloItem* items, int nitems, int stag, int nex, bool lr, bool bt, int ext);
constexpr int PUSH_MAX_ITEMS = 48, PUSH_MAX_GROUPS = 6;
struct PushGroup {
  int j1, j2, i1, i2;        // owned box of the group's staggering
  int elo[4], elen[4];       // edge run of each side
  int dj[4], di[4];          // index shift into the neighbour's numbering (periodic wrap)
  int nk, nex, first, n;     // levels, width, first item, item count
  long long count[4];        // elements per side (0 = side not part of this group)
};
struct PushParams {
  double* p[PUSH_MAX_ITEMS];        // local arrays
  double* q[4][PUSH_MAX_ITEMS];     // the same arrays inside the neighbour's arena, per side
  PushGroup grp[PUSH_MAX_GROUPS];
  int ngroups;
  int mode[4];                      // 0 none, 1 local copy, 2 peer store
  int pNJ[4], pj0[4], pi0[4];       // neighbour's padded box
  long long pplane[4];
  unsigned long long* pflag[4];     // neighbour's arrival counter for the side it sees me on
  unsigned long long* flags;        // my own counters [0..3], CTA counter [4], timeout flag [5]
  unsigned long long seq;
  int mask;                         // remote sides taking part in this round
  long long timeout_cycles;
};

__device__ __forceinline__ void push_cell(const PushGroup& h, int sd, long long e, int& item, int& k, int& js,
                                          int& is, int& jg, int& ig) {
  const int len = h.elen[sd];
  const int iex = (int)(e % h.nex) + 1; e /= h.nex;
  const int r = (int)(e % len); e /= len;
  k = (int)(e % h.nk) + 1; item = h.first + (int)(e / h.nk);
  switch (sd) {
    case 0: js = h.j1 + iex - 1; jg = h.j1 - iex; is = ig = h.elo[sd] + r; break;
    case 1: js = h.j2 - (iex - 1); jg = h.j2 + iex; is = ig = h.elo[sd] + r; break;
    case 2: is = h.i1 + iex - 1; ig = h.i1 - iex; js = jg = h.elo[sd] + r; break;
    default: is = h.i2 - (iex - 1); ig = h.i2 + iex; js = jg = h.elo[sd] + r; break;
  }
}

// One launch = one exchange round: every edge cell of every array of the round
// is stored straight into the neighbour's ghost cell (or copied locally for a
// periodic self-neighbour); the last CTA then publishes the round number to the
// neighbours and waits until theirs has arrived here.
__global__ void moloch_halo_push(Geo g, PushParams h) {
  long long tot = 0;
  for (int q = 0; q < h.ngroups; ++q) tot += h.grp[q].count[0] + h.grp[q].count[1] + h.grp[q].count[2] + h.grp[q].count[3];
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < tot;
       t += (long long)gridDim.x * blockDim.x) {
    int gq = 0, sd = 0; long long e = t;
    while (e >= h.grp[gq].count[sd]) { e -= h.grp[gq].count[sd]; if (++sd == 4) { sd = 0; ++gq; } }
    const PushGroup& G = h.grp[gq];
    int item, k, js, is, jg, ig;
    push_cell(G, sd, e, item, k, js, is, jg, ig);
    if (h.mode[sd] == 1) {
      int item2, k2, js2, is2, jg2, ig2;
      push_cell(G, sd ^ 1, e, item2, k2, js2, is2, jg2, ig2);
      h.p[item][gidx(g, jg, ig, k)] = h.p[item][gidx(g, js2, is2, k)];
    } else {
      const double val = h.p[item][gidx(g, js, is, k)];
      const long long tgt = (long long)(k - 1) * h.pplane[sd] +
                            (long long)(is + G.di[sd] - h.pi0[sd]) * h.pNJ[sd] + (js + G.dj[sd] - h.pj0[sd]);
      h.q[sd][item][tgt] = val;
    }
  }
  if (h.mask == 0) return;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long done = atomicAdd(h.flags + 4, 1ULL);
    if (done == (unsigned long long)gridDim.x - 1ULL) {
      h.flags[4] = 0ULL;
      const unsigned long long seq = h.seq + h.flags[6];
      __threadfence_system();
      for (int sd = 0; sd < 4; ++sd)
        if ((h.mask >> sd) & 1)
          st_release_sys(h.pflag[sd], seq);
      const long long t0 = clock64();
      for (int sd = 0; sd < 4; ++sd) {
        if (!((h.mask >> sd) & 1)) continue;
        for (;;) {
          if (ld_acquire_sys(h.flags + sd) >= seq) break;
          if (ld_acquire_sys(h.flags + 5) != 0ULL) break;                       // an earlier round timed out
          if (clock64() - t0 > h.timeout_cycles) { h.flags[5] = seq; break; }  // neighbour never arrived
        }
      }
    }
  }
}

static double* peer_ptr(Ctx& c, const Peer& pr, double* local) {
  const size_t d = (size_t)((char*)local - c.arena);
  const Layout& L = c.layout;
  for (int sl = 0; sl < SL_COUNT; ++sl) {
    if (L.size[sl] == 0 || d < L.off[sl] || d >= L.off[sl] + L.size[sl]) continue;
    const size_t within = d - L.off[sl];
    const size_t pb = (size_t)c.g.plane * sizeof(double);
    if (within % pb != 0) return nullptr;
    return (double*)(pr.arena + pr.layout.off[sl] + (within / pb) * (size_t)pr.plane * sizeof(double));
  }
  return nullptr;
}

// A round of independent exchanges in one launch (P2P transport); with NCCL the
// specs are exchanged one after the other.
int halo_exchange_multi(Ctx& c, const HaloSpec* specs, int nspecs) {
  // one device round stands for several of the reference's exchanges: named after the widest of them
  NvtxRange nvtx("real8_3d_exchange_left_right_bottom_top");
  if (!c.p2p) {
    for (int q = 0; q < nspecs; ++q)
      if (halo_exchange(c, specs[q].items, specs[q].n, specs[q].stag, specs[q].nex, specs[q].lr, specs[q].bt,
                        specs[q].ext)) return 1;
    return 0;
  }
  const moloch_b200_config& cf = c.cfg;
  const int nbr[4] = {cf.nbr_left, cf.nbr_right, cf.nbr_bottom, cf.nbr_top};
  int sp = 0, done_in_spec = 0;
  while (sp < nspecs) {
    // every rank numbers the rounds identically, whether it takes part or not
    PushParams h;
    memset(&h, 0, sizeof(h));
    h.seq = ++c.halo_seq - c.seq_base;
    h.flags = c.flags;
    h.timeout_cycles = c.halo_timeout_cycles;
    for (int sd = 0; sd < 4; ++sd) h.mode[sd] = (nbr[sd] < 0) ? 0 : (nbr[sd] == cf.rank ? 1 : 2);
    int nit = 0;
    long long tot = 0;
    while (sp < nspecs && h.ngroups < PUSH_MAX_GROUPS && nit < PUSH_MAX_ITEMS) {
      const HaloSpec& S = specs[sp];
      const int take = (S.n - done_in_spec < PUSH_MAX_ITEMS - nit) ? S.n - done_in_spec : PUSH_MAX_ITEMS - nit;
      if (take > 0) {
        PushGroup& G = h.grp[h.ngroups++];
        owned_box(cf, S.stag, G.j1, G.j2, G.i1, G.i2);
        G.nk = S.items[done_in_spec].nk; G.nex = S.nex; G.first = nit; G.n = take;
        for (int sd = 0; sd < 4; ++sd) {
          const bool on = (nbr[sd] >= 0) && ((sd < 2) ? S.lr : S.bt);
          if (sd < 2) { G.elo[sd] = G.i1 - S.ext * c.g.gb; G.elen[sd] = (G.i2 + S.ext * c.g.gt) - G.elo[sd] + 1; }
          else { G.elo[sd] = G.j1 - S.ext * c.g.gl; G.elen[sd] = (G.j2 + S.ext * c.g.gr) - G.elo[sd] + 1; }
          G.count[sd] = on ? (long long)take * G.nk * G.elen[sd] * S.nex : 0;
          tot += G.count[sd];
          if (on && h.mode[sd] == 2) {
            const Peer& pr = c.peer[sd];
            if (!pr.mapped) return fail("halo_exchange: neighbour not peer-mapped (p2p_connect incomplete)");
            int pj1, pj2, pi1, pi2;
            owned_box(pr.cfg, S.stag, pj1, pj2, pi1, pi2);
            if (sd == 0) G.dj[sd] = (pj2 + 1) - G.j1;
            else if (sd == 1) G.dj[sd] = (pj1 - 1) - G.j2;
            else if (sd == 2) G.di[sd] = (pi2 + 1) - G.i1;
            else G.di[sd] = (pi1 - 1) - G.i2;
            h.pNJ[sd] = pr.NJ; h.pj0[sd] = pr.j0; h.pi0[sd] = pr.i0; h.pplane[sd] = pr.plane;
            h.pflag[sd] = (unsigned long long*)(pr.arena + pr.layout.off[SL_FLAGS]) + (sd ^ 1);
            h.mask |= 1 << sd;
          }
        }
        for (int q = 0; q < take; ++q) {
          const HaloItem& it = S.items[done_in_spec + q];
          if (it.nk != G.nk) return fail("halo_exchange: mixed level counts in one batch");
          h.p[nit + q] = it.p;
          for (int sd = 0; sd < 4; ++sd)
            if (G.count[sd] > 0 && h.mode[sd] == 2) {
              h.q[sd][nit + q] = peer_ptr(c, c.peer[sd], it.p);
              if (!h.q[sd][nit + q]) return fail("halo_exchange: array is not addressable in the neighbour's arena");
            }
        }
        nit += take;
        done_in_spec += take;
      }
      if (done_in_spec >= S.n) { ++sp; done_in_spec = 0; }
    }
    if (tot == 0) continue;   // nothing to move for this rank in this round
    const int tb = 256;
    long long nb = (tot + tb - 1) / tb;
    if (nb > 148 * 8) nb = 148 * 8;
    LaunchScope ls(c, KID_HALO);
    moloch_halo_push<<<(unsigned)nb, tb, 0, c.stream>>>(c.g, h);
    MB_CUDA(cudaGetLastError());
  }
  return 0;
}

// Synchronisation-only round of the peer-store transport.  A rank may run at most
// one round ahead of a neighbour (every round waits for the neighbour's arrival),
// so a push is safe whenever the kernels between the neighbour's previous two
// rounds do not read the ghost cells being overwritten.  The one place of the
// step where the same arrays are exchanged in two consecutive rounds with a
// reader in between (ux/vx: uvxtouvstag at the end of advection, then again in
// status_update) gets this fence in between.  No-op for NCCL and single ranks.
int halo_fence(Ctx& c) {
  if (!c.p2p) return 0;
  const moloch_b200_config& cf = c.cfg;
  const int nbr[4] = {cf.nbr_left, cf.nbr_right, cf.nbr_bottom, cf.nbr_top};
  PushParams h;
  memset(&h, 0, sizeof(h));
  h.seq = ++c.halo_seq - c.seq_base;
  h.flags = c.flags;
  h.timeout_cycles = c.halo_timeout_cycles;
  for (int sd = 0; sd < 4; ++sd) {
    if (nbr[sd] < 0 || nbr[sd] == cf.rank) continue;
    const Peer& pr = c.peer[sd];
    if (!pr.mapped) return fail("halo_fence: neighbour not peer-mapped (p2p_connect incomplete)");
    h.pflag[sd] = (unsigned long long*)(pr.arena + pr.layout.off[SL_FLAGS]) + (sd ^ 1);
    h.mask |= 1 << sd;
  }
  if (h.mask == 0) return 0;
  LaunchScope ls(c, KID_HALO);
  moloch_halo_push<<<1, 32, 0, c.stream>>>(c.g, h);
  MB_CUDA(cudaGetLastError());
  return 0;
}

// ---- fused rounds: host side ------------------------------------------------
bool halo_fused_available(const Ctx& c) {
  if (!c.p2p || !c.fuse_halo) return false;
  const int nbr[4] = {c.cfg.nbr_left, c.cfg.nbr_right, c.cfg.nbr_bottom, c.cfg.nbr_top};
  for (int sd = 0; sd < 4; ++sd)
    if (nbr[sd] == c.cfg.rank) return false;   // periodic self-neighbour: local copies, not fusable
  return true;
}

// Allocates the round number (every rank does, in the same program order) and
// fills the signal/wait blocks for all remote neighbours.
int halo_fused_begin(Ctx& c, PushCtl* pc, WaitCtl* wc) {
  const moloch_b200_config& cf = c.cfg;
  const int nbr[4] = {cf.nbr_left, cf.nbr_right, cf.nbr_bottom, cf.nbr_top};
  memset(pc, 0, sizeof(*pc));
  memset(wc, 0, sizeof(*wc));
  wc->seq = ++c.halo_seq - c.seq_base; wc->flags = c.flags; wc->timeout_cycles = c.halo_timeout_cycles;
  for (int sd = 0; sd < 4; ++sd) {
    if (nbr[sd] < 0 || nbr[sd] == cf.rank) continue;
    const Peer& pr = c.peer[sd];
    if (!pr.mapped) return fail("halo_fused_begin: neighbour not peer-mapped (p2p_connect incomplete)");
    pc->pNJ[sd] = pr.NJ; pc->pj0[sd] = pr.j0; pc->pi0[sd] = pr.i0; pc->pplane[sd] = pr.plane;
    wc->pflag[sd] = (unsigned long long*)(pr.arena + pr.layout.off[SL_FLAGS]) + (sd ^ 1);
    pc->mask |= 1 << sd;
  }
  wc->mask = pc->mask;
  if (c.psignal && pc->mask && !(pc->mask & 3)) {   // rows-only: the producer's last edge CTA signals, the consumer only waits
    pc->sig = 1; pc->seq = wc->seq; pc->flags = c.flags;
    for (int sd = 0; sd < 4; ++sd) pc->pflag[sd] = wc->pflag[sd];
    wc->nosig = 1;
  }
  return 0;
}

int halo_push_ctl(Ctx& c, PushCtl* pc) {
  WaitCtl unused;
  const unsigned long long seq = c.halo_seq;
  const int rc = halo_fused_begin(c, pc, &unused);
  c.halo_seq = seq;   // no round is opened
  pc->sig = 0;        // ... and none is signalled
  return rc;
}

// Edges of width nex of one array (exchange_lr / _bt / _lrbt semantics, no corners).
// A periodic self-neighbour cannot be fused (the caller checks fusable()).
int halo_fused_edge(Ctx& c, double* array, int stag, bool lr, bool bt, EdgePush* ep, int nexj, int nexi) {
  const moloch_b200_config& cf = c.cfg;
  const int nbr[4] = {cf.nbr_left, cf.nbr_right, cf.nbr_bottom, cf.nbr_top};
  memset(ep, 0, sizeof(*ep));
  ep->nexj = nexj; ep->nexi = nexi < 0 ? nexj : nexi;
  owned_box(cf, stag, ep->j1, ep->j2, ep->i1, ep->i2);
  for (int sd = 0; sd < 4; ++sd) {
    const bool on = (nbr[sd] >= 0) && ((sd < 2) ? lr : bt);
    if (!on) continue;
    if (nbr[sd] == cf.rank) return fail("halo_fused_edge: self-neighbour rounds are not fusable");
    const Peer& pr = c.peer[sd];
    int pj1, pj2, pi1, pi2;
    owned_box(pr.cfg, stag, pj1, pj2, pi1, pi2);
    if (sd == 0) ep->dj[sd] = (pj2 + 1) - ep->j1;
    else if (sd == 1) ep->dj[sd] = (pj1 - 1) - ep->j2;
    else if (sd == 2) ep->di[sd] = (pi2 + 1) - ep->i1;
    else ep->di[sd] = (pi1 - 1) - ep->i2;
    ep->q[sd] = peer_ptr(c, pr, array);
    if (!ep->q[sd]) return fail("halo_fused_edge: array is not addressable in the neighbour's arena");
  }
  return 0;
}

// `ext` widens the edge run of every side by `ext` ghost points at each end
// that has a neighbour: an lr exchange followed by a bt exchange with ext > 0
// (or the other way round) also fills the corner ghosts.
int halo_exchange(Ctx& c, const HaloItem* items, int nitems, int stag, int nex, bool lr, bool bt, int ext) {
  const moloch_b200_config& cf = c.cfg;
  const int nbr[4] = {cf.nbr_left, cf.nbr_right, cf.nbr_bottom, cf.nbr_top};
  bool any = false;
  for (int sd = 0; sd < 4; ++sd) if (nbr[sd] >= 0 && ((sd < 2) ? lr : bt)) any = true;
  if (c.p2p) {
    HaloSpec sp = {items, nitems, stag, nex, lr, bt, ext};
    return halo_exchange_multi(c, &sp, 1);
  }
  if (!any || nitems == 0) return 0;
  for (int first = 0; first < nitems; first += HALO_MAX_ITEMS) {
    const int n = (nitems - first < HALO_MAX_ITEMS) ? nitems - first : HALO_MAX_ITEMS;
    HaloParams h;
    h.nitems = n; h.nk = items[first].nk; h.nex = nex;
    for (int q = 0; q < n; ++q) {
      if (items[first + q].nk != h.nk) return fail("halo_exchange: mixed level counts in one batch");
      h.p[q] = items[first + q].p;
    }
    owned_box(cf, stag, h.j1, h.j2, h.i1, h.i2);
    long long off = 0, tot = 0;
    bool has_local = false, has_remote = false;
    for (int sd = 0; sd < 4; ++sd) {
      const bool on = (nbr[sd] >= 0) && ((sd < 2) ? lr : bt);
      if (sd < 2) {
        h.elo[sd] = h.i1 - ext * c.g.gb;
        h.elen[sd] = (h.i2 + ext * c.g.gt) - h.elo[sd] + 1;
      } else {
        h.elo[sd] = h.j1 - ext * c.g.gl;
        h.elen[sd] = (h.j2 + ext * c.g.gr) - h.elo[sd] + 1;
      }
      const long long len = h.elen[sd];
      h.mode[sd] = !on ? 0 : (nbr[sd] == cf.rank ? 1 : 2);
      h.count[sd] = on ? (long long)n * h.nk * len * nex : 0;
      h.seg[sd] = off;
      off += h.count[sd]; tot += h.count[sd];
      if (h.mode[sd] == 1) has_local = true;
      if (h.mode[sd] == 2) has_remote = true;
    }
    if (has_local) { if (launch_halo(c, HM_LOCAL, h, nullptr, tot)) return 1; }
    if (has_remote) {
      if (!c.nccl_comm) return fail("halo_exchange: remote neighbour but moloch_b200_comm_init was not called");
      if ((size_t)off > c.halo_buf_doubles) {
        if (sync_stream(c)) return 1;
        if (c.sendbuf) cudaFree(c.sendbuf);
        if (c.recvbuf) cudaFree(c.recvbuf);
        c.halo_buf_doubles = (size_t)off + (size_t)off / 4;
        MB_CUDA(cudaMalloc(&c.sendbuf, c.halo_buf_doubles * sizeof(double)));
        MB_CUDA(cudaMalloc(&c.recvbuf, c.halo_buf_doubles * sizeof(double)));
      }
      if (launch_halo(c, HM_PACK, h, c.sendbuf, tot)) return 1;
      ncclComm_t comm = (ncclComm_t)c.nccl_comm;
      MB_NCCL(g_nccl.GroupStart());
      // sends in side order L,R,B,T; receives in order R,L,T,B so that two
      // messages between the same pair of ranks (periodic, 2 ranks in a
      // direction) match: my "to left" is the peer's "from right".
      for (int sd = 0; sd < 4; ++sd)
        if (h.mode[sd] == 2)
          MB_NCCL(g_nccl.Send(c.sendbuf + h.seg[sd], (size_t)h.count[sd], ncclDouble, nbr[sd], comm, c.stream));
      const int rorder[4] = {1, 0, 3, 2};
      for (int q = 0; q < 4; ++q) {
        const int sd = rorder[q];
        if (h.mode[sd] == 2)
          MB_NCCL(g_nccl.Recv(c.recvbuf + h.seg[sd], (size_t)h.count[sd], ncclDouble, nbr[sd], comm, c.stream));
      }
      MB_NCCL(g_nccl.GroupEnd());
      if (launch_halo(c, HM_UNPACK, h, c.recvbuf, tot)) return 1;
    }
  }
  return 0;
}

}  // namespace mb
