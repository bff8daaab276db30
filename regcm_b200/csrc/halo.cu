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

void halo_free(Ctx& c) {
  if (c.nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)c.nccl_comm);
  c.nccl_comm = nullptr;
  if (c.sendbuf) cudaFree(c.sendbuf);
  if (c.recvbuf) cudaFree(c.recvbuf);
  c.sendbuf = c.recvbuf = nullptr;
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

// `ext` widens the edge run of every side by `ext` ghost points at each end
// that has a neighbour: an lr exchange followed by a bt exchange with ext > 0
// (or the other way round) also fills the corner ghosts.
int halo_exchange(Ctx& c, const HaloItem* items, int nitems, int stag, int nex, bool lr, bool bt, int ext) {
  const moloch_b200_config& cf = c.cfg;
  const int nbr[4] = {cf.nbr_left, cf.nbr_right, cf.nbr_bottom, cf.nbr_top};
  bool any = false;
  for (int sd = 0; sd < 4; ++sd) if (nbr[sd] >= 0 && ((sd < 2) ? lr : bt)) any = true;
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
        MB_CUDA(cudaStreamSynchronize(c.stream));
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
