// kernels.cu -- sm_100a kernels of the MOLOCH dycore step.
//
// Every kernel restates one loop nest (or a fusion of adjacent loop nests) of
// /root/reference/Main/mod_moloch.F90; the cited lines are of that file.
// Arithmetic is FP64 with the reference's operation order and this file is
// compiled with -fmad=false, so +,-,*,/ results are bit-identical to a
// non-contracting CPU evaluation of the Fortran.
#include <algorithm>
#include "common.cuh"

namespace mb {

// ---------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ double dmax(double a, double b) { return (a < b) ? b : a; }
__device__ __forceinline__ double dmin(double a, double b) { return (b < a) ? b : a; }

// local_flow_param :1571-1590 (minnum is a real(rk4) parameter)
__device__ __forceinline__ double flow_param(double num, double den) {
  const double minden = 1.0e-30;
  const double minnum = (double)1.0e-30f;
  if (fabs(den) < minden) return (fabs(num) < minnum) ? 1.0 : 0.0;
  return num / den;
}
// zphi of the WAF limiter :882-883
__device__ __forceinline__ double waf_phi(double rr, double zamu, double is) {
  const double b = dmax(0.0, dmin(2.0, dmax(rr, dmin(2.0 * rr, 1.0))));
  return is + zamu * b - is * b;
}

// Share/pfwsat.inc
__device__ __forceinline__ double pfwsat(double t, double p) {
  const double a0 = 0.611213476e+03, a1 = 0.444007856e+02, a2 = 0.143064234e+01, a3 = 0.264461437e-01,
               a4 = 0.305903558e-03, a5 = 0.196237241e-05, a6 = 0.892344772e-08, a7 = -0.373208410e-10,
               a8 = 0.209339997e-13;
  const double c0 = 0.611123516e+03, c1 = 0.503109514e+02, c2 = 0.188369801e+01, c3 = 0.420547422e-01,
               c4 = 0.614396778e-03, c5 = 0.602780717e-05, c6 = 0.387940929e-07, c7 = 0.149436277e-09,
               c8 = 0.262655803e-12;
  const double td = dmin(dmax(t - tzero, -75.0), 100.0);
  double es;
  if (td >= 0.0)
    es = dmin(a0 + td * (a1 + td * (a2 + td * (a3 + td * (a4 + td * (a5 + td * (a6 + td * (a7 + td * a8))))))),
              0.15 * p);
  else
    es = dmin(c0 + td * (c1 + td * (c2 + td * (c3 + td * (c4 + td * (c5 + td * (c6 + td * (c7 + td * c8))))))),
              0.15 * p);
  return ep2 * (es / (p - es));
}

// Main/mpplib/mod_runparams.F90:184-192
__constant__ double c_qxcheckval[10] = {1.0e-8, 1.0e-16, 1.0e-16, 1.0e-16, 1.0e-16,
                                        1.0e-16, 1.0e-16, 1.0e10, 100.0, 0.01};
__constant__ double c_qxzeroval[10] = {1.0e-8, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0e10, 100.0, 0.01};

#define IX(j, i, k) gidx(g, (j), (i), (k))
#define IX2(j, i) gidx2(g, (j), (i))

constexpr int BX = 32, BY = 8;
static inline dim3 grid3(int nj, int ni, int nk) {
  return dim3((unsigned)((nj + BX - 1) / BX), (unsigned)((ni + BY - 1) / BY), (unsigned)nk);
}
#define THREAD_JIK(jlo, ilo, klo)                      \
  const int j = (jlo) + blockIdx.x * BX + threadIdx.x; \
  const int i = (ilo) + blockIdx.y * BY + threadIdx.y; \
  const int k = (klo) + blockIdx.z;

__device__ __forceinline__ bool in_box(int j, int i, int j1, int j2, int i1, int i2) {
  return j >= j1 && j <= j2 && i >= i1 && i <= i2;
}

// moist factor of temp_to_tvirt/tvirt_to_temp :1608-1648
__device__ __forceinline__ double moist_factor(const Geo& g, const double* qx, long long id) {
  const long long sp = (long long)g.kz * g.plane;
  if (g.ipptls > 0) {
    if (g.ipptls > 1) return 1.0 + ep1 * qx[id] - qx[id + sp] - qx[id + 2 * sp] - qx[id + 3 * sp] - qx[id + 4 * sp];
    return 1.0 + ep1 * qx[id] - qx[id + sp];
  }
  return 1.0 + ep1 * qx[id];
}

// ---------------------------------------------------------------------------
// K1  tetavf = 0.5*(tetav(k-1)+tetav(k))                              :564-566
// ---------------------------------------------------------------------------
__global__ void moloch_tetavf_init(Geo g, const double* __restrict__ tetav, double* __restrict__ tetavf) {
  THREAD_JIK(g.jce1, g.ice1, 2)
  if (j > g.jce2 || i > g.ice2) return;
  const long long id = IX(j, i, k);
  tetavf[id] = 0.5 * (tetav[id - g.plane] + tetav[id]);
}
int k_tetavf_init(Ctx& c) {
  const Geo& g = c.g;
  LaunchScope ls(c, KID_TETAVF);
  moloch_tetavf_init<<<grid3(g.jce2 - g.jce1 + 1, g.ice2 - g.ice1 + 1, g.kz - 1), dim3(BX, BY), 0, c.stream>>>(
      g, c.f[MB_TETAV].p, c.f[MB_TETAVF].p);
  MB_CUDA(cudaGetLastError());
  return 0;
}

// K2..K6 and K10 (sound_div, uvupdate2): kernels_sound.cu

// ---------------------------------------------------------------------------
// K7+K8+K9  vertical part of the divergence, implicit w (Thomas sweeps), new
// Exner function                                                     :627-671
// One thread per column; the finished divergence of the column is parked in
// shared memory (thread-private slots, no barriers).
// ---------------------------------------------------------------------------
// WS_NJ columns per CTA, WS_THREADS threads: 32 x 256 on large grids, 16 x 128 on
// small per-GPU grids (strong scaling) so that the CTAs still cover all SMs.
template <int WS_NJ, int WS_THREADS>
__global__ void __launch_bounds__(WS_THREADS)
moloch_wsolve(Geo g, const double* __restrict__ zdiv, double* s, double* __restrict__ w,
              double* __restrict__ pai, const double* __restrict__ tetav, double* __restrict__ tetavf,
              const double* __restrict__ fmz, const double* __restrict__ fmzf,
              const double* __restrict__ bdywtw, const double* __restrict__ ffilt, double dts, double dtrdz,
              double zcs2, int last, PushCtl pc, EdgePush ep) {
  // CTA = 32 columns x all levels.  Everything that does not depend on the
  // recurrence (finished divergence, explicit w, matrix coefficients) is
  // computed by all threads in parallel over k; only the two Thomas sweeps run
  // serially in k, by one warp, out of shared memory.
  extern __shared__ double sm[];
  const int kz = g.kz;
  double* ZD = sm;                      // finished divergence, levels 1..kz
  double* W = ZD + kz * WS_NJ;          // w / explicit w, levels 1..kz+1
  double* ZU = W + (kz + 1) * WS_NJ;    // zu, then wwkw
  double* ZC = ZU + kz * WS_NJ;         // zd coefficient
  const int nj = g.jci2 - g.jci1 + 1, ni = g.ici2 - g.ici1 + 1;
  const long long ncol = (long long)nj * ni;
  const int lane = threadIdx.x % WS_NJ, row0 = threadIdx.x / WS_NJ;
  constexpr int NR = WS_THREADS / WS_NJ;
  const long long col = (long long)blockIdx.x * WS_NJ + lane;
  const bool valid = col < ncol;
  const long long colc = valid ? col : ncol - 1;
  const int i = g.ici1 + (int)(colc / nj), j = g.jci1 + (int)(colc % nj);
  const long long base = gidx(g, j, i, 1);
  const long long pl = g.plane;
  // :627-630 and the old w
  for (int k = 1 + row0; k <= kz + 1; k += NR) {
    const long long id = base + (k - 1) * pl;
    W[(k - 1) * WS_NJ + lane] = w[id];
    if (k <= kz)
      ZD[(k - 1) * WS_NJ + lane] = zdiv[id] + bdywtw[id] * dtrdz * fmz[id] * (s[id] - s[id + pl]);
  }
  __syncthreads();
  // :636-649 explicit part and tridiagonal coefficients, parallel in k
  for (int k = 2 + row0; k <= kz; k += NR) {
    const long long id = base + (k - 1) * pl;
    const int o = (k - 1) * WS_NJ + lane;
    const double tv_k = tetav[id], tv_km1 = tetav[id - pl];
    const double pai_k = pai[id], pai_km1 = pai[id - pl];
    const double wk = W[o], fmzfk = fmzf[id];
    const double tf = tetavf[id] - wk * fmzfk * dtrdz * (tv_km1 - tv_k);
    if (valid) tetavf[id] = tf;
    const double zrom1w = cpd * tf * fmzfk;
    double zwexpl = wk - zrom1w * dtrdz * (pai_km1 - pai_k) - egrav * dts;
    zwexpl = zwexpl + rdrcv * zrom1w * dtrdz * (pai_km1 * ZD[o - WS_NJ] - pai_k * ZD[o]);
    ZU[o] = zcs2 * fmz[id - pl] * zrom1w * pai_km1 + ffilt[k];
    ZC[o] = zcs2 * fmz[id] * zrom1w * pai_k + ffilt[k];
    W[o] = zwexpl;
  }
  __syncthreads();
  // :650-664 Thomas sweeps, serial in k
  if (row0 == 0) {
    double wkp1 = W[kz * WS_NJ + lane];  // w(kzp1)
    double wwkp1 = 0.0;                  // wwkw(kzp1) :1055-1057
    for (int k = kz; k >= 2; --k) {
      const int o = (k - 1) * WS_NJ + lane;
      const double zu = ZU[o], zd = ZC[o];
      const double zrapp = 1.0 / (1.0 + zd + zu - zd * wwkp1);
      wkp1 = zrapp * (W[o] + zd * wkp1);
      wwkp1 = zrapp * zu;
      W[o] = wkp1;
      ZU[o] = wwkp1;
    }
    double wkm1 = W[lane];  // w(1)
    for (int k = 2; k <= kz; ++k) {
      const int o = (k - 1) * WS_NJ + lane;
      wkm1 = W[o] + ZU[o] * wkm1;
      W[o] = wkm1;
    }
  }
  __syncthreads();
  // :668-671 new Exner function; after the last sub-step also :728-734
  if (valid) {
    for (int k = 1 + row0; k <= kz; k += NR) {
      const long long id = base + (k - 1) * pl;
      const int o = (k - 1) * WS_NJ + lane;
      const double wk = W[o], wk1 = W[o + WS_NJ];
      const double pnew = pai[id] * (1.0 - rdrcv * (ZD[o] + (dtrdz * fmz[id] * (wk - wk1))));
      pai[id] = pnew;
      if (pc.mask) edge_push(pc, ep, j, i, k, pnew);
      if (k >= 2) w[id] = wk;
      if (last) s[id] = (k >= 2) ? (wk + s[id]) * fmzf[id] : 0.0;
    }
    if (last && row0 == 0) s[base + (long long)kz * pl] = 0.0;
  }
  halo_producer_done(pc, blockIdx.x, gridDim.x, WS_NJ, nj, ni, 1);
}
template <int WS_NJ, int WS_THREADS>
static int launch_wsolve(Ctx& c, double dts, bool last, long long ncol, const PushCtl& pc, const EdgePush& ep) {
  const Geo& g = c.g;
  const double dtrdz = dts * c.rdzita;
  const double zcs2 = (dtrdz * dtrdz) * rdrcv;
  const size_t smem = (size_t)(4 * g.kz + 1) * WS_NJ * sizeof(double);
  if (smem > 227 * 1024) return fail("wsolve: kz too large for the shared-memory column tile");
  const double* zsrc = c.cfg.mo_divfilter ? c.zdiv2b : c.f[MB_ZDIV2].p;
  MB_CUDA(cudaFuncSetAttribute(moloch_wsolve<WS_NJ, WS_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  LaunchScope ls(c, KID_WSOLVE);
  moloch_wsolve<WS_NJ, WS_THREADS><<<(unsigned)((ncol + WS_NJ - 1) / WS_NJ), WS_THREADS, smem, c.stream>>>(
      g, zsrc, c.f[MB_S].p, c.f[MB_W].p, c.f[MB_PAI].p, c.f[MB_TETAV].p, c.f[MB_TETAVF].p, c.f[MB_FMZ].p,
      c.f[MB_FMZF].p, c.f[MB_BDYWTW].p, c.prof[MB_FFILT], dts, dtrdz, zcs2, last ? 1 : 0, pc, ep);
  MB_CUDA(cudaGetLastError());
  return 0;
}
int k_wsolve5(Ctx& c, double dts, bool last, const PushCtl& pc, const EdgePush& ep);
int k_wsolve6(Ctx& c, double dts, bool last, const PushCtl& pc, const EdgePush& ep);
int k_wsolve8(Ctx& c, double dts, bool last, const PushCtl& pc, const EdgePush& ep);
int k_wsolve_tm(Ctx& c, double dts, bool last, const PushCtl& pc, const EdgePush& ep);
int k_wsolve(Ctx& c, double dts, bool last, const PushCtl* pcp, const EdgePush* epp) {
  const PushCtl pc = pcp ? *pcp : PushCtl{};
  const EdgePush ep = epp ? *epp : EdgePush{};
  if (c.wsolve_impl == 5) return k_wsolve5(c, dts, last, pc, ep);
  if (c.wsolve_impl == 6 || c.wsolve_impl == 7) return k_wsolve6(c, dts, last, pc, ep);
  if (c.wsolve_impl >= 11 && c.wsolve_impl <= 13 && c.g.kz + 1 <= 42) return k_wsolve_tm(c, dts, last, pc, ep);
  if (c.wsolve_impl >= 8 && c.wsolve_impl <= 13) return k_wsolve8(c, dts, last, pc, ep);
  const Geo& g = c.g;
  const long long ncol = (long long)(g.jci2 - g.jci1 + 1) * (g.ici2 - g.ici1 + 1);
  const bool small = (ncol + 31) / 32 < 148 * 5 * 3;   // fewer than three waves of 32-column CTAs
  return small ? launch_wsolve<16, 128>(c, dts, last, ncol, pc, ep) : launch_wsolve<32, 256>(c, dts, last, ncol, pc, ep);
}

// ---------------------------------------------------------------------------
// K7+K8+K9, thread-per-column variant.  The Thomas recurrences are serial in k
// but independent between columns, so every thread sweeps its own column: no
// barriers, no idle warps.  Memory latency is hidden by a ring of D levels that
// is always in flight (a slot is refilled as soon as it has been consumed);
// the sweep results wait in thread-private shared-memory slots ([k][lane],
// conflict-free) for the upward pass.
// The ring is fed by cp.async (LDGSTS) into thread-private shared-memory slots:
// register-targeted loads of a software pipeline share the warp's few
// scoreboards, so waiting for the oldest level also waits for the youngest one
// and the prefetch distance collapses; cp.async commit groups retire in order,
// so D levels really stay in flight.
// ---------------------------------------------------------------------------
#ifdef MB_HOST_EMU   // tests/emu: deferred copies that land at the matching wait_group
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) { emu::cp_async_enqueue(smem_dst, gsrc, 8); }
__device__ __forceinline__ void cp_async_commit() { emu::cp_async_commit(); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { emu::cp_async_wait(N); }
#else
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif

template <int D>
__global__ void __launch_bounds__(32)
moloch_wsolve5(Geo g, const double* __restrict__ zdiv, double* s, double* w, double* pai,
               const double* __restrict__ tetav, double* tetavf, const double* __restrict__ fmz,
               const double* __restrict__ fmzf, const double* __restrict__ bdywtw,
               const double* __restrict__ ffilt, double dts, double dtrdz, double zcs2, int last, PushCtl pc,
               EdgePush ep) {
  extern __shared__ double sm[];
  const int kz = g.kz;
  double* WP = sm;                        // w after the downward sweep, rows k = 0..kz
  double* WW = WP + (kz + 1) * 32;        // wwkw
  double* ZF = WW + (kz + 1) * 32;        // finished divergence
  double* RING = ZF + (kz + 1) * 32;      // D slots x 9 values x 32 lanes
  const int nj = g.jci2 - g.jci1 + 1, ni = g.ici2 - g.ici1 + 1;
  const long long ncol = (long long)nj * ni;
  const int lane = threadIdx.x;
  const long long col = (long long)blockIdx.x * 32 + lane;
  const bool valid = col < ncol;
  const long long colc = valid ? col : ncol - 1;
  const int i = g.ici1 + (int)(colc / nj), j = g.jci1 + (int)(colc % nj);
  const long long pl = g.plane;
  const long long base = gidx(g, j, i, 1) - pl;   // level k at base + k*pl
  auto fetch = [&](int slot, int k) {
    double* r = RING + slot * (9 * 32) + lane;
    const long long id = base + k * pl;
    cp_async8(r, w + id); cp_async8(r + 32, zdiv + id); cp_async8(r + 64, bdywtw + id);
    cp_async8(r + 96, fmz + id); cp_async8(r + 128, s + id); cp_async8(r + 160, tetav + id);
    cp_async8(r + 192, pai + id); cp_async8(r + 224, fmzf + id); cp_async8(r + 256, tetavf + id);
  };
  // ---- downward pass ----
#pragma unroll
  for (int q = 0; q < D; ++q) {
    if (kz - q >= 1) fetch(q, kz - q);
    cp_async_commit();
  }
  double wkp1 = w[base + (kz + 1) * pl];   // w(kzp1)
  const double w_bottom = wkp1;
  double wwkp1 = 0.0;                       // wwkw(kzp1) :1055-1057
  double s_below = s[base + (kz + 1) * pl];
  double p_w = 0.0, p_tf = 0.0, p_ff = 0.0, p_tv = 0.0, p_pa = 0.0, p_fm = 0.0, p_zd = 0.0;  // level m+1
  double w1 = 0.0;
  for (int t0 = 0; t0 < kz; t0 += D) {
#pragma unroll
    for (int q = 0; q < D; ++q) {
      const int m = kz - (t0 + q);
      cp_async_wait<D - 1>();
      if (m >= 1) {
        const double* r = RING + q * (9 * 32) + lane;
        const double Lw = r[0], Lzdiv = r[32], Lbw = r[64], Lfm = r[96], Ls = r[128], Ltv = r[160],
                     Lpa = r[192], Lff = r[224], Ltf = r[256];
        const double zdm = Lzdiv + Lbw * dtrdz * Lfm * (Ls - s_below);
        s_below = Ls;
        ZF[m * 32 + lane] = zdm;
        if (m < kz) {
          const int k = m + 1;
          const double tfn = p_tf - p_w * p_ff * dtrdz * (Ltv - p_tv);
          if (valid) tetavf[base + k * pl] = tfn;
          const double zrom1w = cpd * tfn * p_ff;
          double zwexpl = p_w - zrom1w * dtrdz * (Lpa - p_pa) - egrav * dts;
          zwexpl = zwexpl + rdrcv * zrom1w * dtrdz * (Lpa * zdm - p_pa * p_zd);
          const double fk = ffilt[k];
          const double zu = zcs2 * Lfm * zrom1w * Lpa + fk;
          const double zd = zcs2 * p_fm * zrom1w * p_pa + fk;
          const double zrapp = 1.0 / (1.0 + zd + zu - zd * wwkp1);
          wkp1 = zrapp * (zwexpl + zd * wkp1);
          wwkp1 = zrapp * zu;
          WP[k * 32 + lane] = wkp1;
          WW[k * 32 + lane] = wwkp1;
        }
        p_w = Lw; p_tf = Ltf; p_ff = Lff; p_tv = Ltv; p_pa = Lpa; p_fm = Lfm; p_zd = zdm;
        if (m == 1) w1 = Lw;
        if (m - D >= 1) fetch(q, m - D);
      }
      cp_async_commit();
    }
  }
  cp_async_wait<0>();
  // ---- upward pass ----
  auto fetchup = [&](int slot, int k) {   // level k needs pai, fmz of level k-1; s, fmzf of level k
    double* r = RING + slot * (9 * 32) + lane;
    const long long id = base + k * pl;
    cp_async8(r, pai + id - pl); cp_async8(r + 32, fmz + id - pl);
    if (last) { cp_async8(r + 64, s + id); cp_async8(r + 96, fmzf + id); }
  };
#pragma unroll
  for (int q = 0; q < D; ++q) {
    if (2 + q <= kz + 1) fetchup(q, 2 + q);
    cp_async_commit();
  }
  double wkm1 = w1;
  for (int t0 = 0; t0 < kz; t0 += D) {
#pragma unroll
    for (int q = 0; q < D; ++q) {
      const int k = 2 + t0 + q;
      cp_async_wait<D - 1>();
      if (k <= kz + 1) {
        const double* r = RING + q * (9 * 32) + lane;
        const double Upa = r[0], Ufm = r[32];
        const double wk = (k <= kz) ? WP[k * 32 + lane] + WW[k * 32 + lane] * wkm1 : w_bottom;
        if (valid) {
          const long long id = base + k * pl;
          const double pnew = Upa * (1.0 - rdrcv * (ZF[(k - 1) * 32 + lane] + (dtrdz * Ufm * (wkm1 - wk))));
          pai[id - pl] = pnew;
          if (pc.mask) edge_push(pc, ep, j, i, k - 1, pnew);
          if (k <= kz) {
            w[id] = wk;
            if (last) s[id] = (wk + r[64]) * r[96];
          }
        }
        wkm1 = wk;
        if (k + D <= kz + 1) fetchup(q, k + D);
      }
      cp_async_commit();
    }
  }
  cp_async_wait<0>();
  if (last && valid) { s[base + pl] = 0.0; s[base + (kz + 1) * pl] = 0.0; }
  halo_producer_done(pc, blockIdx.x, gridDim.x, 32, g.jci2 - g.jci1 + 1, g.ici2 - g.ici1 + 1, 1);
}
template <int D>
static int launch_wsolve5(Ctx& c, double dts, bool last, const PushCtl& pc, const EdgePush& ep) {
  const Geo& g = c.g;
  const long long ncol = (long long)(g.jci2 - g.jci1 + 1) * (g.ici2 - g.ici1 + 1);
  const double dtrdz = dts * c.rdzita;
  const double zcs2 = (dtrdz * dtrdz) * rdrcv;
  const size_t smem = (size_t)(3 * (g.kz + 1) + D * 9) * 32 * sizeof(double);
  if (smem > 227 * 1024) return fail("wsolve: kz too large for the shared-memory sweep slots");
  const double* zsrc = c.cfg.mo_divfilter ? c.zdiv2b : c.f[MB_ZDIV2].p;
  MB_CUDA(cudaFuncSetAttribute(moloch_wsolve5<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  LaunchScope ls(c, KID_WSOLVE);
  moloch_wsolve5<D><<<(unsigned)((ncol + 31) / 32), 32, smem, c.stream>>>(
      g, zsrc, c.f[MB_S].p, c.f[MB_W].p, c.f[MB_PAI].p, c.f[MB_TETAV].p, c.f[MB_TETAVF].p, c.f[MB_FMZ].p,
      c.f[MB_FMZF].p, c.f[MB_BDYWTW].p, c.prof[MB_FFILT], dts, dtrdz, zcs2, last ? 1 : 0, pc, ep);
  MB_CUDA(cudaGetLastError());
  return 0;
}
// ---------------------------------------------------------------------------
// K7+K8+K9, thread-per-column variant 6 (MOLOCH_B200_WSOLVE=6 / set_option("wsolve", 6); measured slower than
// variant 5 on the B200: 246 vs 217 us, r2a -- kept as a bit-identical A/B candidate).
// moloch_wsolve5 is bound by the latency of ONE warp per scheduler: its three sweep arrays (w', wwkw and the
// finished divergence) and the ring take 46 KB of shared memory per warp at kz = 41, so only 4 warps fit an SM.
// Here the finished divergence is not parked but recomputed in the upward pass from the same operands, in the
// same operation order (zdiv2, bdywtw, fmz, s(k), s(k+1): the lines were read by this very warp a few
// microseconds earlier and come from L2), and the ring is D = 4 deep: 31 KB per warp, 7 warps per SM.
// ---------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(32)
moloch_wsolve6(Geo g, const double* __restrict__ zdiv, double* s, double* w, double* pai,
               const double* __restrict__ tetav, double* tetavf, const double* __restrict__ fmz,
               const double* __restrict__ fmzf, const double* __restrict__ bdywtw,
               const double* __restrict__ ffilt, double dts, double dtrdz, double zcs2, int last, PushCtl pc,
               EdgePush ep) {
  extern __shared__ double sm[];
  const int kz = g.kz;
  double* WP = sm;                        // w after the downward sweep, rows k = 0..kz
  double* WW = WP + (kz + 1) * 32;        // wwkw
  double* RING = WW + (kz + 1) * 32;      // D slots x 9 values x 32 lanes
  const int nj = g.jci2 - g.jci1 + 1, ni = g.ici2 - g.ici1 + 1;
  const long long ncol = (long long)nj * ni;
  const int lane = threadIdx.x;
  const long long col = (long long)blockIdx.x * 32 + lane;
  const bool valid = col < ncol;
  const long long colc = valid ? col : ncol - 1;
  const int i = g.ici1 + (int)(colc / nj), j = g.jci1 + (int)(colc % nj);
  const long long pl = g.plane;
  const long long base = gidx(g, j, i, 1) - pl;   // level k at base + k*pl
  auto fetch = [&](int slot, int k) {
    double* r = RING + slot * (9 * 32) + lane;
    const long long id = base + k * pl;
    cp_async8(r, w + id); cp_async8(r + 32, zdiv + id); cp_async8(r + 64, bdywtw + id);
    cp_async8(r + 96, fmz + id); cp_async8(r + 128, s + id); cp_async8(r + 160, tetav + id);
    cp_async8(r + 192, pai + id); cp_async8(r + 224, fmzf + id); cp_async8(r + 256, tetavf + id);
  };
  // ---- downward pass (as in moloch_wsolve5, without the ZF slots) ----
#pragma unroll
  for (int q = 0; q < D; ++q) {
    if (kz - q >= 1) fetch(q, kz - q);
    cp_async_commit();
  }
  double wkp1 = w[base + (kz + 1) * pl];   // w(kzp1)
  const double w_bottom = wkp1;
  double wwkp1 = 0.0;                       // wwkw(kzp1) :1055-1057
  double s_below = s[base + (kz + 1) * pl];
  double p_w = 0.0, p_tf = 0.0, p_ff = 0.0, p_tv = 0.0, p_pa = 0.0, p_fm = 0.0, p_zd = 0.0;  // level m+1
  double w1 = 0.0;
  for (int t0 = 0; t0 < kz; t0 += D) {
#pragma unroll
    for (int q = 0; q < D; ++q) {
      const int m = kz - (t0 + q);
      cp_async_wait<D - 1>();
      if (m >= 1) {
        const double* r = RING + q * (9 * 32) + lane;
        const double Lw = r[0], Lzdiv = r[32], Lbw = r[64], Lfm = r[96], Ls = r[128], Ltv = r[160],
                     Lpa = r[192], Lff = r[224], Ltf = r[256];
        const double zdm = Lzdiv + Lbw * dtrdz * Lfm * (Ls - s_below);
        s_below = Ls;
        if (m < kz) {
          const int k = m + 1;
          const double tfn = p_tf - p_w * p_ff * dtrdz * (Ltv - p_tv);
          if (valid) tetavf[base + k * pl] = tfn;
          const double zrom1w = cpd * tfn * p_ff;
          double zwexpl = p_w - zrom1w * dtrdz * (Lpa - p_pa) - egrav * dts;
          zwexpl = zwexpl + rdrcv * zrom1w * dtrdz * (Lpa * zdm - p_pa * p_zd);
          const double fk = ffilt[k];
          const double zu = zcs2 * Lfm * zrom1w * Lpa + fk;
          const double zd = zcs2 * p_fm * zrom1w * p_pa + fk;
          const double zrapp = 1.0 / (1.0 + zd + zu - zd * wwkp1);
          wkp1 = zrapp * (zwexpl + zd * wkp1);
          wwkp1 = zrapp * zu;
          WP[k * 32 + lane] = wkp1;
          WW[k * 32 + lane] = wwkp1;
        }
        p_w = Lw; p_tf = Ltf; p_ff = Lff; p_tv = Ltv; p_pa = Lpa; p_fm = Lfm; p_zd = zdm;
        if (m == 1) w1 = Lw;
        if (m - D >= 1) fetch(q, m - D);
      }
      cp_async_commit();
    }
  }
  cp_async_wait<0>();
  // ---- upward pass: level k needs pai, fmz, zdiv, bdywtw of level k-1 and s, fmzf of level k ----
  auto fetchup = [&](int slot, int k) {
    double* r = RING + slot * (9 * 32) + lane;
    const long long id = base + k * pl;
    cp_async8(r, pai + id - pl); cp_async8(r + 32, fmz + id - pl); cp_async8(r + 64, s + id);
    cp_async8(r + 128, zdiv + id - pl); cp_async8(r + 160, bdywtw + id - pl);
    if (last) cp_async8(r + 96, fmzf + id);
  };
  double s_km1 = s[base + pl];              // s(1), as the downward pass saw it
#pragma unroll
  for (int q = 0; q < D; ++q) {
    if (2 + q <= kz + 1) fetchup(q, 2 + q);
    cp_async_commit();
  }
  double wkm1 = w1;
  for (int t0 = 0; t0 < kz; t0 += D) {
#pragma unroll
    for (int q = 0; q < D; ++q) {
      const int k = 2 + t0 + q;
      cp_async_wait<D - 1>();
      if (k <= kz + 1) {
        const double* r = RING + q * (9 * 32) + lane;
        const double Upa = r[0], Ufm = r[32], Us = r[64], Uzdiv = r[128], Ubw = r[160];
        // the finished divergence of level k-1, exactly as the downward pass computed it
        const double zdm = Uzdiv + Ubw * dtrdz * Ufm * (s_km1 - Us);
        s_km1 = Us;
        const double wk = (k <= kz) ? WP[k * 32 + lane] + WW[k * 32 + lane] * wkm1 : w_bottom;
        if (valid) {
          const long long id = base + k * pl;
          const double pnew = Upa * (1.0 - rdrcv * (zdm + (dtrdz * Ufm * (wkm1 - wk))));
          pai[id - pl] = pnew;
          if (pc.mask) edge_push(pc, ep, j, i, k - 1, pnew);
          if (k <= kz) {
            w[id] = wk;
            if (last) s[id] = (wk + Us) * r[96];
          }
        }
        wkm1 = wk;
        if (k + D <= kz + 1) fetchup(q, k + D);
      }
      cp_async_commit();
    }
  }
  cp_async_wait<0>();
  if (last && valid) { s[base + pl] = 0.0; s[base + (kz + 1) * pl] = 0.0; }
  halo_producer_done(pc, blockIdx.x, gridDim.x, 32, g.jci2 - g.jci1 + 1, g.ici2 - g.ici1 + 1, 1);
}
template <int D>
static int launch_wsolve6(Ctx& c, double dts, bool last, const PushCtl& pc, const EdgePush& ep) {
  const Geo& g = c.g;
  const long long ncol = (long long)(g.jci2 - g.jci1 + 1) * (g.ici2 - g.ici1 + 1);
  const double dtrdz = dts * c.rdzita;
  const double zcs2 = (dtrdz * dtrdz) * rdrcv;
  const size_t smem = (size_t)(2 * (g.kz + 1) + D * 9) * 32 * sizeof(double);
  if (smem > 227 * 1024) return fail("wsolve: kz too large for the shared-memory sweep slots");
  const double* zsrc = c.cfg.mo_divfilter ? c.zdiv2b : c.f[MB_ZDIV2].p;
  MB_CUDA(cudaFuncSetAttribute(moloch_wsolve6<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  LaunchScope ls(c, KID_WSOLVE);
  moloch_wsolve6<D><<<(unsigned)((ncol + 31) / 32), 32, smem, c.stream>>>(
      g, zsrc, c.f[MB_S].p, c.f[MB_W].p, c.f[MB_PAI].p, c.f[MB_TETAV].p, c.f[MB_TETAVF].p, c.f[MB_FMZ].p,
      c.f[MB_FMZF].p, c.f[MB_BDYWTW].p, c.prof[MB_FFILT], dts, dtrdz, zcs2, last ? 1 : 0, pc, ep);
  MB_CUDA(cudaGetLastError());
  return 0;
}
// ---------------------------------------------------------------------------
// K7+K8+K9, variant 8 (round 2).  What round 1's variants taught (profiles/): the column kernel is bound by the
// bytes it keeps in flight -- HBM answers in 2-3 us under load, so a warp that waits for level k while only
// D levels are under way moves D*2.3 KB per latency -- and the upward pass, which needs two to five values per
// level, ran at the same levels-per-latency as the downward pass with a fifth of its bytes.  Variant 8 therefore
//  * feeds the ring with 16-byte cp.async.cg (two columns per copy, L1 bypassed: an in-flight line does not
//    occupy the small L1 that is left beside 200 KB of shared memory; five copies per lane and level instead of
//    nine).  A warp's 32 columns are consecutive cells of ONE row starting on a 32-byte boundary (tiles by row),
//  * re-partitions the same ring memory for the upward pass into DU = 9*D/4 (9*D/6) slots of 4 (6) values:
//    13+ levels under way instead of 6,
//  * optionally (ZFS = false) recomputes the finished divergence in the upward pass like variant 6, which buys
//    a deeper downward ring in the same shared memory.
// Copies are issued by one lane for two columns, so a warp-level barrier separates a slot's arrival from its
// use and its use from its refill.  Arithmetic and operation order are those of variants 5/6: bit-identical.
// ---------------------------------------------------------------------------
#ifdef MB_HOST_EMU
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc) { emu::cp_async_enqueue(smem_dst, gsrc, 16); }
#else
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
#endif

template <int D, bool ZFS>
__global__ void __launch_bounds__(32)
moloch_wsolve8(Geo g, const double* __restrict__ zdiv, double* s, double* w, double* pai,
               const double* __restrict__ tetav, double* tetavf, const double* __restrict__ fmz,
               const double* __restrict__ fmzf, const double* __restrict__ bdywtw,
               const double* __restrict__ ffilt, double dts, double dtrdz, double zcs2, int last, int ntile_j,
               PushCtl pc, EdgePush ep) {
  extern __shared__ double sm[];
  constexpr int UV = ZFS ? 4 : 6;                    // values per level of the upward pass
  constexpr int DU = (9 * D) / UV;                   // its ring depth in the same memory
  const int kz = g.kz;
  double* WP = sm;                                   // w after the downward sweep, rows k = 0..kz
  double* WW = WP + (kz + 1) * 32;                   // wwkw
  double* ZF = WW + (kz + 1) * 32;                   // finished divergence (ZFS)
  double* RING = ZF + (ZFS ? (kz + 1) * 32 : 0);     // D slots x 9 values x 32 lanes
  const int lane = threadIdx.x;
  const int tj = blockIdx.x % ntile_j, ti = blockIdx.x / ntile_j;
  const int i = g.ici1 + ti, jf = g.jde1 + 32 * tj;  // jf - j0 = HJ + 32*tj: a 32-byte boundary
  const int j = jf + lane;
  const bool valid = (j >= g.jci1 && j <= g.jci2);
  const long long pl = g.plane;
  const long long rowb = gidx(g, jf, i, 1) - pl;     // level k of the tile's first column at rowb + k*pl
  // copy role of the lane: columns 2c, 2c+1 of the arrays of parity `half`
  const int half = lane >> 4, c2 = 2 * (lane & 15);
  const bool cok = (jf + c2 + 1 <= g.j0 + g.NJ - 1);  // the last tile of a row may reach beyond the padded box
  const double* d0 = half ? zdiv : w;
  const double* d1 = half ? fmz : bdywtw;
  const double* d2 = half ? tetav : s;
  const double* d3 = half ? fmzf : pai;
  auto fetch = [&](int slot, int k) {               // slot rows: w, zdiv, bdywtw, fmz, s, tetav, pai, fmzf, tetavf
    if (!cok) return;
    double* r = RING + slot * (9 * 32) + half * 32 + c2;
    const long long o = rowb + k * pl + c2;
    cp_async16(r, d0 + o); cp_async16(r + 64, d1 + o); cp_async16(r + 128, d2 + o); cp_async16(r + 192, d3 + o);
    if (!half) cp_async16(r + 256, tetavf + o);
  };
  // ---- downward pass ----
#pragma unroll
  for (int q = 0; q < D; ++q) {
    if (kz - q >= 1) fetch(q, kz - q);
    cp_async_commit();
  }
  const long long base = rowb + lane;                // this lane's column
  double wkp1 = w[base + (kz + 1) * pl];   // w(kzp1)
  const double w_bottom = wkp1;
  double wwkp1 = 0.0;                       // wwkw(kzp1) :1055-1057
  double s_below = s[base + (kz + 1) * pl];
  double p_w = 0.0, p_tf = 0.0, p_ff = 0.0, p_tv = 0.0, p_pa = 0.0, p_fm = 0.0, p_zd = 0.0;  // level m+1
  double w1 = 0.0;
  for (int t0 = 0; t0 < kz; t0 += D) {
#pragma unroll
    for (int q = 0; q < D; ++q) {
      const int m = kz - (t0 + q);
      cp_async_wait<D - 1>();
      __syncwarp();                         // the slot was filled by several lanes
      if (m >= 1) {
        const double* r = RING + q * (9 * 32) + lane;
        const double Lw = r[0], Lzdiv = r[32], Lbw = r[64], Lfm = r[96], Ls = r[128], Ltv = r[160],
                     Lpa = r[192], Lff = r[224], Ltf = r[256];
        __syncwarp();                       // every lane has read the slot: it may be refilled
        if (m - D >= 1) fetch(q, m - D);
        const double zdm = Lzdiv + Lbw * dtrdz * Lfm * (Ls - s_below);
        s_below = Ls;
        if (ZFS) ZF[m * 32 + lane] = zdm;
        if (m < kz) {
          const int k = m + 1;
          const double tfn = p_tf - p_w * p_ff * dtrdz * (Ltv - p_tv);
          if (valid) tetavf[base + k * pl] = tfn;
          const double zrom1w = cpd * tfn * p_ff;
          double zwexpl = p_w - zrom1w * dtrdz * (Lpa - p_pa) - egrav * dts;
          zwexpl = zwexpl + rdrcv * zrom1w * dtrdz * (Lpa * zdm - p_pa * p_zd);
          const double fk = ffilt[k];
          const double zu = zcs2 * Lfm * zrom1w * Lpa + fk;
          const double zd = zcs2 * p_fm * zrom1w * p_pa + fk;
          const double zrapp = 1.0 / (1.0 + zd + zu - zd * wwkp1);
          wkp1 = zrapp * (zwexpl + zd * wkp1);
          wwkp1 = zrapp * zu;
          WP[k * 32 + lane] = wkp1;
          WW[k * 32 + lane] = wwkp1;
        }
        p_w = Lw; p_tf = Ltf; p_ff = Lff; p_tv = Ltv; p_pa = Lpa; p_fm = Lfm; p_zd = zdm;
        if (m == 1) w1 = Lw;
      }
      cp_async_commit();
    }
  }
  cp_async_wait<0>();
  __syncwarp();
  // ---- upward pass: level k needs pai, fmz [zdiv, bdywtw] of level k-1 and s [, fmzf] of level k ----
  const double* u0 = half ? fmz : pai;               // of level k-1
  const double* u1 = half ? zdiv : s;                // (!ZFS) s of level k / zdiv of level k-1;  (ZFS, last) s / fmzf of level k
  auto fetchup = [&](int slot, int k) {             // slot rows: pai, fmz, s, (zdiv | fmzf), bdywtw, fmzf
    if (!cok) return;
    double* r = RING + slot * (UV * 32) + half * 32 + c2;
    const long long o = rowb + k * pl + c2;
    cp_async16(r, u0 + o - pl);
    if (ZFS) {
      if (last) cp_async16(r + 64, (half ? fmzf : s) + o);
    } else {
      cp_async16(r + 64, u1 + o - (half ? pl : 0));
      if (!half) cp_async16(r + 128, bdywtw + o - pl);
      else if (last) cp_async16(r + 128, fmzf + o);
    }
  };
  double s_km1 = s[base + pl];              // s(1), as the downward pass saw it
#pragma unroll
  for (int q = 0; q < DU; ++q) {
    if (2 + q <= kz + 1) fetchup(q, 2 + q);
    cp_async_commit();
  }
  double wkm1 = w1;
  for (int t0 = 0; t0 < kz; t0 += DU) {
#pragma unroll
    for (int q = 0; q < DU; ++q) {
      const int k = 2 + t0 + q;
      cp_async_wait<DU - 1>();
      __syncwarp();
      if (k <= kz + 1) {
        const double* r = RING + q * (UV * 32) + lane;
        const double Upa = r[0], Ufm = r[32];
        double Us = 0.0, Uff = 0.0, zdm;
        if (ZFS) {
          if (last) { Us = r[64]; Uff = r[96]; }
          zdm = ZF[(k - 1) * 32 + lane];
        } else {
          Us = r[64];
          const double Uzdiv = r[96], Ubw = r[128];
          if (last) Uff = r[160];
          // the finished divergence of level k-1, exactly as the downward pass computed it
          zdm = Uzdiv + Ubw * dtrdz * Ufm * (s_km1 - Us);
          s_km1 = Us;
        }
        __syncwarp();
        if (k + DU <= kz + 1) fetchup(q, k + DU);
        const double wk = (k <= kz) ? WP[k * 32 + lane] + WW[k * 32 + lane] * wkm1 : w_bottom;
        if (valid) {
          const long long id = base + k * pl;
          const double pnew = Upa * (1.0 - rdrcv * (zdm + (dtrdz * Ufm * (wkm1 - wk))));
          pai[id - pl] = pnew;
          if (pc.mask) edge_push(pc, ep, j, i, k - 1, pnew);
          if (k <= kz) {
            w[id] = wk;
            if (last) s[id] = (wk + Us) * Uff;
          }
        }
        wkm1 = wk;
      }
      cp_async_commit();
    }
  }
  cp_async_wait<0>();
  if (last && valid) { s[base + pl] = 0.0; s[base + (kz + 1) * pl] = 0.0; }
  halo_producer_done(pc, blockIdx.x, gridDim.x, 1, ntile_j, g.ici2 - g.ici1 + 1, 1);
}
template <int D, bool ZFS>
static int launch_wsolve8(Ctx& c, double dts, bool last, const PushCtl& pc, const EdgePush& ep) {
  const Geo& g = c.g;
  const int ntile_j = (g.jci2 - g.jde1 + 1 + 31) / 32, ni = g.ici2 - g.ici1 + 1;
  const double dtrdz = dts * c.rdzita;
  const double zcs2 = (dtrdz * dtrdz) * rdrcv;
  const size_t smem = (size_t)((ZFS ? 3 : 2) * (g.kz + 1) + D * 9) * 32 * sizeof(double);
  if (smem > 227 * 1024) return fail("wsolve: kz too large for the shared-memory sweep slots");
  const double* zsrc = c.cfg.mo_divfilter ? c.zdiv2b : c.f[MB_ZDIV2].p;
  MB_CUDA(cudaFuncSetAttribute(moloch_wsolve8<D, ZFS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  LaunchScope ls(c, KID_WSOLVE);
  moloch_wsolve8<D, ZFS><<<(unsigned)(ntile_j * ni), 32, smem, c.stream>>>(
      g, zsrc, c.f[MB_S].p, c.f[MB_W].p, c.f[MB_PAI].p, c.f[MB_TETAV].p, c.f[MB_TETAVF].p, c.f[MB_FMZ].p,
      c.f[MB_FMZF].p, c.f[MB_BDYWTW].p, c.prof[MB_FFILT], dts, dtrdz, zcs2, last ? 1 : 0, ntile_j, pc, ep);
  MB_CUDA(cudaGetLastError());
  return 0;
}
// ---------------------------------------------------------------------------
// K7+K8+K9, variant 11: the sweep arrays live in TENSOR MEMORY.  Variants 5..10 keep w', wwkw and the finished
// divergence of a column in shared memory (1 KB per column), which leaves room for four or five warps per SM:
// one warp per scheduler, and an in-order warp that has its scheduler to itself stalls on every dependent FP64
// instruction (ncu: 23 % of the issue slots used, "wait" the leading stall).  Blackwell's tensor memory -- 256 KB
// per SM, 128 lanes x 512 32-bit columns, otherwise idle in this FP64 stencil code -- is a per-thread scratchpad
// when it is addressed with the 32x32b shape: lane = thread, so a CTA of four warps that allocates 256 columns
// gives each of its 128 threads 128 doubles: the three sweep arrays of a 41-level column (126).  Shared memory
// then only holds the cp.async rings, two CTAs = eight warps fit an SM (two per scheduler), and the ring of a
// warp can be 8 deep.  tcgen05.st in the downward pass, tcgen05.ld one level ahead in the upward pass.
// Same arithmetic and operation order as variant 8: bit-identical.  kz <= 41 (else variant 8).
// ---------------------------------------------------------------------------
constexpr int TM_KZP = 42;          // rows per sweep array in tensor memory
constexpr int TM_COLS = 256;        // columns allocated per CTA (a power of two >= 2*3*TM_KZP)
#ifdef MB_HOST_EMU
struct Tmem {                       // tests/emu: the thread's 128 doubles are a local array
  double cell[3 * TM_KZP];
  __device__ void alloc(unsigned*) {}
  __device__ void release() {}
  __device__ void st(int idx, double v) { cell[idx] = v; }
  __device__ void st_done() {}
  __device__ void ld_issue(int idx, double& dst) { dst = cell[idx]; }
  __device__ void ld_wait() {}
};
#else
struct Tmem {
  unsigned base;                    // tensor-memory address of this warp's lane quarter, column 0
  __device__ __forceinline__ void alloc(unsigned* slot) {   // all threads of the CTA; slot: a word of shared memory
    if ((threadIdx.x >> 5) == 0) {
      const unsigned sa = (unsigned)__cvta_generic_to_shared(slot);
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sa), "r"(TM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    base = *slot + (((threadIdx.x >> 5) & 3u) << 21);       // lane field (bits 31:16) = 32 * (warp % 4)
  }
  __device__ __forceinline__ void release() {               // all threads of the CTA
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if ((threadIdx.x >> 5) == 0)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base & 0xffffu), "r"(TM_COLS) : "memory");
  }
  __device__ __forceinline__ void st(int idx, double v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(base + 2u * (unsigned)idx),
                 "r"(__double2loint(v)), "r"(__double2hiint(v)) : "memory");
  }
  __device__ __forceinline__ void st_done() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
  // the loaded words may only be read after ld_wait()
  int lo, hi;
  __device__ __forceinline__ void ld_issue2(int idx, int& l, int& h) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(l), "=r"(h) : "r"(base + 2u * (unsigned)idx) : "memory");
  }
  __device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
};
#endif

template <int D, bool PUSH>   // PUSH: pai's edge cells are stored into the neighbours' ghost cells (fused round)
__global__ void __launch_bounds__(128)
moloch_wsolve_tm(Geo g, const double* __restrict__ zdiv, double* s, double* w, double* pai,
                 const double* __restrict__ tetav, double* tetavf, const double* __restrict__ fmz,
                 const double* __restrict__ fmzf, const double* __restrict__ bdywtw,
                 const double* __restrict__ ffilt, double dts, double dtrdz, double zcs2, int last, int ntile_j,
                 int ntiles, PushCtl pc, EdgePush ep) {
  extern __shared__ double sm[];
  constexpr int UV = 4;                               // values per level of the upward pass
  constexpr int DU = (9 * D) / UV;                    // its ring depth in the same memory
  const int kz = g.kz;
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  double* RING = sm + wq * (D * 9 * 32);              // this warp's D slots x 9 values x 32 lanes
  unsigned* slot = reinterpret_cast<unsigned*>(sm + 4 * (D * 9 * 32));
  Tmem tm;
  tm.alloc(slot);
  const int tile = blockIdx.x * 4 + wq;
  if (tile < ntiles) {      // (warp-uniform; the warps of the CTA only meet again in tm.release())
    const int tj = tile % ntile_j, ti = tile / ntile_j;
    const int i = g.ici1 + ti, jf = g.jde1 + 32 * tj;   // jf - j0 = HJ + 32*tj: a 32-byte boundary
    const int j = jf + lane;
    const bool valid = (j >= g.jci1 && j <= g.jci2);
    const long long pl = g.plane;
    const long long rowb = gidx(g, jf, i, 1) - pl;      // level k of the tile's first column at rowb + k*pl
    const int half = lane >> 4, c2 = 2 * (lane & 15);
    const bool cok = (jf + c2 + 1 <= g.j0 + g.NJ - 1);
    const double* d0 = half ? zdiv : w;
    const double* d1 = half ? fmz : bdywtw;
    const double* d2 = half ? tetav : s;
    const double* d3 = half ? fmzf : pai;
    auto fetch = [&](int sl, int k) {                  // slot rows: w, zdiv, bdywtw, fmz, s, tetav, pai, fmzf, tetavf
      if (!cok) return;
      double* r = RING + sl * (9 * 32) + half * 32 + c2;
      const long long o = rowb + k * pl + c2;
      cp_async16(r, d0 + o); cp_async16(r + 64, d1 + o); cp_async16(r + 128, d2 + o); cp_async16(r + 192, d3 + o);
      if (!half) cp_async16(r + 256, tetavf + o);
    };
    // ---- downward pass ----
#pragma unroll
    for (int q = 0; q < D; ++q) {
      if (kz - q >= 1) fetch(q, kz - q);
      cp_async_commit();
    }
    const long long base = rowb + lane;                // this lane's column
    double wkp1 = w[base + (kz + 1) * pl];   // w(kzp1)
    const double w_bottom = wkp1;
    double wwkp1 = 0.0;                       // wwkw(kzp1) :1055-1057
    double s_below = s[base + (kz + 1) * pl];
    double p_w = 0.0, p_tf = 0.0, p_ff = 0.0, p_tv = 0.0, p_pa = 0.0, p_fm = 0.0, p_zd = 0.0;  // level m+1
    double w1 = 0.0;
    for (int t0 = 0; t0 < kz; t0 += D) {
#pragma unroll
      for (int q = 0; q < D; ++q) {
        const int m = kz - (t0 + q);
        cp_async_wait<D - 1>();
        __syncwarp();                         // the slot was filled by several lanes
        if (m >= 1) {
          const double* r = RING + q * (9 * 32) + lane;
          const double Lw = r[0], Lzdiv = r[32], Lbw = r[64], Lfm = r[96], Ls = r[128], Ltv = r[160],
                       Lpa = r[192], Lff = r[224], Ltf = r[256];
          __syncwarp();                       // every lane has read the slot: it may be refilled
          if (m - D >= 1) fetch(q, m - D);
          const double zdm = Lzdiv + Lbw * dtrdz * Lfm * (Ls - s_below);
          s_below = Ls;
          tm.st(2 * TM_KZP + m, zdm);
          if (m < kz) {
            const int k = m + 1;
            const double tfn = p_tf - p_w * p_ff * dtrdz * (Ltv - p_tv);
            if (valid) tetavf[base + k * pl] = tfn;
            const double zrom1w = cpd * tfn * p_ff;
            double zwexpl = p_w - zrom1w * dtrdz * (Lpa - p_pa) - egrav * dts;
            zwexpl = zwexpl + rdrcv * zrom1w * dtrdz * (Lpa * zdm - p_pa * p_zd);
            const double fk = ffilt[k];
            const double zu = zcs2 * Lfm * zrom1w * Lpa + fk;
            const double zd = zcs2 * p_fm * zrom1w * p_pa + fk;
            const double zrapp = 1.0 / (1.0 + zd + zu - zd * wwkp1);
            wkp1 = zrapp * (zwexpl + zd * wkp1);
            wwkp1 = zrapp * zu;
            tm.st(k, wkp1);
            tm.st(TM_KZP + k, wwkp1);
          }
          p_w = Lw; p_tf = Ltf; p_ff = Lff; p_tv = Ltv; p_pa = Lpa; p_fm = Lfm; p_zd = zdm;
          if (m == 1) w1 = Lw;
        }
        cp_async_commit();
      }
    }
    cp_async_wait<0>();
    tm.st_done();
    __syncwarp();
    // ---- upward pass: level k needs pai, fmz of level k-1 and (last) s, fmzf of level k from memory;
    //      w', wwkw of level k and the divergence of level k-1 from tensor memory, loaded one level ahead ----
    const double* u0 = half ? fmz : pai;
    auto fetchup = [&](int sl, int k) {               // slot rows: pai, fmz, s, fmzf
      if (!cok) return;
      double* r = RING + sl * (UV * 32) + half * 32 + c2;
      const long long o = rowb + k * pl + c2;
      cp_async16(r, u0 + o - pl);
      if (last) cp_async16(r + 64, (half ? fmzf : s) + o);
    };
#pragma unroll
    for (int q = 0; q < DU; ++q) {
      if (2 + q <= kz + 1) fetchup(q, 2 + q);
      cp_async_commit();
    }
#ifdef MB_HOST_EMU
    double n_wp = 0.0, n_ww = 0.0, n_zf = 0.0;
    if (2 <= kz) { tm.ld_issue(2, n_wp); tm.ld_issue(TM_KZP + 2, n_ww); }
    tm.ld_issue(2 * TM_KZP + 1, n_zf);
#else
    int a0 = 0, a1 = 0, b0 = 0, b1 = 0, c0 = 0, c1 = 0;
    if (2 <= kz) { tm.ld_issue2(2, a0, a1); tm.ld_issue2(TM_KZP + 2, b0, b1); }
    tm.ld_issue2(2 * TM_KZP + 1, c0, c1);
#endif
    double wkm1 = w1;
    for (int t0 = 0; t0 < kz; t0 += DU) {
#pragma unroll
      for (int q = 0; q < DU; ++q) {
        const int k = 2 + t0 + q;
        cp_async_wait<DU - 1>();
        __syncwarp();
        if (k <= kz + 1) {
          const double* r = RING + q * (UV * 32) + lane;
          const double Upa = r[0], Ufm = r[32];
          double Us = 0.0, Uff = 0.0;
          if (last) { Us = r[64]; Uff = r[96]; }
          __syncwarp();
          if (k + DU <= kz + 1) fetchup(q, k + DU);
          tm.ld_wait();
#ifdef MB_HOST_EMU
          const double c_wp = n_wp, c_ww = n_ww, zdm = n_zf;
          if (k + 1 <= kz) { tm.ld_issue(k + 1, n_wp); tm.ld_issue(TM_KZP + k + 1, n_ww); }
          if (k + 1 <= kz + 1) tm.ld_issue(2 * TM_KZP + k, n_zf);
#else
          const double c_wp = __hiloint2double(a1, a0), c_ww = __hiloint2double(b1, b0), zdm = __hiloint2double(c1, c0);
          if (k + 1 <= kz) { tm.ld_issue2(k + 1, a0, a1); tm.ld_issue2(TM_KZP + k + 1, b0, b1); }
          if (k + 1 <= kz + 1) tm.ld_issue2(2 * TM_KZP + k, c0, c1);
#endif
          const double wk = (k <= kz) ? c_wp + c_ww * wkm1 : w_bottom;
          if (valid) {
            const long long id = base + k * pl;
            const double pnew = Upa * (1.0 - rdrcv * (zdm + (dtrdz * Ufm * (wkm1 - wk))));
            pai[id - pl] = pnew;
            if (k <= kz) {
              w[id] = wk;
              if (last) s[id] = (wk + Us) * Uff;
            }
          }
          wkm1 = wk;
        }
        cp_async_commit();
      }
    }
    cp_async_wait<0>();
    tm.ld_wait();
    if (last && valid) { s[base + pl] = 0.0; s[base + (kz + 1) * pl] = 0.0; }
    // fused exchange of pai (:673): the edge columns store the levels they have just written (own stores: no
    // fence needed) into the neighbours' ghost cells, after the sweeps so that the sweeps carry no trace of it
    if (PUSH && valid) {
      if (pc.mask & 3) {          // left/right neighbours: the general per-cell push
        for (int k = 1; k <= kz; ++k) edge_push(pc, ep, j, i, k, pai[base + k * pl]);
      } else {
        const ColPush cp = col_push_init(pc, ep, j, i, true);
        if (cp.t2 || cp.t3)
          for (int k = 1; k <= kz; ++k) col_push(pc, cp, k, pai[base + k * pl]);
      }
    }
  }
  tm.release();
  if (PUSH) halo_producer_done(pc, blockIdx.x, gridDim.x, 4, ntile_j, g.ici2 - g.ici1 + 1, 1);
}
template <int D, bool PUSH>
static int launch_wsolve_tm_t(Ctx& c, double dts, bool last, const PushCtl& pc, const EdgePush& ep) {
  const Geo& g = c.g;
  const int ntile_j = (g.jci2 - g.jde1 + 1 + 31) / 32, ni = g.ici2 - g.ici1 + 1, ntiles = ntile_j * ni;
  const double dtrdz = dts * c.rdzita;
  const double zcs2 = (dtrdz * dtrdz) * rdrcv;
  // at least 78 KB: no more than two CTAs per SM, i.e. no more CTAs than tensor-memory allocations of 256 columns
  const size_t smem = std::max((size_t)(4 * D * 9) * 32 * sizeof(double) + 16, (size_t)78 * 1024);
  const double* zsrc = c.cfg.mo_divfilter ? c.zdiv2b : c.f[MB_ZDIV2].p;
  MB_CUDA(cudaFuncSetAttribute(moloch_wsolve_tm<D, PUSH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  LaunchScope ls(c, KID_WSOLVE);
  moloch_wsolve_tm<D, PUSH><<<(unsigned)((ntiles + 3) / 4), 128, smem, c.stream>>>(
      g, zsrc, c.f[MB_S].p, c.f[MB_W].p, c.f[MB_PAI].p, c.f[MB_TETAV].p, c.f[MB_TETAVF].p, c.f[MB_FMZ].p,
      c.f[MB_FMZF].p, c.f[MB_BDYWTW].p, c.prof[MB_FFILT], dts, dtrdz, zcs2, last ? 1 : 0, ntile_j, ntiles, pc, ep);
  MB_CUDA(cudaGetLastError());
  return 0;
}
template <int D>
static int launch_wsolve_tm(Ctx& c, double dts, bool last, const PushCtl& pc, const EdgePush& ep) {
  return pc.mask ? launch_wsolve_tm_t<D, true>(c, dts, last, pc, ep) : launch_wsolve_tm_t<D, false>(c, dts, last, pc, ep);
}
// variant 11: ring of 8 (74 KB of shared memory per CTA of four warps); variant 12: ring of 6 (55 KB)
int k_wsolve_tm(Ctx& c, double dts, bool last, const PushCtl& pc, const EdgePush& ep) {
  if (c.wsolve_impl == 13) return launch_wsolve_tm<4>(c, dts, last, pc, ep);
  return c.wsolve_impl == 12 ? launch_wsolve_tm<6>(c, dts, last, pc, ep) : launch_wsolve_tm<8>(c, dts, last, pc, ep);
}
// variant 8: three sweep arrays + ring of 6 (46 KB per warp at kz = 41, 4 warps per SM); variant 9: two sweep
// arrays (divergence recomputed) + ring of 9 (42 KB, 5 warps per SM); variant 10: two + ring of 12 (49 KB, 4 warps)
int k_wsolve8(Ctx& c, double dts, bool last, const PushCtl& pc, const EdgePush& ep) {
  if (c.wsolve_impl == 9) return launch_wsolve8<9, false>(c, dts, last, pc, ep);
  if (c.wsolve_impl == 10) return launch_wsolve8<12, false>(c, dts, last, pc, ep);
  return launch_wsolve8<6, true>(c, dts, last, pc, ep);
}
// variant 6: ring of 4 levels (31 KB per warp at kz = 41: 7 warps per SM); variant 7: ring of 6 (36 KB: 6 warps)
int k_wsolve6(Ctx& c, double dts, bool last, const PushCtl& pc, const EdgePush& ep) {
  return c.wsolve_impl == 7 ? launch_wsolve6<6>(c, dts, last, pc, ep) : launch_wsolve6<4>(c, dts, last, pc, ep);
}
int k_wsolve5(Ctx& c, double dts, bool last, const PushCtl& pc, const EdgePush& ep) {
  // One warp per CTA and a latency-bound column sweep: what matters on small
  // per-GPU grids is the number of CTA waves.  A shallower ring needs less shared
  // memory (5 instead of 4 CTAs per SM at kz = 41): take it when it saves a wave.
  const Geo& g = c.g;
  const long long nblk = ((long long)(g.jci2 - g.jci1 + 1) * (g.ici2 - g.ici1 + 1) + 31) / 32;
  auto waves = [&](int d) {
    const long long smem = (long long)(3 * (g.kz + 1) + d * 9) * 32 * 8 + 1024;
    long long per_sm = (227 * 1024) / smem;
    if (per_sm > 32) per_sm = 32;
    if (per_sm < 1) per_sm = 1;
    return (nblk + 148 * per_sm - 1) / (148 * per_sm);
  };
  if (waves(4) == 1 && waves(6) == 2) return launch_wsolve5<4>(c, dts, last, pc, ep);
  return launch_wsolve5<6>(c, dts, last, pc, ep);
}

// ---------------------------------------------------------------------------
// K11  s = (w+s)*fmzf, s(1)=s(kzp1)=0                                 :728-734
// ---------------------------------------------------------------------------
__global__ void moloch_sfinish(Geo g, double* __restrict__ s, const double* __restrict__ w,
                               const double* __restrict__ fmzf) {
  THREAD_JIK(g.jci1, g.ici1, 1)
  if (j > g.jci2 || i > g.ici2) return;
  const long long id = IX(j, i, k);
  if (k == 1 || k == g.kz + 1) s[id] = 0.0;
  else s[id] = (w[id] + s[id]) * fmzf[id];
}
int k_sfinish(Ctx& c) {
  const Geo& g = c.g;
  LaunchScope ls(c, KID_SFINISH);
  moloch_sfinish<<<grid3(g.jci2 - g.jci1 + 1, g.ici2 - g.ici1 + 1, g.kz + 1), dim3(BX, BY), 0, c.stream>>>(
      g, c.f[MB_S].p, c.f[MB_W].p, c.f[MB_FMZF].p);
  MB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// K12+K13  uvstagtouvx :1537-1567 and zstagtoh :1451-1458
// ---------------------------------------------------------------------------
// FUSED (peer-store transport inside dynamical_core): the kernel first waits for the neighbours' u, v edges
// (pushed, 2 wide, by the last sub-step's uvupdate) and stores its own ux / vx edges, 2 wide, into the
// neighbours' ghost cells: the physical-boundary rows and columns among them are final here, the others are
// overwritten by curvature's push after the advection (uvxtouvstag's exchange, :1485-1486).
template <bool FUSED>
__global__ void moloch_destagger(Geo g, const double* __restrict__ u, const double* __restrict__ v,
                                 const double* __restrict__ w, double* __restrict__ ux,
                                 double* __restrict__ vx, double* __restrict__ wx, WaitCtl wc, PushCtl pc,
                                 EdgePush eux, EdgePush evx) {
  if (FUSED) halo_sync(wc, 2, g.jce2 - g.jce1 + 1, g.ice2 - g.ice1 + 1, BX, BY);
  THREAD_JIK(g.jce1, g.ice1, 1)
  if (j > g.jce2 || i > g.ice2) return;
  const long long id = IX(j, i, k);
  const int kz = g.kz;
  double uxn, vxn;
  // ux on jci1:jci2 (4th order) and the physical-boundary columns (2nd order)
  if (j >= g.jci1 && j <= g.jci2) {
    uxn = 0.5625 * (u[id + 1] + u[id]) - 0.0625 * (u[id + 2] + u[id - 1]);
  } else {
    // j == jce1 with has_bdyleft: u(jde1),u(jdi1) = u(j),u(j+1); j == jce2 with
    // has_bdyright: u(jde2),u(jdi2) = u(j+1),u(j)
    if (j == g.jce1) uxn = 0.5 * (u[id] + u[id + 1]);
    else uxn = 0.5 * (u[id + 1] + u[id]);
  }
  ux[id] = uxn;
  if (i >= g.ici1 && i <= g.ici2) {
    vxn = 0.5625 * (v[id + g.NJ] + v[id]) - 0.0625 * (v[id + 2 * g.NJ] + v[id - g.NJ]);
  } else {
    if (i == g.ice1) vxn = 0.5 * (v[id] + v[id + g.NJ]);
    else vxn = 0.5 * (v[id + g.NJ] + v[id]);
  }
  vx[id] = vxn;
  if (FUSED && pc.mask) { edge_push(pc, eux, j, i, k, uxn); edge_push(pc, evx, j, i, k, vxn); }
  const long long pl = g.plane;
  if (k == 1) wx[id] = 0.5 * (w[id + pl] + w[id]);
  else if (k == kz) wx[id] = 0.5 * (w[id + pl] + w[id]);
  else wx[id] = 0.5625 * (w[id + pl] + w[id]) - 0.0625 * (w[id + 2 * pl] + w[id - pl]);
}
int k_destagger(Ctx& c, const WaitCtl* wc, const PushCtl* pc, const EdgePush* eux, const EdgePush* evx) {
  const Geo& g = c.g;
  const WaitCtl w0 = wc ? *wc : WaitCtl{};
  const PushCtl p0 = pc ? *pc : PushCtl{};
  const EdgePush e0 = eux ? *eux : EdgePush{}, e1 = evx ? *evx : EdgePush{};
  LaunchScope ls(c, KID_DESTAG);
  const dim3 grid = grid3(g.jce2 - g.jce1 + 1, g.ice2 - g.ice1 + 1, g.kz);
  if (w0.mask || p0.mask)
    moloch_destagger<true><<<grid, dim3(BX, BY), 0, c.stream>>>(
        g, c.f[MB_U].p, c.f[MB_V].p, c.f[MB_W].p, c.f[MB_UX].p, c.f[MB_VX].p, c.f[MB_WX].p, w0, p0, e0, e1);
  else
    moloch_destagger<false><<<grid, dim3(BX, BY), 0, c.stream>>>(
        g, c.f[MB_U].p, c.f[MB_V].p, c.f[MB_W].p, c.f[MB_UX].p, c.f[MB_VX].p, c.f[MB_WX].p, w0, p0, e0, e1);
  MB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// K14  wafone vertical advection, twice with dt/2                     :863-922
// One thread per column; the column is staged in shared memory (thread-private
// slots), both half-steps run on it and only wz goes back to HBM.
// ---------------------------------------------------------------------------
constexpr int WZ_TB = 64;
__device__ __forceinline__ void waf_vertical_pass(const Geo& g, const double* q, double* out,
                                                  const double* __restrict__ s,
                                                  const double* __restrict__ fmz,
                                                  const double* __restrict__ fmzf, long long base,
                                                  double dtrdz) {
  // q, out: shared-memory columns with stride WZ_TB, level k at [(k-1)*WZ_TB]
  const int kz = g.kz;
  const long long pl = g.plane;
  double fprev = 0.0;  // wfw(1) = 0 :864
  double sk = s[base];
  double fmzfk = fmzf[base];
  for (int k = 1; k <= kz; ++k) {
    const long long id = base + (k - 1) * pl;
    const double sk1 = s[id + pl];
    const double fmzfk1 = fmzf[id + pl];
    const double qk = q[(k - 1) * WZ_TB];
    double fnext = 0.0;  // wfw(kzp1) = 0 :865
    if (k < kz) {
      const double zamu = sk1 * dtrdz;
      double is; int k1, k1p1;
      if (zamu >= 0.0) { is = 1.0; k1 = k + 1; k1p1 = k1 + 1; if (k1p1 > kz) k1p1 = kz; }
      else { is = -1.0; k1 = k - 1; k1p1 = k; if (k1 < 1) k1 = 1; }
      const double qk1 = q[k * WZ_TB];
      const double rr = flow_param(q[(k1 - 1) * WZ_TB] - q[(k1p1 - 1) * WZ_TB], qk - qk1);
      const double zphi = waf_phi(rr, zamu, is);
      fnext = 0.5 * sk1 * ((1.0 + zphi) * qk1 + (1.0 - zphi) * qk);
    }
    const double fm = fmz[id];
    const double zrfmu = dtrdz * fm / fmzfk;
    const double zrfmd = dtrdz * fm / fmzfk1;
    const double zdv = (sk * zrfmu - sk1 * zrfmd) * qk;
    out[(k - 1) * WZ_TB] = qk - fprev * zrfmu + fnext * zrfmd + zdv;
    fprev = fnext; sk = sk1; fmzfk = fmzfk1;
  }
}
__global__ void __launch_bounds__(WZ_TB)
moloch_waf_vertical(Geo g, double* const* __restrict__ tab, int first, double* __restrict__ wzall,
                    const double* __restrict__ s, const double* __restrict__ fmz,
                    const double* __restrict__ fmzf, double dtrdz) {
  extern __shared__ double sm[];
  const int nj = g.jce2 - g.jce1 + 1, ni = g.ice2 - g.ice1 + 1;
  const long long col = (long long)blockIdx.x * WZ_TB + threadIdx.x;
  if (col >= (long long)nj * ni) return;
  const int i = g.ice1 + (int)(col / nj), j = g.jce1 + (int)(col % nj);
  const int kz = g.kz;
  const double* __restrict__ pp = tab[first + blockIdx.y];
  double* __restrict__ wz = wzall + (long long)blockIdx.y * kz * g.plane;
  double* a = sm + threadIdx.x;
  double* b = sm + (size_t)kz * WZ_TB + threadIdx.x;
  const long long base = IX(j, i, 1);
  for (int k = 1; k <= kz; ++k) a[(k - 1) * WZ_TB] = pp[base + (k - 1) * g.plane];
  waf_vertical_pass(g, a, b, s, fmz, fmzf, base, dtrdz);
  waf_vertical_pass(g, b, a, s, fmz, fmzf, base, dtrdz);  // do_vadvtwice :894
  for (int k = 1; k <= kz; ++k) wz[base + (k - 1) * g.plane] = a[(k - 1) * WZ_TB];
}
int k_waf_z(Ctx& c, int first, int count, double dta) {
  const Geo& g = c.g;
  const double dtrdz = 0.5 * (dta * c.rdzita);  // :857-860
  const long long ncol = (long long)(g.jce2 - g.jce1 + 1) * (g.ice2 - g.ice1 + 1);
  const size_t smem = (size_t)2 * g.kz * WZ_TB * sizeof(double);
  MB_CUDA(cudaFuncSetAttribute(moloch_waf_vertical, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  LaunchScope ls(c, KID_WAF_Z);
  moloch_waf_vertical<<<dim3((unsigned)((ncol + WZ_TB - 1) / WZ_TB), (unsigned)count), WZ_TB, smem, c.stream>>>(
      g, c.d_ptrtab, first, c.wzall, c.f[MB_S].p, c.f[MB_FMZ].p, c.f[MB_FMZF].p, dtrdz);
  MB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// K15  wafone meridional flux + update -> p0                :929-953 / :987-1010
// The flux at a V face is recomputed by the two cells that share it (no zpby
// array in HBM); both evaluations are bit-identical.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double waf_flux_y(const Geo& g, const double* __restrict__ wz,
                                             const double* __restrict__ v, const double* __restrict__ mv,
                                             int j, int i, int k, double dtrdy) {
  const long long id = IX(j, i, k);
  const double vv = v[id];
  const double zamu = g.lrotllr ? vv * dtrdy : vv * mv[IX2(j, i)] * dtrdy;
  double is; int ih;
  if (zamu > 0.0) { is = 1.0; ih = i - 1; } else { is = -1.0; ih = min(i + 1, g.imax); }
  const int ihm1 = max(ih - 1, g.imin);
  const double w0 = wz[id], wm = wz[id - g.NJ];
  const double rr = flow_param(wz[IX(j, ih, k)] - wz[IX(j, ihm1, k)], w0 - wm);
  const double zphi = waf_phi(rr, zamu, is);
  return 0.5 * vv * ((1.0 + zphi) * wm + (1.0 - zphi) * w0);
}
__global__ void moloch_waf_meridional(Geo g, double* const* __restrict__ tab, int first,
                                      const double* __restrict__ wzall, double* __restrict__ p0all,
                                      const double* __restrict__ v, const double* __restrict__ fmz,
                                      const double* __restrict__ rfmzu, const double* __restrict__ rfmzv,
                                      const double* __restrict__ mx, const double* __restrict__ mx2,
                                      const double* __restrict__ mv, const double* __restrict__ rmv,
                                      double dtrdy) {
  const int kz = g.kz;
  const int j = g.jce1 + blockIdx.x * BX + threadIdx.x;
  const int i = g.ici1 + blockIdx.y * BY + threadIdx.y;
  const int fld = blockIdx.z / kz, k = 1 + blockIdx.z % kz;
  if (j > g.jce2 || i > g.ici2) return;
  const double* __restrict__ pp = tab[first + fld];
  const double* __restrict__ wz = wzall + (long long)fld * kz * g.plane;
  double* __restrict__ p0 = p0all + (long long)fld * kz * g.plane;
  const long long id = IX(j, i, k);
  const long long i2 = IX2(j, i);
  const double fs = waf_flux_y(g, wz, v, mv, j, i, k, dtrdy);
  const double fn = waf_flux_y(g, wz, v, mv, j, i + 1, k, dtrdy);
  if (g.lrotllr) {
    const double zhxvtn = dtrdy * rmv[i2 + g.NJ] * mx[i2];
    const double zhxvts = dtrdy * rmv[i2] * mx[i2];
    const double zrfmn = zhxvtn * fmz[id] * rfmzv[id + g.NJ];
    const double zrfms = zhxvts * fmz[id] * rfmzv[id];
    const double zdv = (v[id + g.NJ] * zrfmn - v[id] * zrfms) * pp[id];
    p0[id] = wz[id] + (fs * zrfms - fn * zrfmn + zdv);
  } else {
    const double zrfmn = dtrdy * fmz[id] * rfmzu[id + g.NJ];  // sic: rfmzu :1004-1005
    const double zrfms = dtrdy * fmz[id] * rfmzu[id];
    const double zdv = (v[id + g.NJ] * rmv[i2 + g.NJ] * zrfmn - v[id] * rmv[i2] * zrfms) * pp[id];
    p0[id] = wz[id] + mx2[i2] * (fs * zrfms - fn * zrfmn + zdv);
  }
}
int k_waf_y(Ctx& c, int first, int count, double dta) {
  const Geo& g = c.g;
  LaunchScope ls(c, KID_WAF_Y);
  moloch_waf_meridional<<<grid3(g.jce2 - g.jce1 + 1, g.ici2 - g.ici1 + 1, g.kz * count), dim3(BX, BY), 0,
                          c.stream>>>(g, c.d_ptrtab, first, c.wzall, c.p0all, c.f[MB_V].p, c.f[MB_FMZ].p,
                                      c.f[MB_RFMZU].p, c.f[MB_RFMZV].p, c.f[MB_MSFX].p, c.mx2,
                                      c.f[MB_MSFV].p, c.rmv, dta * c.rdx);
  MB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// K16  wafone zonal flux + update -> pp                    :959-982 / :1015-1038
// ---------------------------------------------------------------------------
__device__ __forceinline__ double waf_flux_x(const Geo& g, const double* __restrict__ p0,
                                             const double* __restrict__ u, const double* __restrict__ mu,
                                             int j, int i, int k, double dtrdx) {
  const long long id = IX(j, i, k);
  const double uu = u[id];
  const double zamu = uu * mu[IX2(j, i)] * dtrdx;
  double is; int jh;
  if (zamu > 0.0) { is = 1.0; jh = j - 1; } else { is = -1.0; jh = min(j + 1, g.jmax); }
  const int jhm1 = max(jh - 1, g.jmin);
  const double q0 = p0[id], qm = p0[id - 1];
  const double rr = flow_param(p0[IX(jh, i, k)] - p0[IX(jhm1, i, k)], q0 - qm);
  const double zphi = waf_phi(rr, zamu, is);
  return 0.5 * uu * ((1.0 + zphi) * qm + (1.0 - zphi) * q0);
}
__global__ void moloch_waf_zonal(Geo g, double* const* __restrict__ tab, int first,
                                 const double* __restrict__ p0all, const double* __restrict__ u,
                                 const double* __restrict__ fmz, const double* __restrict__ rfmzu,
                                 const double* __restrict__ mx, const double* __restrict__ mx2,
                                 const double* __restrict__ mu, const double* __restrict__ rmu,
                                 double dtrdx) {
  const int kz = g.kz;
  const int j = g.jci1 + blockIdx.x * BX + threadIdx.x;
  const int i = g.ici1 + blockIdx.y * BY + threadIdx.y;
  const int fld = blockIdx.z / kz, k = 1 + blockIdx.z % kz;
  if (j > g.jci2 || i > g.ici2) return;
  double* __restrict__ pp = tab[first + fld];
  const double* __restrict__ p0 = p0all + (long long)fld * kz * g.plane;
  const long long id = IX(j, i, k);
  const long long i2 = IX2(j, i);
  const double fw = waf_flux_x(g, p0, u, mu, j, i, k, dtrdx);
  const double fe = waf_flux_x(g, p0, u, mu, j + 1, i, k, dtrdx);
  if (g.lrotllr) {
    const double zcostx = dtrdx * mx[i2];
    const double zrfmw = zcostx * fmz[id] * rfmzu[id];
    const double zrfme = zcostx * fmz[id] * rfmzu[id + 1];
    const double zdv = (u[id + 1] * zrfme - u[id] * zrfmw) * pp[id];
    pp[id] = p0[id] + fw * zrfmw - fe * zrfme + zdv;
  } else {
    const double zrfmw = dtrdx * fmz[id] * rfmzu[id];
    const double zrfme = dtrdx * fmz[id] * rfmzu[id + 1];
    const double zdv = (u[id + 1] * rmu[i2 + 1] * zrfme - u[id] * rmu[i2] * zrfmw) * pp[id];
    pp[id] = p0[id] + mx2[i2] * (fw * zrfmw - fe * zrfme + zdv);
  }
}
int k_waf_x(Ctx& c, int first, int count, double dta) {
  const Geo& g = c.g;
  LaunchScope ls(c, KID_WAF_X);
  moloch_waf_zonal<<<grid3(g.jci2 - g.jci1 + 1, g.ici2 - g.ici1 + 1, g.kz * count), dim3(BX, BY), 0, c.stream>>>(
      g, c.d_ptrtab, first, c.p0all, c.f[MB_U].p, c.f[MB_FMZ].p, c.f[MB_RFMZU].p, c.f[MB_MSFX].p, c.mx2,
      c.f[MB_MSFU].p, c.rmu, dta * c.rdx);
  MB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// K17  curvature terms                                                :811-825
// ---------------------------------------------------------------------------
template <bool FUSED>
__global__ void moloch_curvature(Geo g, double* __restrict__ ux, double* __restrict__ vx,
                                 const double* __restrict__ mx, const double* __restrict__ mu,
                                 const double* __restrict__ mv, const double* __restrict__ rlat, double rdx,
                                 double dta, PushCtl pc, EdgePush eux, EdgePush evx) {
  THREAD_JIK(g.jci1, g.ici1, 1)
  if (j <= g.jci2 && i <= g.ici2) {
  const long long id = IX(j, i, k);
  const long long i2 = IX2(j, i);
  double tanx, tany;
  if (g.lrotllr) {
    // the MB_RLAT slot holds sin(degrad*0.5*(rlat(i)+rlat(i+1))), 1-based from
    // i = ide1, evaluated once on the host at set_profile time (:813-814)
    tanx = rlat[i - g.ide1 + 1] * mx[i2] * rearthrad;
    tany = tanx;
  } else {
    tanx = (mu[i2 - 1] - mu[i2]) * rdx;
    tany = (mv[i2 - g.NJ] - mv[i2]) * rdx;
  }
  const double uxn = ux[id] + ux[id] * vx[id] * tanx * dta;
  ux[id] = uxn;
  const double vxn = vx[id] - uxn * uxn * tany * dta;
  vx[id] = vxn;
  if (FUSED && pc.mask) { edge_push(pc, eux, j, i, k, uxn); edge_push(pc, evx, j, i, k, vxn); }
  }
  if (FUSED) halo_producer_done(pc, blockIdx.y, gridDim.y, BY, 1, g.ici2 - g.ici1 + 1, gridDim.x * gridDim.z);
}
int k_curvature(Ctx& c, double dta, const PushCtl* pc, const EdgePush* eux, const EdgePush* evx) {
  const Geo& g = c.g;
  const PushCtl p0 = pc ? *pc : PushCtl{};
  const EdgePush e0 = eux ? *eux : EdgePush{}, e1 = evx ? *evx : EdgePush{};
  LaunchScope ls(c, KID_CURV);
  const dim3 grid = grid3(g.jci2 - g.jci1 + 1, g.ici2 - g.ici1 + 1, g.kz);
  if (p0.mask)
    moloch_curvature<true><<<grid, dim3(BX, BY), 0, c.stream>>>(
        g, c.f[MB_UX].p, c.f[MB_VX].p, c.f[MB_MSFX].p, c.f[MB_MSFU].p, c.f[MB_MSFV].p, c.prof[MB_RLAT], c.rdx,
        dta, p0, e0, e1);
  else
    moloch_curvature<false><<<grid, dim3(BX, BY), 0, c.stream>>>(
        g, c.f[MB_UX].p, c.f[MB_VX].p, c.f[MB_MSFX].p, c.f[MB_MSFU].p, c.f[MB_MSFV].p, c.prof[MB_RLAT], c.rdx,
        dta, p0, e0, e1);
  MB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// K18+K19  uvxtouvstag :1490-1520 and htozstag :1467-1474
// ---------------------------------------------------------------------------
template <bool FUSED>
__global__ void moloch_restagger(Geo g, const double* __restrict__ ux, const double* __restrict__ vx,
                                 const double* __restrict__ wx, double* __restrict__ u,
                                 double* __restrict__ v, double* __restrict__ w, int with_w, WaitCtl wc) {
  if (FUSED) halo_sync(wc, 2, g.jde2 - g.jde1 + 1, g.ide2 - g.ide1 + 1, BX, BY);   // ux, vx ghosts (2 wide)
  THREAD_JIK(g.jde1, g.ide1, 1)
  if (j > g.jde2 || i > g.ide2) return;
  const long long id = IX(j, i, k);
  const int kz = g.kz;
  if (i >= g.ici1 && i <= g.ici2) {
    if (j >= g.jdii1 && j <= g.jdii2) {
      u[id] = 0.5625 * (ux[id] + ux[id - 1]) - 0.0625 * (ux[id + 1] + ux[id - 2]);
    } else if (g.br && j == g.jdi2) {
      u[id] = 0.5 * (ux[id - 1] + ux[id]);  // ux(jci2), ux(jce2) = ux(j-1), ux(j)
    } else if (g.bl && j == g.jdi1) {
      u[id] = 0.5 * (ux[id] + ux[id - 1]);  // ux(jci1), ux(jce1) = ux(j), ux(j-1)
    }
  }
  if (j >= g.jci1 && j <= g.jci2) {
    if (i >= g.idii1 && i <= g.idii2) {
      v[id] = 0.5625 * (vx[id] + vx[id - g.NJ]) - 0.0625 * (vx[id + g.NJ] + vx[id - 2 * g.NJ]);
    } else if (g.bt && i == g.idi2) {
      v[id] = 0.5 * (vx[id - g.NJ] + vx[id]);
    } else if (g.bb && i == g.idi1) {
      v[id] = 0.5 * (vx[id] + vx[id - g.NJ]);
    }
  }
  if (with_w && j <= g.jce2 && i <= g.ice2 && k >= 2) {
    const long long pl = g.plane;
    if (k == 2) w[id] = 0.5 * (wx[id] + wx[id - pl]);
    else if (k == kz) w[id] = 0.5 * (wx[id] + wx[id - pl]);
    else w[id] = 0.5625 * (wx[id] + wx[id - pl]) - 0.0625 * (wx[id + pl] + wx[id - 2 * pl]);
  }
}
int k_restagger(Ctx& c, bool with_w, const WaitCtl* wc) {
  const Geo& g = c.g;
  const WaitCtl w0 = wc ? *wc : WaitCtl{};
  LaunchScope ls(c, KID_RESTAG);
  const dim3 grid = grid3(g.jde2 - g.jde1 + 1, g.ide2 - g.ide1 + 1, g.kz);
  if (w0.mask)
    moloch_restagger<true><<<grid, dim3(BX, BY), 0, c.stream>>>(
        g, c.f[MB_UX].p, c.f[MB_VX].p, c.f[MB_WX].p, c.f[MB_U].p, c.f[MB_V].p, c.f[MB_W].p, with_w ? 1 : 0, w0);
  else
    moloch_restagger<false><<<grid, dim3(BX, BY), 0, c.stream>>>(
        g, c.f[MB_UX].p, c.f[MB_VX].p, c.f[MB_WX].p, c.f[MB_U].p, c.f[MB_V].p, c.f[MB_W].p, with_w ? 1 : 0, w0);
  MB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// K20  tvirt = tetav*pai, t = tvirt/moist                  :1121-1125,:1629-1648
// ---------------------------------------------------------------------------
__global__ void moloch_tvirt_temp(Geo g, const double* __restrict__ tetav, const double* __restrict__ pai,
                                  const double* __restrict__ qx, double* __restrict__ tvirt,
                                  double* __restrict__ t) {
  THREAD_JIK(g.jce1, g.ice1, 1)
  if (j > g.jce2 || i > g.ice2) return;
  const long long id = IX(j, i, k);
  const double tv = tetav[id] * pai[id];
  tvirt[id] = tv;
  if (in_box(j, i, g.jci1, g.jci2, g.ici1, g.ici2)) t[id] = tv / moist_factor(g, qx, id);
}
int k_tvirt_temp(Ctx& c) {
  const Geo& g = c.g;
  LaunchScope ls(c, KID_TVIRT);
  moloch_tvirt_temp<<<grid3(g.jce2 - g.jce1 + 1, g.ice2 - g.ice1 + 1, g.kz), dim3(BX, BY), 0, c.stream>>>(
      g, c.f[MB_TETAV].p, c.f[MB_PAI].p, c.f[MB_QX].p, c.f[MB_TVIRT].p, c.f[MB_T].p);
  MB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// K21  p, rho, qsat :348-352 and extrapolate_surface_pressure :1592-1606
// ---------------------------------------------------------------------------
__global__ void moloch_diag_prq(Geo g, const double* __restrict__ pai, const double* __restrict__ t,
                                double* __restrict__ p, double* __restrict__ rho, double* __restrict__ qsat) {
  THREAD_JIK(g.jce1, g.ice1, 1)
  if (j > g.jce2 || i > g.ice2) return;
  const long long id = IX(j, i, k);
  const double pp = pow(pai[id], cpovr) * p00;
  const double tt = t[id];
  p[id] = pp;
  rho[id] = pp / (rgas * tt);
  qsat[id] = pfwsat(tt, pp);
}
__global__ void moloch_diag_ps(Geo g, const double* __restrict__ tvirt, const double* __restrict__ z,
                               const double* __restrict__ p, double* __restrict__ ps) {
  const int j = g.jci1 + blockIdx.x * BX + threadIdx.x;
  const int i = g.ici1 + blockIdx.y * BY + threadIdx.y;
  if (j > g.jci2 || i > g.ici2) return;
  const int kz = g.kz;
  const long long a = IX(j, i, kz), b = IX(j, i, kz - 1);
  double lrt = (tvirt[b] - tvirt[a]) / (z[b] - z[a]);
  if (lrt < -govcp) lrt = -govcp;
  else if (lrt > -0.005) lrt = 0.65 * lrt - 0.35 * lrate;
  const double tv = tvirt[a] - lrt * 0.5 * z[a];
  ps[IX2(j, i)] = p[a] * exp(govr * z[a] / tv);
}
int k_diagnostics(Ctx& c) {
  const Geo& g = c.g;
  {
    LaunchScope ls(c, KID_DIAG);
    moloch_diag_prq<<<grid3(g.jce2 - g.jce1 + 1, g.ice2 - g.ice1 + 1, g.kz), dim3(BX, BY), 0, c.stream>>>(
        g, c.f[MB_PAI].p, c.f[MB_T].p, c.f[MB_P].p, c.f[MB_RHO].p, c.f[MB_QSAT].p);
    MB_CUDA(cudaGetLastError());
  }
  {
    LaunchScope ls(c, KID_PS);
    moloch_diag_ps<<<grid3(g.jci2 - g.jci1 + 1, g.ici2 - g.ici1 + 1, 1), dim3(BX, BY), 0, c.stream>>>(
        g, c.f[MB_TVIRT].p, c.f[MB_ZETA].p, c.f[MB_P].p, c.f[MB_PS].p);
    MB_CUDA(cudaGetLastError());
  }
  return 0;
}

// ---------------------------------------------------------------------------
// K22  status_update (without its trailing uvxtouvstag)             :1410-1438
// ---------------------------------------------------------------------------
// FUSED (peer-store transport, fusion level 2): the two stand-alone rounds that follow the kernel in the reference's
// order -- the synchronisation-only round that keeps a neighbour's uvxtouvstag (end of advection) from reading
// ux, vx ghosts that are already being overwritten, and exchange of ux, vx (:1485-1486 as called from :1431) --
// are folded in: the edge CTAs wait for the neighbours' word "everything before my status_update has completed"
// (wf), every owned edge cell of ux / vx (updated or not: the boundary update may have changed the others) is
// stored into the neighbours' ghost cells, and uvxtouvstag waits for the neighbours' status_update (wx).
#ifndef MB_SU_G
#define MB_SU_G 5
#endif
constexpr int SU_G = MB_SU_G;   // species per group of moloch_status_update
template <bool FUSED>
__global__ void moloch_status_update(Geo g, double* __restrict__ t, double* __restrict__ ux,
                                     double* __restrict__ vx, double* __restrict__ qx,
                                     double* __restrict__ trac, const double* __restrict__ tten,
                                     const double* __restrict__ uten, const double* __restrict__ vten,
                                     const double* __restrict__ qxten, const double* __restrict__ chiten,
                                     const double* __restrict__ pai, const double* __restrict__ p,
                                     double* __restrict__ tvirt, double* __restrict__ tetav,
                                     double* __restrict__ rho, double* __restrict__ qsat, double dtinc, WaitCtl wf,
                                     PushCtl pc, EdgePush eux, EdgePush evx) {
  if (FUSED) halo_sync(wf, 2, g.jce2 - g.jce1 + 1, g.ice2 - g.ice1 + 1, BX, BY);
  THREAD_JIK(g.jce1, g.ice1, 1)
  if (j > g.jce2 || i > g.ice2) return;
  const long long id = IX(j, i, k);
  const long long sp = (long long)g.kz * g.plane;
  double tt = t[id];
  const bool inner = in_box(j, i, g.jci1, g.jci2, g.ici1, g.ici2);
  if (FUSED && pc.mask) {
    double uxn = ux[id], vxn = vx[id];
    if (inner) { uxn = uxn + dtinc * uten[id]; vxn = vxn + dtinc * vten[id]; }
    edge_push(pc, eux, j, i, k, uxn);
    edge_push(pc, evx, j, i, k, vxn);
  }
  if (inner) {
    tt = tt + dtinc * tten[id];
    t[id] = tt;
    ux[id] = ux[id] + dtinc * uten[id];
    vx[id] = vx[id] + dtinc * vten[id];
    // species in groups of SU_G with every load of a group issued before its first store (the compiler cannot move a
    // load of qx / trac across a store to the same array: one species at a time leaves one load pair in flight)
#pragma unroll 1
    for (int n0 = 0; n0 < g.nqx; n0 += SU_G) {
      double q[SU_G], dq[SU_G];
#pragma unroll
      for (int m = 0; m < SU_G; ++m) {
        const bool on = n0 + m < g.nqx;
        q[m] = on ? qx[id + (n0 + m) * sp] : 0.0;
        dq[m] = on ? qxten[id + (n0 + m) * sp] : 0.0;
      }
#pragma unroll
      for (int m = 0; m < SU_G; ++m)
        if (n0 + m < g.nqx) {
          double v = q[m] + dtinc * dq[m];
          if (v < c_qxcheckval[n0 + m]) v = c_qxzeroval[n0 + m];
          qx[id + (n0 + m) * sp] = v;
        }
    }
#pragma unroll 1
    for (int n0 = 0; n0 < g.ntr; n0 += SU_G) {
      double q[SU_G], dq[SU_G];
#pragma unroll
      for (int m = 0; m < SU_G; ++m) {
        const bool on = n0 + m < g.ntr;
        q[m] = on ? trac[id + (n0 + m) * sp] : 0.0;
        dq[m] = on ? chiten[id + (n0 + m) * sp] : 0.0;
      }
#pragma unroll
      for (int m = 0; m < SU_G; ++m)
        if (n0 + m < g.ntr) {
          double v = q[m] + dtinc * dq[m];
          if (v < 0.0) v = 0.0;
          trac[id + (n0 + m) * sp] = v;
        }
    }
  }
  const double tv = tt * moist_factor(g, qx, id);
  tvirt[id] = tv;
  tetav[id] = tv / pai[id];
  const double pp = p[id];
  rho[id] = pp / (rgas * tt);
  qsat[id] = pfwsat(tt, pp);
}
int k_status_update(Ctx& c, double dtinc, const WaitCtl* wf, const PushCtl* pc, const EdgePush* eux, const EdgePush* evx) {
  const Geo& g = c.g;
  const WaitCtl w0 = wf ? *wf : WaitCtl{};
  const PushCtl p0 = pc ? *pc : PushCtl{};
  const EdgePush e0 = eux ? *eux : EdgePush{}, e1 = evx ? *evx : EdgePush{};
  LaunchScope ls(c, KID_STATUS);
  const dim3 grid = grid3(g.jce2 - g.jce1 + 1, g.ice2 - g.ice1 + 1, g.kz);
#define SU_ARGS g, c.f[MB_T].p, c.f[MB_UX].p, c.f[MB_VX].p, c.f[MB_QX].p, c.f[MB_TRAC].p, c.f[MB_TTEN].p, c.f[MB_UTEN].p, \
      c.f[MB_VTEN].p, c.f[MB_QXTEN].p, c.f[MB_CHITEN].p, c.f[MB_PAI].p, c.f[MB_P].p, c.f[MB_TVIRT].p, c.f[MB_TETAV].p, \
      c.f[MB_RHO].p, c.f[MB_QSAT].p, dtinc, w0, p0, e0, e1
  if (w0.mask || p0.mask) moloch_status_update<true><<<grid, dim3(BX, BY), 0, c.stream>>>(SU_ARGS);
  else moloch_status_update<false><<<grid, dim3(BX, BY), 0, c.stream>>>(SU_ARGS);
#undef SU_ARGS
  MB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// K0  reset_tendencies                                              :1044-1083
// Whole-array clears: cells outside the reference's loop ranges (ghosts, pads)
// are never read.
// ---------------------------------------------------------------------------
int k_reset_tendencies(Ctx& c) {
  // tten, uten, vten, qxten, chiten, s, zdiv2 are consecutive arena slots: one
  // clear.  (wwkw, also cleared by the reference, only exists inside moloch_wsolve.)
  const Layout& L = c.layout;
  const size_t lo = L.off[MB_TTEN], hi = L.off[MB_ZDIV2] + L.size[MB_ZDIV2];
  LaunchScope ls(c, KID_RESET);
  MB_CUDA(cudaMemsetAsync(c.arena + lo, 0, hi - lo, c.stream));
  return 0;
}

// ---------------------------------------------------------------------------
// init_moloch device part: mx2, rmx, rmu, rmv over the whole padded box (the
// ghost rows of msf* were delivered by the host, so no exchange is needed:
// the exchanged values of :269-272 are the same products)          :263-278
// ---------------------------------------------------------------------------
__global__ void moloch_init_static(Geo g, const double* __restrict__ mx, const double* __restrict__ mu,
                                   const double* __restrict__ mv, double* __restrict__ mx2,
                                   double* __restrict__ rmx, double* __restrict__ rmu,
                                   double* __restrict__ rmv, double* __restrict__ w) {
  const long long n = g.plane;
  for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < n;
       id += (long long)gridDim.x * blockDim.x) {
    const double a = mx[id], b = mu[id], cc = mv[id];
    mx2[id] = a * a;
    rmx[id] = (a != 0.0) ? 1.0 / a : 0.0;
    rmu[id] = (b != 0.0) ? 1.0 / b : 0.0;
    rmv[id] = (cc != 0.0) ? 1.0 / cc : 0.0;
    w[id] = 0.0;  // w(:,:,1) = 0
  }
}
int k_init_static(Ctx& c) {
  const Geo& g = c.g;
  LaunchScope ls(c, KID_INIT);
  moloch_init_static<<<148, 256, 0, c.stream>>>(g, c.f[MB_MSFX].p, c.f[MB_MSFU].p, c.f[MB_MSFV].p, c.mx2,
                                                 c.rmx, c.rmu, c.rmv, c.f[MB_W].p);
  MB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// host hand-off: padded device box <-> contiguous staging buffer
// ---------------------------------------------------------------------------
__global__ void moloch_box_copy(Geo g, double* __restrict__ dev, double* __restrict__ stage, int ja, int ia,
                                int ka, int nj, int ni, int nk, int pack) {
  const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y * BY + threadIdx.y;
  if (j >= nj || i >= ni) return;
  for (int k = blockIdx.z; k < nk; k += gridDim.z) {
    const long long d = gidx(g, ja + j, ia + i, ka + k);
    const long long t = ((long long)k * ni + i) * nj + j;
    if (pack) stage[t] = dev[d]; else dev[d] = stage[t];
  }
}
int k_box_copy(Ctx& c, double* dev, double* stage, int ja, int ia, int ka, int nj, int ni, int nk, bool pack,
               cudaStream_t on) {
  dim3 grid((unsigned)((nj + BX - 1) / BX), (unsigned)((ni + BY - 1) / BY), (unsigned)(nk < 64 ? nk : 64));
  if (on) {   // a copy stream of the hand-off: the profiling events belong to the context's stream
    c.launches++;
    moloch_box_copy<<<grid, dim3(BX, BY), 0, on>>>(c.g, dev, stage, ja, ia, ka, nj, ni, nk, pack ? 1 : 0);
  } else {
    LaunchScope ls(c, KID_BOX);
    moloch_box_copy<<<grid, dim3(BX, BY), 0, c.stream>>>(c.g, dev, stage, ja, ia, ka, nj, ni, nk, pack ? 1 : 0);
  }
  MB_CUDA(cudaGetLastError());
  return 0;
}

// last node of a captured graph: the rounds the graph opened are added to the device-side base of the round numbers
__global__ void moloch_seq_bump(unsigned long long* flags, unsigned long long rounds) { flags[6] += rounds; }
int k_seq_bump(Ctx& c, unsigned long long rounds) {
  moloch_seq_bump<<<1, 1, 0, c.stream>>>(c.flags, rounds);
  MB_CUDA(cudaGetLastError());
  return 0;
}

// One slab of the physics hand-off: every array's rows [ia, ia+ni) x [ja, ja+nj) x [ka, ka+nk) between its
// padded device box and its packed (k, i, j) run inside the slab's staging block.  blockIdx.y = array.
__global__ void moloch_slab_copy(Geo g, SlabTable t, double* __restrict__ stage, int pack) {
  const SlabArray a = t.a[blockIdx.y];
  const long long rows = (long long)a.nk * a.ni;
  double* __restrict__ st = stage + a.off;
  for (long long r = (long long)blockIdx.x * blockDim.y + threadIdx.y; r < rows; r += (long long)gridDim.x * blockDim.y) {
    const int k = (int)(r / a.ni), i = (int)(r % a.ni);
    const long long d0 = gidx(g, a.ja, a.ia + i, a.ka + k);
    const long long s0 = (long long)k * a.w + (((a.a0 + k * a.p8) & 127) >> 3) + (long long)i * a.nj;
    for (int j = threadIdx.x; j < a.nj; j += blockDim.x) {
      if (pack) st[s0 + j] = a.dev[d0 + j]; else a.dev[d0 + j] = st[s0 + j];
    }
  }
}
int k_slab_copy(Ctx& c, const SlabTable& t, double* stage, bool pack, cudaStream_t on) {
  if (t.n <= 0) return 0;
  long long rows = 1;
  for (int q = 0; q < t.n; ++q) rows = std::max(rows, (long long)t.a[q].nk * t.a[q].ni);
  const unsigned gx = (unsigned)std::min<long long>((rows + 3) / 4, 148 * 4);
  c.launches++;   // (a hand-off stream: the profiling events belong to the context's stream)
  moloch_slab_copy<<<dim3(gx, (unsigned)t.n), dim3(64, 4), 0, on>>>(c.g, t, stage, pack ? 1 : 0);
  MB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mb
