// kernels_sound.cu -- the acoustic sub-step of `sound`
// (/root/reference/Main/mod_moloch.F90:568-723) in its round-2 form: three kernels per
// sub-step instead of four, no zdiv2 exchange.
//
//   moloch_sound_div     K3 + K4 + K6 (:582-618, 531-543): partial s, the horizontal
//                        divergence zdiv2 and its 5-point filter in ONE pass over a
//                        shared-memory tile.  The tile carries a one-cell ring, and on the
//                        sides of the rank that have a neighbour the ring cells are the
//                        neighbour's cells, computed here from u, v ghosts (the same
//                        expressions on the same values: bit-identical to what the
//                        neighbour computes).  The reference's exchange of zdiv2
//                        (:745, and the redundant :535) therefore does not exist on the
//                        device: u and v travel a little wider instead (u: 2 columns / 1
//                        row, v: 1 column / 2 rows), once per sub-step.
//   moloch_wsolve*       K7 + K8 + K9 (kernels.cu)
//   moloch_uvupdate2     K5 + K10 (:738-765, 677-721): divergence damping applied on the
//                        fly (u + xdam*(zdiv2(j)-zdiv2(j-1)) is formed in registers, the
//                        snapshot copies ud/vd of :573-578 and the damped copies of round 1
//                        are gone), then the momentum update.
//
// Algorithmic traffic per cell and sub-step (SURVEY.md App. D counts 9 + 5 + 2 + 10 = 26 doubles
// for K2..K6 + K10): sound_div reads u, v, rfmzu, rfmzv, fmz and writes s, zdiv2, zdiv2b (8),
// uvupdate2 reads u, v, zdiv2, tetav, pai, bdywtu, bdywtv and writes u, v (9).
//
// Same arithmetic and operation order as the reference loops; compiled -fmad=false.
#include "common.cuh"

namespace mb {

#ifdef MB_HOST_EMU
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
#endif

constexpr int SA_TJ = 64;             // core tile: 64 columns x 8 rows, 256 threads, two cells per thread
constexpr int SA_TI = 8;
constexpr int SA_NT = 256;
constexpr int SA_W = SA_TJ + 6;       // shared-memory row; column c <-> j = jt - 2 + c (core starts at c = 2)
constexpr int SA_RU = SA_TI + 2;      // rows of the U arrays and of Z:  r <-> i = it - 1 + r
constexpr int SA_RV = SA_TI + 3;      // rows of the V arrays (one more: v(i+1) of the top ring row)
constexpr int SA_ZC = SA_TJ + 2;      // Z columns per row: c = 1 .. TJ+2  (j = jt-1 .. jt+TJ)

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }

__global__ void __launch_bounds__(SA_NT)
moloch_sound_div(Geo g, const double* __restrict__ u, const double* __restrict__ v, double* __restrict__ s,
                 double* __restrict__ w, double* __restrict__ zdiv2, double* __restrict__ zdiv2b,
                 const double* __restrict__ fmz, const double* __restrict__ rfmzu,
                 const double* __restrict__ rfmzv, const double* __restrict__ hx,
                 const double* __restrict__ hy, const double* __restrict__ mx,
                 const double* __restrict__ mx2, const double* __restrict__ rmu,
                 const double* __restrict__ rmv, const double* __restrict__ gzitak,
                 const double* __restrict__ xknu, double dtrdx, double dtrdy, int do_filter, int ring_store,
                 WaitCtl wc) {
  // u, v ghosts of a fused round: the tiles that reach them wait for the neighbours' word
  halo_sync(wc, 3, g.jde2 - g.jde1 + 1, g.ide2 - g.ide1 + 1, SA_TJ, SA_TI);
  extern __shared__ double sm[];
  double* RU = sm;                         // u                       [SA_RU][SA_W]
  double* PU = RU + SA_RU * SA_W;          // dtrdx*u*rfmzu[*rmu]
  double* RV = PU + SA_RU * SA_W;          // v                       [SA_RV][SA_W]
  double* PV = RV + SA_RV * SA_W;          // dtrdy*v*rfmzv*rmv
  double* Z = PV + SA_RV * SA_W;           // zdiv2                   [SA_RU][SA_W]
  const int k = 1 + blockIdx.z;
  const int jt = g.jde1 + blockIdx.x * SA_TJ, it = g.ide1 + blockIdx.y * SA_TI;
  const int tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;
  const int jb = g.j0 + g.NJ - 1, ib = g.i0 + g.NI - 1;      // last column / row of the padded box
  const long long kbase = (long long)(k - 1) * g.plane;
  const bool rot = g.lrotllr != 0;

  // ---- phase 1: u, v of level k and their metric products, one warp per row --------------------
  for (int t = wq; t < SA_RU + SA_RV; t += SA_NT / 32) {
    const bool isu = t < SA_RU;
    const int r = isu ? t : t - SA_RU;
    const int i = min(it - 1 + r, ib);
    const double* __restrict__ f3 = isu ? u : v;
    const double* __restrict__ m3 = isu ? rfmzu : rfmzv;
    const double* __restrict__ m2 = isu ? rmu : rmv;
    const double dtr = isu ? dtrdx : dtrdy;
    const bool use_m2 = isu ? !rot : true;
    double* raw = (isu ? RU : RV) + r * SA_W;
    double* prd = (isu ? PU : PV) + r * SA_W;
    const long long row2 = (long long)(i - g.i0) * g.NJ - g.j0;     // + j
    const long long row3 = kbase + row2;
    {
      const int j = jt + 2 * lane;
      double2 a = make_double2(0.0, 0.0), b = a, m = make_double2(1.0, 1.0);
      if (j + 1 <= jb) {
        a = ld2(f3 + row3 + j); b = ld2(m3 + row3 + j);
        if (use_m2) m = ld2(m2 + row2 + j);
      } else if (j <= jb) {
        a.x = f3[row3 + j]; b.x = m3[row3 + j];
        if (use_m2) m.x = m2[row2 + j];
      }
      const int c = 2 + 2 * lane;
      raw[c] = a.x; raw[c + 1] = a.y;
      prd[c] = use_m2 ? dtr * a.x * b.x * m.x : dtr * a.x * b.x;
      prd[c + 1] = use_m2 ? dtr * a.y * b.y * m.y : dtr * a.y * b.y;
    }
    if (lane < 3) {        // the ring columns: j = jt-1 (c = 1), jt+TJ (c = TJ+2), jt+TJ+1 (c = TJ+3, u only)
      const int c = (lane == 0) ? 1 : SA_TJ + 1 + lane;
      const int j = jt - 2 + c;
      double a = 0.0, b = 0.0, m = 1.0;
      if (j <= jb) {
        a = f3[row3 + j]; b = m3[row3 + j];
        if (use_m2) m = m2[row2 + j];
      }
      raw[c] = a;
      prd[c] = use_m2 ? dtr * a * b * m : dtr * a * b;
    }
  }
  __syncthreads();
  // ---- phase 2: zdiv2 on the tile and its ring (:602-618) ---------------------------------------
  const double* __restrict__ mc = rot ? mx : mx2;
  for (int p = tid; p < SA_RU * SA_ZC; p += SA_NT) {
    const int r = p / SA_ZC, c = 1 + p % SA_ZC;
    const int j = min(jt - 2 + c, jb), i = min(it - 1 + r, ib);
    const long long i2 = (long long)(i - g.i0) * g.NJ + (j - g.j0);
    const double zum = PU[r * SA_W + c], zup = PU[r * SA_W + c + 1];
    const double zvm = PV[r * SA_W + c], zvp = PV[(r + 1) * SA_W + c];
    Z[r * SA_W + c] = fmz[kbase + i2] * mc[i2] * ((zup - zum) + (zvp - zvm));
  }
  __syncthreads();
  // ---- phase 3: outputs of the thread's two cells ------------------------------------------------
  const int tx = tid & 31, ty = tid >> 5;
  const int i = it + ty, r = ty + 1;
  const int ja = jt + 2 * tx, c = 2 + 2 * tx;
  if (i > g.ice2) return;
  const long long ida = kbase + (long long)(i - g.i0) * g.NJ + (ja - g.j0);
  const double za = Z[r * SA_W + c], zb = Z[r * SA_W + c + 1];
  // zdiv2 on the external cross range; the left / bottom ring cells that uvupdate2's damping reads
  // (zdiv2(j-1), zdiv2(i-1) at the first owned column / row) are stored by the edge tiles
  if (ja + 1 <= g.jce2) *reinterpret_cast<double2*>(zdiv2 + ida) = make_double2(za, zb);
  else if (ja <= g.jce2) zdiv2[ida] = za;
  if (ring_store) {
    if (blockIdx.x == 0 && tx == 0 && g.gl) zdiv2[ida - 1] = Z[r * SA_W + 1];
    if (blockIdx.y == 0 && ty == 0 && g.gb) {
      if (ja <= g.jce2) zdiv2[ida - g.NJ] = Z[c];
      if (ja + 1 <= g.jce2) zdiv2[ida - g.NJ + 1] = Z[c + 1];
    }
  }
  const bool rowin = (i >= g.ici1 && i <= g.ici2);
  const bool ina = rowin && ja >= g.jci1 && ja <= g.jci2, inb = rowin && ja + 1 >= g.jci1 && ja + 1 <= g.jci2;
  if (!ina && !inb) return;
  if (do_filter) {   // :536-542 (Jacobi: old values everywhere)
    const double xk = xknu[k];
    const double fa = za + xk * (Z[r * SA_W + c - 1] + zb + Z[(r - 1) * SA_W + c] + Z[(r + 1) * SA_W + c] - 4.0 * za);
    const double fb = zb + xk * (za + Z[r * SA_W + c + 2] + Z[(r - 1) * SA_W + c + 1] + Z[(r + 1) * SA_W + c + 1] - 4.0 * zb);
    if (ina && inb) *reinterpret_cast<double2*>(zdiv2b + ida) = make_double2(fa, fb);
    else if (ina) zdiv2b[ida] = fa;
    else zdiv2b[ida + 1] = fb;
  }
  // :582-597 partial s
  const long long i2a = (long long)(i - g.i0) * g.NJ + (ja - g.j0);
  const double u0 = RU[r * SA_W + c], u1 = RU[r * SA_W + c + 1], u2 = RU[r * SA_W + c + 2];
  const double va0 = RV[r * SA_W + c], va1 = RV[(r + 1) * SA_W + c];
  const double vb0 = RV[r * SA_W + c + 1], vb1 = RV[(r + 1) * SA_W + c + 1];
  const double2 hx01 = ld2(hx + i2a);
  const double hx2 = hx[min(i2a + 2, (long long)g.plane - 1)];
  const double2 hy0 = ld2(hy + i2a), hy1 = ld2(hy + i2a + g.NJ);
  if (k >= 2) {
    const long long im = ida - g.plane;
    const double2 um01 = ld2(u + im);
    const double um2 = (ja + 2 <= jb) ? u[im + 2] : 0.0;
    const double2 vm0 = ld2(v + im), vm1 = ld2(v + im + g.NJ);
    const double gk = gzitak[k];
    const double zuha = (u0 + um01.x) * hx01.x + (u1 + um01.y) * hx01.y;
    const double zvha = (va0 + vm0.x) * hy0.x + (va1 + vm1.x) * hy1.x;
    const double zuhb = (u1 + um01.y) * hx01.y + (u2 + um2) * hx2;
    const double zvhb = (vb0 + vm0.y) * hy0.y + (vb1 + vm1.y) * hy1.y;
    const double sa = -0.25 * (zuha + zvha) * gk, sb = -0.25 * (zuhb + zvhb) * gk;
    if (ina && inb) *reinterpret_cast<double2*>(s + ida) = make_double2(sa, sb);
    else if (ina) s[ida] = sa;
    else s[ida + 1] = sb;
  }
  if (k == g.kz) {
    const double ska = -0.5 * ((u0 * hx01.x + u1 * hx01.y) + (va0 * hy0.x + va1 * hy1.x));
    const double skb = -0.5 * ((u1 * hx01.y + u2 * hx2) + (vb0 * hy0.y + vb1 * hy1.y));
    const long long idp = ida + g.plane;
    if (ina) { s[idp] = ska; w[idp] = -ska; }
    if (inb) { s[idp + 1] = skb; w[idp + 1] = -skb; }
  }
}

int k_sound_div(Ctx& c, double dts, const WaitCtl* wc) {
  const Geo& g = c.g;
  const double dtrdx = dts * c.rdx, dtrdy = dts * c.rdx;
  const WaitCtl w0 = wc ? *wc : WaitCtl{};
  const size_t smem = (size_t)(3 * SA_RU + 2 * SA_RV) * SA_W * sizeof(double);
  const dim3 grid((unsigned)((g.jde2 - g.jde1 + 1 + SA_TJ - 1) / SA_TJ), (unsigned)((g.ide2 - g.ide1 + 1 + SA_TI - 1) / SA_TI),
                  (unsigned)g.kz);
  LaunchScope ls(c, KID_SOUND_PRE);
  moloch_sound_div<<<grid, SA_NT, smem, c.stream>>>(
      g, c.f[MB_U].p, c.f[MB_V].p, c.f[MB_S].p, c.f[MB_W].p, c.f[MB_ZDIV2].p, c.zdiv2b, c.f[MB_FMZ].p,
      c.f[MB_RFMZU].p, c.f[MB_RFMZV].p, c.f[MB_HX].p, c.f[MB_HY].p, c.f[MB_MSFX].p, c.mx2, c.rmu, c.rmv,
      c.prof[MB_GZITAK], c.prof[MB_XKNU], dtrdx, dtrdy, c.cfg.mo_divfilter ? 1 : 0, c.cfg.mo_divdamp ? 1 : 0, w0);
  MB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// K5 + K10  divergence damping (:746-764) and horizontal momentum update (:677-721)
// ---------------------------------------------------------------------------
constexpr int UBX = 32, UBY = 8;
template <bool FUSED>
__global__ void __launch_bounds__(UBX * UBY)
moloch_uvupdate2(Geo g, double* __restrict__ u, double* __restrict__ v, const double* __restrict__ zdiv2,
                 const double* __restrict__ tetav, const double* __restrict__ pai,
                 const double* __restrict__ bdywtu, const double* __restrict__ bdywtv,
                 const double* __restrict__ coru, const double* __restrict__ corv,
                 const double* __restrict__ hx, const double* __restrict__ hy,
                 const double* __restrict__ mu, const double* __restrict__ mv,
                 const double* __restrict__ gzitakh, const double* __restrict__ xkdamp, double dts,
                 double dtrdx, double dtrdy, double dxrdt, int damped, WaitCtl wc, PushCtl pc, EdgePush eu,
                 EdgePush ev) {
  if (FUSED) halo_sync(wc);   // pai ghosts of a fused round
  const int j = g.jde1 + blockIdx.x * UBX + threadIdx.x;
  const int i = g.ide1 + blockIdx.y * UBY + threadIdx.y;
  const int k = 1 + blockIdx.z;
  const bool inside = (j <= g.jde2 && i <= g.ide2);
  const bool du = inside && j >= g.jdi1 && j <= g.jdi2 && i >= g.ici1 && i <= g.ici2;
  const bool dv = inside && j >= g.jci1 && j <= g.jci2 && i >= g.idi1 && i <= g.idi2;
  if (!(du || dv)) return;
  const long long id = gidx(g, j, i, k);
  const long long i2 = gidx2(g, j, i);
  const double tv0 = tetav[id], pai0 = pai[id];
  const double zfz = egrav * dts;
  const double gk = gzitakh[k];
  // u, v of the start of the sub-step: the reference's ud, vd (:573-578), which the Coriolis terms use
  const double uold = u[id], vold = v[id];
  const double z0 = damped ? zdiv2[id] : 0.0;
  if (du) {
    double ub = uold;
    if (damped) {   // :746-750
      const double xdam = dxrdt * xkdamp[k] * mu[i2];
      ub = uold + xdam * (z0 - zdiv2[id - 1]);
    }
    const double zcx = dtrdx * mu[i2];
    const double zrom1u = 0.5 * cpd * (tetav[id - 1] + tv0);
    const double zcor1u = coru[i2] * dts * vold;
    const double un = ub + bdywtu[id] * (zcor1u - zfz * hx[i2] * gk - zcx * zrom1u * (pai0 - pai[id - 1]));
    u[id] = un;
    if (FUSED && pc.mask) edge_push(pc, eu, j, i, k, un);
  }
  if (dv) {
    double vb = vold;
    if (damped) {   // :752-763
      const double xdam = g.lrotllr ? dxrdt * xkdamp[k] : dxrdt * xkdamp[k] * mv[i2];
      vb = vold + xdam * (z0 - zdiv2[id - g.NJ]);
    }
    const double zcy = g.lrotllr ? dtrdy : dtrdy * mv[i2];
    const double zrom1v = 0.5 * cpd * (tetav[id - g.NJ] + tv0);
    const double zcor1v = corv[i2] * dts * uold;
    const double vn = vb + bdywtv[id] * (-zcor1v - zfz * hy[i2] * gk - zcy * zrom1v * (pai0 - pai[id - g.NJ]));
    v[id] = vn;
    if (FUSED && pc.mask) edge_push(pc, ev, j, i, k, vn);
  }
}

int k_uvupdate2(Ctx& c, double dts, const WaitCtl* wc, const PushCtl* pc, const EdgePush* eu, const EdgePush* ev) {
  const Geo& g = c.g;
  const WaitCtl w0 = wc ? *wc : WaitCtl{};
  const PushCtl p0 = pc ? *pc : PushCtl{};
  const EdgePush e0 = eu ? *eu : EdgePush{}, e1 = ev ? *ev : EdgePush{};
  LaunchScope ls(c, KID_UVUPDATE);
  const dim3 grid((unsigned)((g.jde2 - g.jde1 + 1 + UBX - 1) / UBX), (unsigned)((g.ide2 - g.ide1 + 1 + UBY - 1) / UBY),
                  (unsigned)g.kz);
#define UV2_ARGS g, c.f[MB_U].p, c.f[MB_V].p, c.f[MB_ZDIV2].p, c.f[MB_TETAV].p, c.f[MB_PAI].p, c.f[MB_BDYWTU].p, \
      c.f[MB_BDYWTV].p, c.f[MB_CORU].p, c.f[MB_CORV].p, c.f[MB_HX].p, c.f[MB_HY].p, c.f[MB_MSFU].p, c.f[MB_MSFV].p, \
      c.prof[MB_GZITAKH], c.prof[MB_XKDAMP], dts, dts * c.rdx, dts * c.rdx, c.cfg.dx / dts, c.cfg.mo_divdamp ? 1 : 0, \
      w0, p0, e0, e1
  if (w0.mask || p0.mask) moloch_uvupdate2<true><<<grid, dim3(UBX, UBY), 0, c.stream>>>(UV2_ARGS);
  else moloch_uvupdate2<false><<<grid, dim3(UBX, UBY), 0, c.stream>>>(UV2_ARGS);
#undef UV2_ARGS
  MB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mb
