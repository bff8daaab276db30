// kernels_sound.cu -- the acoustic sub-step of `sound`
// (/root/reference/Main/mod_moloch.F90:568-723) in its round-2 form: three kernels per
// sub-step instead of four, no zdiv2 exchange.
//
//   moloch_sound_div     K3 + K4 + K6 (:582-618, 531-543): partial s, the horizontal
//                        divergence zdiv2 and its 5-point filter in ONE pass of row-marching
//                        warps.  A strip carries a one-cell ring, and on the
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


// moloch_sound_div: a warp owns a strip of R rows x 64 columns of one level (lanes 1..30 hold the 60 useful
// columns, two per lane; lanes 0 and 31 only supply the neighbouring zdiv2 values) and marches through the rows
// i0-1 .. i0+R: the U-point and V-point products, the zdiv2 rows i-1, i, i+1 and the raw u, v rows that the
// partial s needs form rolling register windows, the j-neighbours come from warp shuffles.  No shared memory,
// no barriers; every array is read with 128-bit loads, 12 per row and lane.
constexpr int SD_J = 60;              // useful columns of a strip
constexpr int SD_WARPS = 4;           // strips per CTA, stacked in i

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }
__device__ __forceinline__ double shfl_dn1(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ double shfl_up1(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }

#ifndef MB_SD_MINB
#define MB_SD_MINB 3
#endif
#ifndef MB_SD_DEPTH
#define MB_SD_DEPTH 1
#endif
#ifndef MB_SD_UNROLL
#define MB_SD_UNROLL 1
#endif
constexpr int SD_UNROLL = MB_SD_UNROLL;
template <int R>
__global__ void __launch_bounds__(32 * SD_WARPS, MB_SD_MINB)
moloch_sound_div(Geo g, const double* __restrict__ u, const double* __restrict__ v, double* __restrict__ s,
                 double* __restrict__ w, double* __restrict__ zdiv2, double* __restrict__ zdiv2b,
                 const double* __restrict__ fmz, const double* __restrict__ rfmzu,
                 const double* __restrict__ rfmzv, const double* __restrict__ hx,
                 const double* __restrict__ hy, const double* __restrict__ mx,
                 const double* __restrict__ mx2, const double* __restrict__ rmu,
                 const double* __restrict__ rmv, const double* __restrict__ gzitak,
                 const double* __restrict__ xknu, double dtrdx, double dtrdy, int do_filter, int ring_store,
                 WaitCtl wc) {
  // u, v ghosts of a fused round: the CTAs that reach them wait for the neighbours' word
  halo_sync(wc, 3, g.jde2 - g.jde1 + 1, g.ide2 - g.ide1 + 1, SD_J, R * SD_WARPS);
  const int k = 1 + blockIdx.z;
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  const int i0 = g.ide1 + (blockIdx.y * SD_WARPS + wq) * R;        // first useful row of the strip
  if (i0 > g.ice2) return;
  const int ja = g.jde1 + blockIdx.x * SD_J - 2 + 2 * lane;        // this lane's columns: ja, ja+1
  const int jb = g.j0 + g.NJ - 1, ib = g.i0 + g.NI - 1;             // last column / row of the padded box
  const bool jok = ja + 1 <= jb;                                    // (ja - j0 is even: the pair is in or out)
  const bool rot = g.lrotllr != 0;
  const double* __restrict__ mc = rot ? mx : mx2;
  const long long kbase = (long long)(k - 1) * g.plane;
  const long long col = (long long)(jok ? ja : jb - 1) - g.j0;
  const double xk = xknu[k], gk = gzitak[k];
  const bool useful = lane >= 1 && lane <= 30;
  const bool ina = useful && ja >= g.jci1 && ja <= g.jci2, inb = useful && ja + 1 >= g.jci1 && ja + 1 <= g.jci2;
  const bool exa = useful && ja >= g.jce1 && ja <= g.jce2, exb = useful && ja + 1 >= g.jce1 && ja + 1 <= g.jce2;
  auto row2 = [&](int i) { return (long long)(min(i, ib) - g.i0) * g.NJ + col; };
  // Everything iteration t reads from memory is loaded SD_DEPTH iterations ahead (12 independent 128-bit loads
  // per row and lane in flight while the earlier rows are computed) -- a load that is issued where it is used
  // stalls the in-order warp, also when it hits L2.  Row set iz: the row iz of the U-point arrays, of fmz, the
  // metric factors, hx, hy and u(k-1), v(k-1); the row iz+1 of the V-point arrays.
  struct RowSet { double2 ur, fu, mu, vr, fv, mv, fz, m2, hy, vm, hx, um; };
  auto load_row = [&](int iz) {
    RowSet L;
    const long long o2 = row2(iz), o3 = kbase + o2, p2 = row2(iz + 1), p3 = kbase + p2;
    L.ur = ld2(u + o3); L.fu = ld2(rfmzu + o3);
    L.mu = rot ? make_double2(1.0, 1.0) : ld2(rmu + o2);
    L.vr = ld2(v + p3); L.fv = ld2(rfmzv + p3); L.mv = ld2(rmv + p2);
    L.fz = ld2(fmz + o3); L.m2 = ld2(mc + o2);
    L.hy = ld2(hy + o2); L.hx = ld2(hx + o2);
    L.vm = L.um = make_double2(0.0, 0.0);
    if (k >= 2) { L.vm = ld2(v + o3 - g.plane); L.um = ld2(u + o3 - g.plane); }
    return L;
  };
  double2 vr0 = make_double2(0.0, 0.0), vr1;                        // raw v of rows iz-1, iz
  double2 vmp = vr0, hyp = vr0, hxp = vr0, ump = vr0;               // v(k-1), hy, hx, u(k-1) of row iz-1
  double2 ur1 = vr0; double ur1e = 0.0;                             // raw u of row iz-1 (+ column ja+2)
  double2 zm = vr0, zc = vr0;                                       // zdiv2 of rows iz-2, iz-1
  double2 pvc;                                                      // V-point products of row iz
  {
    const long long o2 = row2(i0 - 1), o3 = kbase + o2;
    vr1 = ld2(v + o3);
    const double2 f = ld2(rfmzv + o3), m = ld2(rmv + o2);
    pvc = make_double2(dtrdy * vr1.x * f.x * m.x, dtrdy * vr1.y * f.y * m.y);
  }
  RowSet cur = load_row(i0 - 1);
#if MB_SD_DEPTH == 2
  RowSet nx1 = load_row(i0);
#endif
#pragma unroll SD_UNROLL
  for (int t = 0; t < R + 2; ++t) {
    const int iz = i0 - 1 + t;                                      // the zdiv2 row computed in this iteration
#if MB_SD_DEPTH == 2
    RowSet nx2 = nx1;
    if (t + 2 < R + 2) nx2 = load_row(iz + 2);
#else
    RowSet nx1 = cur;
    if (t + 1 < R + 2) nx1 = load_row(iz + 1);
#endif
    // ---- zdiv2 of row iz (:602-618) ----
    const double2 ur = cur.ur;
    double2 pu;
    if (rot) pu = make_double2(dtrdx * ur.x * cur.fu.x, dtrdx * ur.y * cur.fu.y);
    else pu = make_double2(dtrdx * ur.x * cur.fu.x * cur.mu.x, dtrdx * ur.y * cur.fu.y * cur.mu.y);
    const double pue = shfl_dn1(pu.x);                              // the product at column ja+2
    const double ure = shfl_dn1(ur.x);
    const double2 vr2 = cur.vr;
    const double2 pvn = make_double2(dtrdy * vr2.x * cur.fv.x * cur.mv.x, dtrdy * vr2.y * cur.fv.y * cur.mv.y);
    const double2 zn = make_double2(cur.fz.x * cur.m2.x * ((pu.y - pu.x) + (pvn.x - pvc.x)),
                                    cur.fz.y * cur.m2.y * ((pue - pu.y) + (pvn.y - pvc.y)));
    const double2 hyn = cur.hy, vmn = cur.vm;
    // ---- outputs of row i = iz-1 ----
    const int i = iz - 1;
    if (t >= 2 && i <= g.ice2) {
      const long long oi3 = kbase + row2(i);
      const double zw = shfl_up1(zc.y), ze = shfl_dn1(zc.x);       // zdiv2 at columns ja-1, ja+2
      if (exa && exb) st2(zdiv2 + oi3, zc.x, zc.y);
      else if (exa) zdiv2[oi3] = zc.x;
      else if (exb) zdiv2[oi3 + 1] = zc.y;
      const bool rowin = (i >= g.ici1 && i <= g.ici2);
      if (do_filter && rowin) {   // :536-542 (Jacobi: old values everywhere)
        const double fa = zc.x + xk * (zw + zc.y + zm.x + zn.x - 4.0 * zc.x);
        const double fb = zc.y + xk * (zc.x + ze + zm.y + zn.y - 4.0 * zc.y);
        if (ina && inb) st2(zdiv2b + oi3, fa, fb);
        else if (ina) zdiv2b[oi3] = fa;
        else if (inb) zdiv2b[oi3 + 1] = fb;
      }
      // :582-597 partial s (all lanes shuffle, the interior ones store)
      const double2 hx01 = hxp, um01 = ump;
      const double hx2 = shfl_dn1(hx01.x);
      const double um2 = shfl_dn1(um01.x);
      if (rowin) {
        if (k >= 2) {
          const double zuha = (ur1.x + um01.x) * hx01.x + (ur1.y + um01.y) * hx01.y;
          const double zvha = (vr0.x + vmp.x) * hyp.x + (vr1.x + vmn.x) * hyn.x;
          const double zuhb = (ur1.y + um01.y) * hx01.y + (ur1e + um2) * hx2;
          const double zvhb = (vr0.y + vmp.y) * hyp.y + (vr1.y + vmn.y) * hyn.y;
          const double sa = -0.25 * (zuha + zvha) * gk, sb = -0.25 * (zuhb + zvhb) * gk;
          if (ina && inb) st2(s + oi3, sa, sb);
          else if (ina) s[oi3] = sa;
          else if (inb) s[oi3 + 1] = sb;
        }
        if (k == g.kz) {
          const double ska = -0.5 * ((ur1.x * hx01.x + ur1.y * hx01.y) + (vr0.x * hyp.x + vr1.x * hyn.x));
          const double skb = -0.5 * ((ur1.y * hx01.y + ur1e * hx2) + (vr0.y * hyp.y + vr1.y * hyn.y));
          const long long op = oi3 + g.plane;
          if (ina) { s[op] = ska; w[op] = -ska; }
          if (inb) { s[op + 1] = skb; w[op + 1] = -skb; }
        }
      }
      // the ring cells uvupdate2's damping reads: zdiv2(jce1-1, i) and zdiv2(j, ice1-1)
      if (ring_store && blockIdx.x == 0 && lane == 0 && g.gl) zdiv2[oi3 + 1] = zc.y;
    }
    // roll the windows
    zm = zc; zc = zn; pvc = pvn;
    vr0 = vr1; vr1 = vr2; vmp = vmn; hyp = hyn; hxp = cur.hx; ump = cur.um;
    ur1 = ur; ur1e = ure;
#if MB_SD_DEPTH == 2
    cur = nx1; nx1 = nx2;
#else
    cur = nx1;
#endif
    if (ring_store && t == 0 && g.gb && i0 == g.ide1) {   // zc is now the bottom ring row ice1-1
      const long long o = kbase + row2(i0 - 1);
      if (exa && exb) st2(zdiv2 + o, zc.x, zc.y);
      else if (exa) zdiv2[o] = zc.x;
      else if (exb) zdiv2[o + 1] = zc.y;
    }
  }
}

template <int R>
static int launch_sound_div(Ctx& c, double dts, const WaitCtl& w0) {
  const Geo& g = c.g;
  const double dtrdx = dts * c.rdx, dtrdy = dts * c.rdx;
  const int nj = g.jde2 - g.jde1 + 1, ni = g.ide2 - g.ide1 + 1;
  const dim3 grid((unsigned)((nj + SD_J - 1) / SD_J), (unsigned)((ni + R * SD_WARPS - 1) / (R * SD_WARPS)), (unsigned)g.kz);
  LaunchScope ls(c, KID_SOUND_PRE);
  moloch_sound_div<R><<<grid, 32 * SD_WARPS, 0, c.stream>>>(
      g, c.f[MB_U].p, c.f[MB_V].p, c.f[MB_S].p, c.f[MB_W].p, c.f[MB_ZDIV2].p, c.zdiv2b, c.f[MB_FMZ].p,
      c.f[MB_RFMZU].p, c.f[MB_RFMZV].p, c.f[MB_HX].p, c.f[MB_HY].p, c.f[MB_MSFX].p, c.mx2, c.rmu, c.rmv,
      c.prof[MB_GZITAK], c.prof[MB_XKNU], dtrdx, dtrdy, c.cfg.mo_divfilter ? 1 : 0, c.cfg.mo_divdamp ? 1 : 0, w0);
  MB_CUDA(cudaGetLastError());
  return 0;
}
int k_sound_div(Ctx& c, double dts, const WaitCtl* wc) {
  const WaitCtl w0 = wc ? *wc : WaitCtl{};
  // small per-GPU grids: shorter strips, so that the warps still cover all SMs
  const long long strips8 = (long long)((c.g.jde2 - c.g.jde1 + SD_J) / SD_J) * ((c.g.ide2 - c.g.ide1 + 8) / 8) * c.g.kz;
#ifndef MB_SD_R
#define MB_SD_R 8
#endif
  return strips8 < 148 * 48 ? launch_sound_div<4>(c, dts, w0) : launch_sound_div<MB_SD_R>(c, dts, w0);
}

// ---------------------------------------------------------------------------
// K5 + K10  divergence damping (:746-764) and horizontal momentum update (:677-721)
// ---------------------------------------------------------------------------
// Two cells (j, j+1) per thread, every 3-D array read with one 128-bit load per thread; the value at j-1
// (tetav, pai, zdiv2: the U-point differences) comes from the lane to the left, lane 0 loads it.
constexpr int UBX = 32, UBY = 4;      // 64 columns x 4 rows per CTA
template <bool FUSED>
__device__ __forceinline__ void
uvupdate2_cells(const Geo& g, double* __restrict__ u, double* __restrict__ v, const double* __restrict__ zdiv2,
                const double* __restrict__ tetav, const double* __restrict__ pai,
                const double* __restrict__ bdywtu, const double* __restrict__ bdywtv,
                const double* __restrict__ coru, const double* __restrict__ corv,
                const double* __restrict__ hx, const double* __restrict__ hy,
                const double* __restrict__ mu, const double* __restrict__ mv,
                const double* __restrict__ gzitakh, const double* __restrict__ xkdamp, double dts,
                double dtrdx, double dtrdy, double dxrdt, int damped, const PushCtl& pc, const EdgePush& eu,
                const EdgePush& ev) {
  const int lane = threadIdx.x;
  const int ja = g.jde1 + (blockIdx.x * UBX + lane) * 2;
  const int i = g.ide1 + blockIdx.y * UBY + threadIdx.y;
  const int k = 1 + blockIdx.z;
  if (i > g.ide2) return;                                   // (warp-uniform: a warp is one row)
  const bool jin = ja <= g.jde2;                            // ja - j0 is even and NJ is: ja+1 is in the box then
  const long long i2 = gidx2(g, jin ? ja : g.jde1, i);
  const long long id = (long long)(k - 1) * g.plane + i2;
  // every load of the thread is issued before the first use (all addresses lie in the padded box):
  // 16 independent 128-bit loads in flight per thread
  const double2 tv = ld2(tetav + id), pa = ld2(pai + id);
  const double2 tvs = ld2(tetav + id - g.NJ), pas = ld2(pai + id - g.NJ);
  const double2 uo = ld2(u + id), vo = ld2(v + id);
  const double2 bwu = ld2(bdywtu + id), bwv = ld2(bdywtv + id);
  double2 zz = make_double2(0.0, 0.0), zs = zz;
  if (damped) { zz = ld2(zdiv2 + id); zs = ld2(zdiv2 + id - g.NJ); }
  const double2 mju = ld2(mu + i2), cou = ld2(coru + i2), hxx = ld2(hx + i2);
  const double2 cov = ld2(corv + i2), hyy = ld2(hy + i2);
  double2 mjv = make_double2(1.0, 1.0);
  if (!g.lrotllr) mjv = ld2(mv + i2);
  const double gk = gzitakh[k];
  const double xkd = damped ? dxrdt * xkdamp[k] : 0.0;
  // column ja-1: the left lane's second value (lane 0: its own load)
  double tvw = shfl_up1(tv.y), paw = shfl_up1(pa.y), zw = shfl_up1(zz.y);
  if (lane == 0) { tvw = tetav[id - 1]; paw = pai[id - 1]; if (damped) zw = zdiv2[id - 1]; }
  if (!jin) return;
  const bool rowu = (i >= g.ici1 && i <= g.ici2), rowv = (i >= g.idi1 && i <= g.idi2);
  const bool dua = rowu && ja >= g.jdi1 && ja <= g.jdi2, dub = rowu && ja + 1 >= g.jdi1 && ja + 1 <= g.jdi2;
  const bool dva = rowv && ja >= g.jci1 && ja <= g.jci2, dvb = rowv && ja + 1 >= g.jci1 && ja + 1 <= g.jci2;
  const double zfz = egrav * dts;
  // uo, vo: u, v of the start of the sub-step -- the reference's ud, vd (:573-578), which the Coriolis terms use
  if (dua || dub) {
    double ua = uo.x, ub = uo.y;
    if (damped) {   // :746-750
      ua = uo.x + xkd * mju.x * (zz.x - zw);
      ub = uo.y + xkd * mju.y * (zz.y - zz.x);
    }
    const double una = ua + bwu.x * (cou.x * dts * vo.x - zfz * hxx.x * gk - dtrdx * mju.x * (0.5 * cpd * (tvw + tv.x)) * (pa.x - paw));
    const double unb = ub + bwu.y * (cou.y * dts * vo.y - zfz * hxx.y * gk - dtrdx * mju.y * (0.5 * cpd * (tv.x + tv.y)) * (pa.y - pa.x));
    if (dua && dub) st2(u + id, una, unb);
    else if (dua) u[id] = una;
    else u[id + 1] = unb;
    if (FUSED && pc.mask) {
      if (dua) edge_push(pc, eu, ja, i, k, una);
      if (dub) edge_push(pc, eu, ja + 1, i, k, unb);
    }
  }
  if (dva || dvb) {
    double va = vo.x, vb = vo.y;
    if (damped) {   // :752-763
      const double xa = g.lrotllr ? xkd : xkd * mjv.x, xb = g.lrotllr ? xkd : xkd * mjv.y;
      va = vo.x + xa * (zz.x - zs.x);
      vb = vo.y + xb * (zz.y - zs.y);
    }
    const double zcya = g.lrotllr ? dtrdy : dtrdy * mjv.x, zcyb = g.lrotllr ? dtrdy : dtrdy * mjv.y;
    const double vna = va + bwv.x * (-(cov.x * dts * uo.x) - zfz * hyy.x * gk - zcya * (0.5 * cpd * (tvs.x + tv.x)) * (pa.x - pas.x));
    const double vnb = vb + bwv.y * (-(cov.y * dts * uo.y) - zfz * hyy.y * gk - zcyb * (0.5 * cpd * (tvs.y + tv.y)) * (pa.y - pas.y));
    if (dva && dvb) st2(v + id, vna, vnb);
    else if (dva) v[id] = vna;
    else v[id + 1] = vnb;
    if (FUSED && pc.mask) {
      if (dva) edge_push(pc, ev, ja, i, k, vna);
      if (dvb) edge_push(pc, ev, ja + 1, i, k, vnb);
    }
  }
}
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
  // (the cells return early on rows / columns outside the box; the producer's count below needs every thread)
  uvupdate2_cells<FUSED>(g, u, v, zdiv2, tetav, pai, bdywtu, bdywtv, coru, corv, hx, hy, mu, mv, gzitakh, xkdamp, dts,
                         dtrdx, dtrdy, dxrdt, damped, pc, eu, ev);
  if (FUSED) halo_producer_done(pc, blockIdx.y, gridDim.y, UBY, 1, g.ide2 - g.ide1 + 1, gridDim.x * gridDim.z);
}

int k_uvupdate2(Ctx& c, double dts, const WaitCtl* wc, const PushCtl* pc, const EdgePush* eu, const EdgePush* ev) {
  const Geo& g = c.g;
  const WaitCtl w0 = wc ? *wc : WaitCtl{};
  const PushCtl p0 = pc ? *pc : PushCtl{};
  const EdgePush e0 = eu ? *eu : EdgePush{}, e1 = ev ? *ev : EdgePush{};
  LaunchScope ls(c, KID_UVUPDATE);
  const dim3 grid((unsigned)((g.jde2 - g.jde1 + 1 + 2 * UBX - 1) / (2 * UBX)), (unsigned)((g.ide2 - g.ide1 + 1 + UBY - 1) / UBY),
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
