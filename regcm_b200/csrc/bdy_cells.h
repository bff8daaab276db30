// bdy_cells.h -- per-cell bodies of the lateral-boundary, mkslice and TKE
// kernels.
//
// Each function restates the work one (j,i,k) iteration of a reference loop
// does; kernels_bdy.cu maps CUDA threads onto them.  The file is host/device
// neutral so that the CPU test-suite can instantiate exactly the same bodies
// with g++ (tests/emu) and compare them bit for bit with the oracle without a
// GPU; the product never runs them on the host.
//
// Reference: Main/mod_bdycod.F90 (bdyval MOLOCH branch :1618-1875, motopnudge
// :4049-4081, morelax_external/_fraction :3962-4031, mospectral_nudge
// :3898-3960), Main/chemlib/mod_che_bdyco.F90 (:391-535, :965-1026),
// Main/mod_slice.F90:115-173, Main/mod_moloch.F90 (boundary :448-529, zstagtoh
// :1445-1459, htozstag :1461-1475, status_update :1419-1424).
#pragma once
#include <math.h>
#include "geo.h"

namespace mb {

// Main/mpplib/mod_runparams.F90:184-192 (n = 0-based species index)
MB_HD double qx_checkval(int n) { return n == 0 ? 1.0e-8 : (n <= 6 ? 1.0e-16 : (n == 7 ? 1.0e10 : (n == 8 ? 100.0 : 0.01))); }
MB_HD double qx_zeroval(int n) { return n == 0 ? 1.0e-8 : (n <= 6 ? 0.0 : (n == 7 ? 1.0e10 : (n == 8 ? 100.0 : 0.01))); }

struct BdyArgs {
  Geo g;
  // state (mo_atm, sfs%psa)
  double *u, *v, *w, *t, *pai, *qx, *trac, *ps, *tke, *ux, *vx, *tvirt, *tetav;
  // v3dbound / v2dbound b0, b1
  const double *dub0, *dub1, *dvb0, *dvb1, *xtb0, *xtb1, *xpaib0, *xpaib1, *xqb0, *xqb1, *xlb0, *xlb1, *xib0,
      *xib1, *xpsb0, *xpsb1, *chib0, *chib1;
  const int *ib_cr, *ib_ud, *ib_vd;   // ba_cr/ba_ud/ba_vd %ibnd on the padded plane
  const double *hefc, *tnudge, *fcx;  // hefc[(k-1)*nspgx + ib-1], tnudge[k-1], fcx[ib-1]
  double x0, x1;      // time weights of this call: x1 = (xbctime+dt)*rtb
  double xc0, xc1;    // tracer weights: x1 = (xbctime+dt)/dtbdys (mod_che_bdyco.F90:507)
  double dtsec, tkemin;
  int nspgx, iqfrst, present_qc, present_qi, tke_on, nztop, top_nudge, ichem, ichebdy;
};

MB_HD double lin2(const double* b0, const double* b1, long long id, double x0, double x1) {
  return x0 * b0[id] + x1 * b1[id];
}

// ---- bdyval, west (side 0) / east (side 1): corners excluded  :1641-1767 -----
// One call = one (i,k) of the boundary column; i runs over ide1:ide2.
MB_HD void bdyval_we_cell(const BdyArgs& a, int side, int i, int k) {
  const Geo& g = a.g;
  if (side == 0 ? !g.bl : !g.br) return;
  const int jd = side ? g.jde2 : g.jde1, jc = side ? g.jce2 : g.jce1, jin = side ? g.jci2 : g.jci1;
  const double sgn = side ? -1.0 : 1.0;   // inflow: u > 0 on the west, u < 0 on the east side
  const bool in_ic = i >= g.ici1 && i <= g.ici2, in_id = i >= g.idi1 && i <= g.idi2;
  const long long idd = gidx(g, jd, i, k), idc = gidx(g, jc, i, k), idin = gidx(g, jin, i, k);
  const long long sp = (long long)g.kz * g.plane;
  if (in_id) a.v[idc] = lin2(a.dvb0, a.dvb1, idc, a.x0, a.x1);
  if (!in_ic) return;
  if (k == 1) { const long long i2 = gidx2(g, jc, i); a.ps[i2] = a.x0 * a.xpsb0[i2] + a.x1 * a.xpsb1[i2]; }
  const double ub = lin2(a.dub0, a.dub1, idd, a.x0, a.x1);
  a.u[idd] = ub;
  a.t[idc] = lin2(a.xtb0, a.xtb1, idc, a.x0, a.x1);
  a.pai[idc] = lin2(a.xpaib0, a.xpaib1, idc, a.x0, a.x1);
  a.qx[idc] = lin2(a.xqb0, a.xqb1, idc, a.x0, a.x1);
  if (a.present_qc) a.qx[idc + sp] = lin2(a.xlb0, a.xlb1, idc, a.x0, a.x1);
  if (a.present_qi && g.ipptls > 1) a.qx[idc + 2 * sp] = lin2(a.xib0, a.xib1, idc, a.x0, a.x1);
  const bool inflow = sgn * ub > 0.0;
  for (int n = a.iqfrst; n <= g.nqx; ++n) {
    if ((a.present_qc && n == 2) || (a.present_qi && n == 3)) continue;
    const double qxint = a.qx[idin + (n - 1) * sp];
    a.qx[idc + (n - 1) * sp] = inflow ? qx_zeroval(n - 1) : qxint;
  }
  a.w[idc] = inflow ? 0.0 : a.w[idin];
  if (a.tke_on) {
    if (k == 1) a.tke[idc] = a.tkemin;
    else {
      // u(jd,i,k-1) is written by another thread of this launch: re-evaluate it
      const double ukm1 = lin2(a.dub0, a.dub1, idd - g.plane, a.x0, a.x1);
      a.tke[idc + g.plane] = (sgn * (ub + ukm1) > 0.0) ? a.tkemin : a.tke[idin + g.plane];
    }
  }
}

// ---- bdyval, south (side 0) / north (side 1): corners included  :1771-1897 ----
// One call = one (j,k) of the boundary row; j runs over jde1:jde2.  Runs after
// the west/east pass: the corner cells read what that pass stored at (jce1,ici1).
MB_HD void bdyval_sn_cell(const BdyArgs& a, int side, int j, int k) {
  const Geo& g = a.g;
  if (side == 0 ? !g.bb : !g.bt) return;
  const int id = side ? g.ide2 : g.ide1, ic = side ? g.ice2 : g.ice1, iin = side ? g.ici2 : g.ici1;
  const double sgn = side ? -1.0 : 1.0;
  const long long idu = gidx(g, j, ic, k);
  const long long sp = (long long)g.kz * g.plane;
  a.u[idu] = lin2(a.dub0, a.dub1, idu, a.x0, a.x1);   // j = jde1:jde2
  if (j > g.jce2) return;
  const long long idv = gidx(g, j, id, k), idc = idu, idin = gidx(g, j, iin, k);
  if (k == 1) { const long long i2 = gidx2(g, j, ic); a.ps[i2] = a.x0 * a.xpsb0[i2] + a.x1 * a.xpsb1[i2]; }
  const double vb = lin2(a.dvb0, a.dvb1, idv, a.x0, a.x1);
  a.v[idv] = vb;
  a.t[idc] = lin2(a.xtb0, a.xtb1, idc, a.x0, a.x1);
  a.pai[idc] = lin2(a.xpaib0, a.xpaib1, idc, a.x0, a.x1);
  a.qx[idc] = lin2(a.xqb0, a.xqb1, idc, a.x0, a.x1);
  if (a.present_qc) a.qx[idc + sp] = lin2(a.xlb0, a.xlb1, idc, a.x0, a.x1);
  if (a.present_qi && g.ipptls > 1) a.qx[idc + 2 * sp] = lin2(a.xib0, a.xib1, idc, a.x0, a.x1);
  const bool inflow = sgn * vb > 0.0;
  for (int n = a.iqfrst; n <= g.nqx; ++n) {
    if ((a.present_qc && n == 2) || (a.present_qi && n == 3)) continue;
    const double qxint = a.qx[idin + (n - 1) * sp];
    a.qx[idc + (n - 1) * sp] = inflow ? qx_zeroval(n - 1) : qxint;
  }
  a.w[idc] = inflow ? 0.0 : a.w[idin];
  if (a.tke_on) {
    if (k == 1) a.tke[idc] = a.tkemin;
    else {
      const double vkm1 = lin2(a.dvb0, a.dvb1, idv - g.plane, a.x0, a.x1);
      a.tke[idc + g.plane] = (sgn * (vb + vkm1) > 0.0) ? a.tkemin : a.tke[idin + g.plane];
    }
  }
}

// ---- chem_bdyval_uncoupled  (Main/chemlib/mod_che_bdyco.F90:391-535) -----------
// west/east pass: one (i,k,n); i runs over ice1:ice2
MB_HD void chem_bdyval_we_cell(const BdyArgs& a, int side, int i, int k, int n) {
  const Geo& g = a.g;
  if (side == 0 ? !g.bl : !g.br) return;
  double* c = a.trac + (long long)n * g.kz * g.plane;
  const int jc = side ? g.jce2 : g.jce1;
  const long long idc = gidx(g, jc, i, k);
  if (a.ichebdy == 0) {
    const int jd = side ? g.jde2 : g.jde1, jdi = side ? g.jdi2 : g.jdi1, jin = side ? g.jci2 : g.jci1;
    const double trint = c[gidx(g, jin, i, k)];
    const double windavg = a.u[gidx(g, jd, i, k)] - a.u[gidx(g, jdi, i, k)];
    const bool out = side ? (windavg > 0.0) : (windavg < 0.0);
    c[idc] = out ? trint : 0.0;
  } else if (i >= g.ici1 && i <= g.ici2) {
    const long long off = (long long)n * g.kz * g.plane + idc;
    c[idc] = a.xc0 * a.chib0[off] + a.xc1 * a.chib1[off];
  }
}
// south/north pass: one (j,k,n); j runs over jce1:jce2
MB_HD void chem_bdyval_sn_cell(const BdyArgs& a, int side, int j, int k, int n) {
  const Geo& g = a.g;
  if (side == 0 ? !g.bb : !g.bt) return;
  double* c = a.trac + (long long)n * g.kz * g.plane;
  const int ic = side ? g.ice2 : g.ice1;
  const long long idc = gidx(g, j, ic, k);
  if (a.ichebdy == 0) {
    if (j < g.jci1 || j > g.jci2) return;
    const int id = side ? g.ide2 : g.ide1, idi = side ? g.idi2 : g.idi1, iin = side ? g.ici2 : g.ici1;
    const double trint = c[gidx(g, j, iin, k)];
    const double windavg = a.v[gidx(g, j, id, k)] - a.v[gidx(g, j, idi, k)];
    const bool out = side ? (windavg > 0.0) : (windavg < 0.0);
    c[idc] = out ? trint : 0.0;
  } else {
    const long long off = (long long)n * g.kz * g.plane + idc;
    c[idc] = a.xc0 * a.chib0[off] + a.xc1 * a.chib1[off];
  }
}

// ---- motopnudge + morelax_* of every variable, one (j,i,k) of the dot box -----
// The reference applies motopnudge (t,u,v on k <= nztop, w on level 2) and then
// the Davies relaxation variable by variable; every update is pointwise, so
// one pass per cell in the same order gives the same bits.  x0/x1 here are the
// weights AFTER bdyval advanced xbctime (:2653 precedes :4058 and :4016).
MB_HD double relax_to(double f, double xf, double fext) { return (1.0 - xf) * f + xf * fext; }
MB_HD void bdy_relax_cell(const BdyArgs& a, int j, int i, int k) {
  const Geo& g = a.g;
  const long long id = gidx(g, j, i, k), i2 = gidx2(g, j, i);
  const long long sp = (long long)g.kz * g.plane;
  const double trtau = 1.0 / (2.0 * 3600.0), wrtau = 1.0 / (3600.0 / 4.0);
  const bool topk = a.top_nudge && k <= a.nztop;
  const double xft = topk ? a.tnudge[k - 1] * trtau * a.dtsec : 0.0;
  const bool sponge = a.nspgx > 0;
  const bool in_jc = j >= g.jci1 && j <= g.jci2, in_ic = i >= g.ici1 && i <= g.ici2;
  // Only the sponge ring (ibnd > 0) and the nztop nudged layers are touched: every
  // other cell is left alone (no load, no store), which is most of the grid.
  if (in_ic && j >= g.jdi1 && j <= g.jdi2) {          // u: motopnudge :4066-4070, morelax(ba_ud) [F90:478]
    const int ib = sponge ? a.ib_ud[i2] : 0;
    if (topk || ib > 0) {
      double f = a.u[id];
      const double fext = lin2(a.dub0, a.dub1, id, a.x0, a.x1);
      if (topk) f = relax_to(f, xft, fext);
      if (ib > 0) f = relax_to(f, a.hefc[(k - 1) * a.nspgx + (ib - 1)], fext);
      a.u[id] = f;
    }
  }
  if (in_jc && i >= g.idi1 && i <= g.idi2) {          // v: :4071-4075, morelax(ba_vd) [F90:479]
    const int ib = sponge ? a.ib_vd[i2] : 0;
    if (topk || ib > 0) {
      double f = a.v[id];
      const double fext = lin2(a.dvb0, a.dvb1, id, a.x0, a.x1);
      if (topk) f = relax_to(f, xft, fext);
      if (ib > 0) f = relax_to(f, a.hefc[(k - 1) * a.nspgx + (ib - 1)], fext);
      a.v[id] = f;
    }
  }
  if (!(in_jc && in_ic)) return;
  const int ib = sponge ? a.ib_cr[i2] : 0;
  const double xf = ib > 0 ? a.hefc[(k - 1) * a.nspgx + (ib - 1)] : 0.0;
  if (topk || ib > 0) {                               // t: :4061-4065, morelax [F90:480]
    double f = a.t[id];
    const double fext = lin2(a.xtb0, a.xtb1, id, a.x0, a.x1);
    if (topk) f = relax_to(f, xft, fext);
    if (ib > 0) f = relax_to(f, xf, fext);
    a.t[id] = f;
  }
  if ((a.top_nudge && k == 2) || ib > 0) {            // w: :4056-4060 (wfac = 0), morelax_fraction(frac = 0) [F90:483]
    double f = a.w[id];
    if (a.top_nudge && k == 2) { const double fext = 0.0 * f; const double xw = wrtau * a.dtsec; f = (1.0 - xw) * f + xw * fext; }
    if (ib > 0) f = (1.0 - xf) * f + xf * f * 0.0;
    a.w[id] = f;
  }
  if (ib > 0) {
    a.pai[id] = relax_to(a.pai[id], xf, lin2(a.xpaib0, a.xpaib1, id, a.x0, a.x1));       // [F90:481]
    a.qx[id] = relax_to(a.qx[id], xf, lin2(a.xqb0, a.xqb1, id, a.x0, a.x1));             // [F90:482]
    if (g.ipptls > 0) {
      if (a.present_qc) a.qx[id + sp] = relax_to(a.qx[id + sp], xf, lin2(a.xlb0, a.xlb1, id, a.x0, a.x1));
      if (g.ipptls > 1 && a.present_qi)
        a.qx[id + 2 * sp] = relax_to(a.qx[id + 2 * sp], xf, lin2(a.xib0, a.xib1, id, a.x0, a.x1));
    }
    if (a.ichem && a.ichebdy != 0) {                  // morelax_chiten (mod_che_bdyco.F90:965-1026)
      const double xc = a.fcx[ib - 1];
      for (int n = 0; n < g.ntr; ++n) {
        const long long off = id + n * sp;
        a.trac[off] = relax_to(a.trac[off], xc, a.xc0 * a.chib0[off] + a.xc1 * a.chib1[off]);
      }
    }
  }
}

// moist factor of temp_to_tvirt [F90:1608-1627]
MB_HD double moist_factor_hd(const Geo& g, const double* qx, long long id) {
  const long long sp = (long long)g.kz * g.plane;
  if (g.ipptls > 0) {
    if (g.ipptls > 1) return 1.0 + ep1 * qx[id] - qx[id + sp] - qx[id + 2 * sp] - qx[id + 3 * sp] - qx[id + 4 * sp];
    return 1.0 + ep1 * qx[id] - qx[id + sp];
  }
  return 1.0 + ep1 * qx[id];
}

// ---- end of `boundary`: uvstagtouvx, temp_to_tvirt, tetav  [F90:521-527] --------
// one (j,i,k) of the cross box, after the width-2 exchange of u (lr) and v (bt)
MB_HD void bdy_finish_cell(const BdyArgs& a, int j, int i, int k) {
  const Geo& g = a.g;
  const long long id = gidx(g, j, i, k);
  const double* u = a.u; const double* v = a.v;
  if (j >= g.jci1 && j <= g.jci2) a.ux[id] = 0.5625 * (u[id + 1] + u[id]) - 0.0625 * (u[id + 2] + u[id - 1]);
  else if (j == g.jce1) a.ux[id] = 0.5 * (u[id] + u[id + 1]);       // has_bdyleft: u(jde1), u(jdi1)
  else a.ux[id] = 0.5 * (u[id + 1] + u[id]);                        // has_bdyright: u(jde2), u(jdi2)
  if (i >= g.ici1 && i <= g.ici2) a.vx[id] = 0.5625 * (v[id + g.NJ] + v[id]) - 0.0625 * (v[id + 2 * g.NJ] + v[id - g.NJ]);
  else if (i == g.ice1) a.vx[id] = 0.5 * (v[id] + v[id + g.NJ]);
  else a.vx[id] = 0.5 * (v[id + g.NJ] + v[id]);
  const double tv = a.t[id] * moist_factor_hd(g, a.qx, id);
  a.tvirt[id] = tv;
  a.tetav[id] = tv / a.pai[id];
}

// ---- mospectral_nudge + lowpass_filter  (Main/mod_bdycod.F90:3898-3960) -------------
// Single rank: row_reduce/column_reduce are copies of the first `count`
// elements of the contiguous sx(ide1:ide2,1:2km) / sy(jde1:jde2,1:2lm) arrays
// (count = nk*(i2-i1+1), mod_mppparam.F90:20620-20664).  Where i1:i2 (j1:j2) is
// shorter than the allocation, the last elements of sxg (syg) are NOT updated
// and keep what the last full-length call left there; `stale` carries those
// values (see k_spectral_nudge).  The levels of one variable are independent
// and are processed in parallel; all sums run in the reference's order.
struct SpecArgs {
  Geo g;
  double* f;                 // t, u or v
  const double *b0, *b1;     // its boundary buffers
  double *zn, *g1;           // 3-D scratch: the departure and its zonally filtered version
  double *sx, *sy;           // [kz][2km][ni], [kz][2lm][nj]
  const double *sx_stale, *sy_stale;   // sxg / syg as the last full-length call left them (absolute index)
  const double *bvx, *bvy;   // bvx[(kk-1)*nj + (j-jde1)], bvy[(l-1)*ni + (i-ide1)]
  const double* cnudge;
  double x0, x1;
  int km2, lm2, ni, nj;
  int j1, j2, i1, i2, jj1, jj2, ii1, ii2;
  int count_x, count_y;      // reduced elements of sx / sy
};
// the three calls of `boundary` [F90:499-506]; NB the reference passes jci1,jci1
// as the j update range of t
MB_HD void spec_ranges(const Geo& g, int var, SpecArgs& a) {
  if (var == 0) { a.j1 = g.jce1; a.j2 = g.jce2; a.i1 = g.ice1; a.i2 = g.ice2; a.jj1 = g.jci1; a.jj2 = g.jci1; a.ii1 = g.ici1; a.ii2 = g.ici2; }
  else if (var == 1) { a.j1 = g.jde1; a.j2 = g.jde2; a.i1 = g.ice1; a.i2 = g.ice2; a.jj1 = g.jdi1; a.jj2 = g.jdi2; a.ii1 = g.ici1; a.ii2 = g.ici2; }
  else { a.j1 = g.jce1; a.j2 = g.jce2; a.i1 = g.ide1; a.i2 = g.ide2; a.jj1 = g.jci1; a.jj2 = g.jci2; a.ii1 = g.idi1; a.ii2 = g.idi2; }
  a.count_x = a.km2 * (a.i2 - a.i1 + 1);
  a.count_y = a.lm2 * (a.j2 - a.j1 + 1);
}
// zn1 = (x0*b0 + x1*b1) - f on j1:j2, i1:i2
MB_HD void spec_zn_cell(const SpecArgs& a, int j, int i, int k) {
  const long long id = gidx(a.g, j, i, k);
  a.zn[id] = (a.x0 * a.b0[id] + a.x1 * a.b1[id]) - a.f[id];
}
// sx(i,kk) = sum_{j=jj1..jj2} zn1(j,i)*bvx(j,kk)
MB_HD void spec_sx_cell(const SpecArgs& a, int i, int kk, int k) {
  const Geo& g = a.g;
  double acc = 0.0;
  const long long row = gidx(g, 0, i, k);
  for (int j = a.jj1; j <= a.jj2; ++j) acc = acc + a.zn[row + j] * a.bvx[(kk - 1) * a.nj + (j - g.jde1)];
  a.sx[((long long)(k - 1) * a.km2 + (kk - 1)) * a.ni + (i - g.ide1)] = acc;
}
MB_HD double spec_sxg(const SpecArgs& a, int i, int kk, int k) {
  const int e = (kk - 1) * a.ni + (i - a.g.ide1);
  return e < a.count_x ? a.sx[(long long)(k - 1) * a.km2 * a.ni + e] : a.sx_stale[e];
}
MB_HD double spec_syg(const SpecArgs& a, int j, int l, int k) {
  const int e = (l - 1) * a.nj + (j - a.g.jde1);
  return e < a.count_y ? a.sy[(long long)(k - 1) * a.lm2 * a.nj + e] : a.sy_stale[e];
}
// f(j,i) = sum_kk sxg(i,kk)*bvx(j,kk) on j1:j2, i1:i2 (zonal reconstruction)
MB_HD void spec_g1_cell(const SpecArgs& a, int j, int i, int k) {
  double acc = 0.0;
  for (int kk = 1; kk <= a.km2; ++kk) acc = acc + spec_sxg(a, i, kk, k) * a.bvx[(kk - 1) * a.nj + (j - a.g.jde1)];
  a.g1[gidx(a.g, j, i, k)] = acc;
}
// sy(j,l) = sum_{i=ii1..ii2} f(j,i)*bvy(i,l)
MB_HD void spec_sy_cell(const SpecArgs& a, int j, int l, int k) {
  const Geo& g = a.g;
  double acc = 0.0;
  for (int i = a.ii1; i <= a.ii2; ++i) acc = acc + a.g1[gidx(g, j, i, k)] * a.bvy[(l - 1) * a.ni + (i - g.ide1)];
  a.sy[((long long)(k - 1) * a.lm2 + (l - 1)) * a.nj + (j - g.jde1)] = acc;
}
// f(j,i,k) += cnudge(k) * sum_l syg(j,l)*bvy(i,l) on jj1:jj2, ii1:ii2
MB_HD void spec_update_cell(const SpecArgs& a, int j, int i, int k) {
  double acc = 0.0;
  for (int l = 1; l <= a.lm2; ++l) acc = acc + spec_syg(a, j, l, k) * a.bvy[(l - 1) * a.ni + (i - a.g.ide1)];
  const long long id = gidx(a.g, j, i, k);
  a.f[id] = a.f[id] + a.cnudge[k - 1] * acc;
}

// ---- mkslice, idynamic == 3  (Main/mod_slice.F90:115-173) ------------------------
struct SliceArgs {
  Geo g;
  const double *pai, *t, *p, *rho, *qsat, *w, *ps, *zq;
  double *qx, *trac, *pf3d, *th3d, *rhb3d, *wpx3d, *rhox2d, *tp2d, *th700;
  const double *xlat, *za;
  double *ptrop, *ktrop, *kmxpbl;
  double rhmin, rhmax, calday, dayspy;
  int ichem, icldmstrat, irceideal, ibltyp;
};
constexpr double rovcp_hd = rgas * (1.0 / cpd);
// one (j,i,k) of the cross box, k = 1..kz
MB_HD void mkslice_cell(const SliceArgs& a, int j, int i, int k) {
  const Geo& g = a.g;
  const long long id = gidx(g, j, i, k);
  const long long sp = (long long)g.kz * g.plane;
  if (k >= 2) a.pf3d[id] = p00 * pow(0.5 * (a.pai[id] + a.pai[id - g.plane]), cpovr);
  a.th3d[id] = a.t[id] * pow(p00 / a.p[id], rovcp_hd);
  if (!(j >= g.jci1 && j <= g.jci2 && i >= g.ici1 && i <= g.ici2)) return;
  for (int n = 0; n < g.nqx; ++n)
    if (a.qx[id + n * sp] < qx_checkval(n)) a.qx[id + n * sp] = qx_zeroval(n);
  if (a.ichem)
    for (int n = 0; n < g.ntr; ++n)
      if (a.trac[id + n * sp] < 1.0e-50) a.trac[id + n * sp] = 0.0;
  a.rhb3d[id] = fmin(fmax(a.qx[id] / a.qsat[id], a.rhmin), a.rhmax);
  a.wpx3d[id] = -egrav * a.rho[id] * 0.5 * (a.w[id + g.plane] + a.w[id]);
}
// one (j,i) column of the cross box, after mkslice_cell of every level
MB_HD void mkslice_col(const SliceArgs& a, int j, int i) {
  const Geo& g = a.g;
  const int kz = g.kz;
  const long long i2 = gidx2(g, j, i), pl = g.plane;
  const long long id1 = gidx(g, j, i, 1), idz = gidx(g, j, i, kz);
  a.pf3d[idz + pl] = a.ps[i2];
  a.pf3d[id1] = a.pf3d[id1 + pl] - egrav * a.rho[id1] * (a.zq[id1] - a.zq[id1 + pl]);
  if (!(j >= g.jci1 && j <= g.jci2 && i >= g.ici1 && i <= g.ici2)) return;
  a.rhox2d[i2] = a.ps[i2] / (rgas * a.t[idz]);
  a.tp2d[i2] = a.t[idz] * pow(a.ps[i2] / a.p[idz], rovcp_hd);
  if (a.icldmstrat == 1) {
    double th = a.th3d[idz];
    for (int k = 2; k <= kz - 1; ++k) {
      const long long id = id1 + (long long)(k - 1) * pl;
      if (a.p[id] > 70000.0) {
        const double w1 = (a.p[id] - 70000.0) / (a.p[id] - a.p[id - pl]);
        const double w2 = 1.0 - w1;
        th = a.th3d[id - pl] * w1 + a.th3d[id] * w2;
        break;
      }
    }
    a.th700[i2] = th;
  }
}

// common tail of mkslice (Main/mod_slice.F90:342-384): tropopause pressure (Mateus, Mendes, Pires 2022),
// its level index and the highest level the PBL may reach; one interior (j,i) column
MB_HD void mkslice_trop_col(const SliceArgs& a, int j, int i) {
  const Geo& g = a.g;
  if (!(j >= g.jci1 && j <= g.jci2 && i >= g.ici1 && i <= g.ici2)) return;
  const double twopi = 3.14159265358979323846 * 2.0;      // Share/mod_constants.F90:305,323
  const double anorth[6] = {7.9925, 8.3329, 24.1731, -1.8069, 0.1082, -0.1493};   // mod_slice.F90:42-48
  const double asouth[6] = {8.1797, 8.1455, -23.4839, 1.1464, 0.0798, -0.1491};
  const int kz = g.kz;
  const long long i2 = gidx2(g, j, i), pl = g.plane, id1 = gidx(g, j, i, 1);
  if (a.irceideal != 1) {
    const double xl = a.xlat[i2];
    const double* c = xl > 0.0 ? anorth : asouth;
    const double ztrop = c[0] + c[1] / pow(1.0 + exp(-(xl - c[2]) / c[3]), c[4]) +
                         c[5] * cos((twopi * (a.calday - 28.0)) / a.dayspy);
    a.ptrop[i2] = p00 * exp(-ztrop / 8.4);
  }
  const double pt = a.ptrop[i2];
  int kt = kz;
  for (int k = kz - 1; k >= 2; --k) {
    kt = k;
    if (a.p[id1 + (long long)(k - 1) * pl] < pt) break;
  }
  a.ktrop[i2] = (double)kt;
  if (a.ibltyp == 1) {
    int km = kz;
    for (int k = kz - 1; k >= 2; --k) {
      if (a.za[id1 + (long long)(k - 1) * pl] > 5000.0) break;     // Saharan heat lows 5-6 km
      km = k;
    }
    a.kmxpbl[i2] = (double)km;
  }
}

// ---- massck, idynamic == 3  (Main/mod_massck.F90:77-175) and the ps guard ---------
struct MassArgs {
  Geo g;
  const double *rho, *zq, *qx, *u, *v, *ps;
  double *rows;     // [2][kz*ni]: row sums of dry mass and water mass
  double *lev;      // [4][kz]: per level: boundary flux sums (dry, water), interior sums (dry, water)
  double *psrow;    // [3][ni]: row max, row min, non-finite count of ps
  double *out;      // tdrym, tdadv, tqmass, tqadv, psmax, psmin, nonfinite
  double dxsq, dt, dx;
  int ni;
};
// dz = zetaf(k) - zetaf(k+1)  (Main/mod_params.F90:3389)
MB_HD double mass_dz(const MassArgs& a, long long id) { return a.zq[id] - a.zq[id + a.g.plane]; }
// one interior row (i,k): sum over j = jci1:jci2 in the reference's order (:79-83, :139-146)
MB_HD void massck_row(const MassArgs& a, int i, int k) {
  const Geo& g = a.g;
  const long long sp = (long long)g.kz * g.plane;
  double dry = 0.0, wat = 0.0;
  if (i >= g.ici1 && i <= g.ici2) {
    for (int j = g.jci1; j <= g.jci2; ++j) {
      const long long id = gidx(g, j, i, k);
      dry = dry + a.dxsq * mass_dz(a, id) * a.rho[id];
    }
    for (int n = 0; n < g.nqx; ++n)
      for (int j = g.jci1; j <= g.jci2; ++j) {
        const long long id = gidx(g, j, i, k);
        wat = wat + a.qx[id + n * sp] * a.dxsq * mass_dz(a, id) * a.rho[id];
      }
  }
  const long long r = (long long)(k - 1) * a.ni + (i - g.ice1);
  a.rows[r] = dry;
  a.rows[(long long)g.kz * a.ni + r] = wat;
}
// one level: the level's row sums added in row order, and its boundary fluxes (:85-118, :150-185)
MB_HD void massck_bdy_level(const MassArgs& a, int k) {
  const Geo& g = a.g;
  const long long sp = (long long)g.kz * g.plane;
  const long long nr = (long long)g.kz * a.ni;
  double rd = 0.0, rw = 0.0;
  for (int r = 0; r < a.ni; ++r) {
    rd = rd + a.rows[(long long)(k - 1) * a.ni + r];
    rw = rw + a.rows[nr + (long long)(k - 1) * a.ni + r];
  }
  a.lev[2 * g.kz + k - 1] = rd;
  a.lev[3 * g.kz + k - 1] = rw;
  double dry = 0.0, wat = 0.0;
  for (int pass = 0; pass <= g.nqx; ++pass) {   // pass 0: dry air, pass n: water species n
    double acc = 0.0;
    const long long off = pass > 0 ? (long long)(pass - 1) * sp : 0;
    if (g.bl) for (int i = g.ice1; i <= g.ice2; ++i) {
      const long long idc = gidx(g, g.jce1, i, k), idd = gidx(g, g.jde1, i, k);
      acc = acc + (pass > 0 ? a.qx[idc + off] * a.u[idd] : a.u[idd]) * a.dt * a.dx * mass_dz(a, idc) * a.rho[idc];
    }
    if (g.br) for (int i = g.ice1; i <= g.ice2; ++i) {
      const long long idc = gidx(g, g.jce2, i, k), idd = gidx(g, g.jde2, i, k);
      acc = acc - (pass > 0 ? a.qx[idc + off] * a.u[idd] : a.u[idd]) * a.dt * a.dx * mass_dz(a, idc) * a.rho[idc];
    }
    if (g.bb) for (int j = g.jci1; j <= g.jci2; ++j) {
      const long long idc = gidx(g, j, g.ice1, k), idd = gidx(g, j, g.ide1, k);
      acc = acc + (pass > 0 ? a.qx[idc + off] * a.v[idd] : a.v[idd]) * a.dt * a.dx * mass_dz(a, idc) * a.rho[idc];
    }
    if (g.bt) for (int j = g.jci1; j <= g.jci2; ++j) {
      const long long idc = gidx(g, j, g.ice2, k), idd = gidx(g, j, g.ide2, k);
      acc = acc - (pass > 0 ? a.qx[idc + off] * a.v[idd] : a.v[idd]) * a.dt * a.dx * mass_dz(a, idc) * a.rho[idc];
    }
    if (pass == 0) dry = acc; else wat = wat + acc;
  }
  a.lev[k - 1] = dry;
  a.lev[g.kz + k - 1] = wat;
}
// ps over one interior row [F90:408-409]
MB_HD void ps_row(const MassArgs& a, int i) {
  const Geo& g = a.g;
  double mx = -1.0e300, mn = 1.0e300, bad = 0.0;
  if (i >= g.ici1 && i <= g.ici2)
    for (int j = g.jci1; j <= g.jci2; ++j) {
      const double p = a.ps[gidx2(g, j, i)];
      if (!(p - p == 0.0)) bad = bad + 1.0;      // NaN or Inf
      else { if (p > mx) mx = p; if (p < mn) mn = p; }
    }
  a.psrow[i - g.ice1] = mx; a.psrow[a.ni + i - g.ice1] = mn; a.psrow[2 * a.ni + i - g.ice1] = bad;
}
// final sums, one thread: the kz level sums
MB_HD void massck_final(const MassArgs& a) {
  const Geo& g = a.g;
  double dry = 0.0, wat = 0.0, fd = 0.0, fw = 0.0;
  for (int k = 0; k < g.kz; ++k) {
    fd = fd + a.lev[k]; fw = fw + a.lev[g.kz + k];
    dry = dry + a.lev[2 * g.kz + k]; wat = wat + a.lev[3 * g.kz + k];
  }
  double mx = -1.0e300, mn = 1.0e300, bad = 0.0;
  for (int i = 0; i < a.ni; ++i) {
    if (a.psrow[i] > mx) mx = a.psrow[i];
    if (a.psrow[a.ni + i] < mn) mn = a.psrow[a.ni + i];
    bad = bad + a.psrow[2 * a.ni + i];
  }
  a.out[0] = dry; a.out[1] = fd; a.out[2] = wat; a.out[3] = fw; a.out[4] = mx; a.out[5] = mn; a.out[6] = bad;
}

// ---- tendency diagnostics (idiag, ichdiag)  [F90:1092-1103, 1127-1139, 455-466, 508-519] ----------
struct DiagArgs {
  Geo g;
  const double *t, *qv, *trac;
  double *ten0, *qen0, *chiten0;      // snapshots
  double *dt_out, *dq_out, *dc_out;   // tdiag%adh|bdy, qdiag%adh|bdy, cadvhdiag|cbdydiag
  double rdt;
  int idiag, ichdiag;
};
// ten0 = t, qen0 = qv, chiten0 = trac on the interior
MB_HD void diag_snap_cell(const DiagArgs& a, int j, int i, int k) {
  const long long id = gidx(a.g, j, i, k), sp = (long long)a.g.kz * a.g.plane;
  if (a.idiag) { a.ten0[id] = a.t[id]; a.qen0[id] = a.qv[id]; }
  if (a.ichdiag) for (int n = 0; n < a.g.ntr; ++n) a.chiten0[id + n * sp] = a.trac[id + n * sp];
}
// (x - x0) * rdt
MB_HD void diag_diff_cell(const DiagArgs& a, int j, int i, int k) {
  const long long id = gidx(a.g, j, i, k), sp = (long long)a.g.kz * a.g.plane;
  if (a.idiag) {
    a.dt_out[id] = (a.t[id] - a.ten0[id]) * a.rdt;
    a.dq_out[id] = (a.qv[id] - a.qen0[id]) * a.rdt;
  }
  if (a.ichdiag) for (int n = 0; n < a.g.ntr; ++n) a.dc_out[id + n * sp] = (a.trac[id + n * sp] - a.chiten0[id + n * sp]) * a.rdt;
}

// ---- TKE (ibltyp == 2) -------------------------------------------------------------
// zstagtoh(tke,tkex) [F90:1445-1459]: one (j,i,k), k = 1..kz
MB_HD void zstagtoh_cell(const Geo& g, const double* fl, double* hl, int j, int i, int k) {
  const long long id = gidx(g, j, i, k), pl = g.plane;
  if (k == 1 || k == g.kz) hl[id] = 0.5 * (fl[id + pl] + fl[id]);
  else hl[id] = 0.5625 * (fl[id + pl] + fl[id]) - 0.0625 * (fl[id + 2 * pl] + fl[id - pl]);
}
// htozstag(tkex,tke) [F90:1461-1475]: one (j,i,k), k = 2..kz
MB_HD void htozstag_cell(const Geo& g, const double* hl, double* fl, int j, int i, int k) {
  const long long id = gidx(g, j, i, k), pl = g.plane;
  if (k == 2 || k == g.kz) fl[id] = 0.5 * (hl[id] + hl[id - pl]);
  else fl[id] = 0.5625 * (hl[id] + hl[id - pl]) - 0.0625 * (hl[id + pl] + hl[id - 2 * pl]);
}
// status_update of tke [F90:1419-1424]: one (j,i,k) of the interior, k = 1..kz+1
MB_HD void tke_update_cell(const Geo& g, double* tke, const double* tketen, double dtinc, double tkemin, int j,
                           int i, int k) {
  const long long id = gidx(g, j, i, k);
  double x = tke[id] + dtinc * tketen[id];
  if (x < tkemin) x = tkemin;
  tke[id] = x;
}

}  // namespace mb
