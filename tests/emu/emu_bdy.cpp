// emu_bdy.cpp -- TEST INFRASTRUCTURE: host-compiled instantiation of the cell
// functions of regcm_b200/csrc/bdy_cells.h.
//
// The GPU kernels of kernels_bdy.cu are one-line wrappers that map a CUDA
// thread onto one call of a cell function.  This file includes exactly the same
// header with g++ (-ffp-contract=off) and calls the cell functions over the same
// index ranges as the launch grids, in the product's launch order, on host
// arrays with the product's padded layout.  The CPU test-suite compares the
// result bit for bit with the oracle, so index ranges, branches and operation
// order of the device code are checked here, without a GPU; `order` = 1 walks
// every grid backwards, which exposes any dependence between the threads of a
// launch.  It is NOT a CPU fallback: nothing under regcm_b200/ builds, loads or
// calls it.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "bdy_cells.h"

using namespace mb;

namespace {
thread_local std::string g_err;
int fail(const std::string& m) { g_err = m; return 1; }

struct Emu {
  moloch_b200_config cfg;
  Geo g;
  std::vector<double> f[MB_NFIELDS];
  int nk[MB_NFIELDS], nspec[MB_NFIELDS];
  std::vector<int> ibnd[3];
  std::vector<double> tab[MB_NTABLES];
  double xbctime = 0.0, tspectral = 0.0, calday = 1.0, dayspy = 365.2422;
  int order = 0;
  std::vector<double> zn, g1, sx, sy, sx_stale, sy_stale;   // mospectral_nudge scratch
};

double* P(Emu& e, int id) { return e.f[id].empty() ? nullptr : e.f[id].data(); }

BdyArgs bdy_args(Emu& e, double xbctime) {
  BdyArgs a;
  std::memset(&a, 0, sizeof(a));
  const moloch_b200_config& c = e.cfg;
  a.g = e.g;
  a.u = P(e, MB_U); a.v = P(e, MB_V); a.w = P(e, MB_W); a.t = P(e, MB_T); a.pai = P(e, MB_PAI);
  a.qx = P(e, MB_QX); a.trac = P(e, MB_TRAC); a.ps = P(e, MB_PS); a.tke = P(e, MB_TKE);
  a.ux = P(e, MB_UX); a.vx = P(e, MB_VX); a.tvirt = P(e, MB_TVIRT); a.tetav = P(e, MB_TETAV);
  a.dub0 = P(e, MB_DUB0); a.dub1 = P(e, MB_DUB1); a.dvb0 = P(e, MB_DVB0); a.dvb1 = P(e, MB_DVB1);
  a.xtb0 = P(e, MB_XTB0); a.xtb1 = P(e, MB_XTB1); a.xpaib0 = P(e, MB_XPAIB0); a.xpaib1 = P(e, MB_XPAIB1);
  a.xqb0 = P(e, MB_XQB0); a.xqb1 = P(e, MB_XQB1); a.xlb0 = P(e, MB_XLB0); a.xlb1 = P(e, MB_XLB1);
  a.xib0 = P(e, MB_XIB0); a.xib1 = P(e, MB_XIB1); a.xpsb0 = P(e, MB_XPSB0); a.xpsb1 = P(e, MB_XPSB1);
  a.chib0 = P(e, MB_CHIB0); a.chib1 = P(e, MB_CHIB1);
  a.ib_cr = e.ibnd[0].empty() ? nullptr : e.ibnd[0].data();
  a.ib_ud = e.ibnd[1].empty() ? nullptr : e.ibnd[1].data();
  a.ib_vd = e.ibnd[2].empty() ? nullptr : e.ibnd[2].data();
  a.hefc = e.tab[MB_TAB_HEFC].data(); a.tnudge = e.tab[MB_TAB_TNUDGE].data(); a.fcx = e.tab[MB_TAB_FCX].data();
  const double rtb = 1.0 / c.dtbdys;
  a.x1 = (xbctime + c.dtsec) * rtb; a.x0 = 1.0 - a.x1;
  a.xc1 = (xbctime + c.dtsec) / c.dtbdys; a.xc0 = 1.0 - a.xc1;
  a.dtsec = c.dtsec; a.tkemin = c.tkemin;
  a.nspgx = c.nspgx; a.iqfrst = c.iqfrst; a.present_qc = c.present_qc; a.present_qi = c.present_qi;
  a.tke_on = c.ibltyp == 2; a.nztop = c.nztop; a.top_nudge = c.mo_top_nudge;
  a.ichem = c.ichem && c.ntr > 0; a.ichebdy = c.ichebdy;
  return a;
}

// walk lo..hi forwards (order 0) or backwards (order 1)
template <class F> void walk(int order, int lo, int hi, F f) {
  if (order == 0) for (int x = lo; x <= hi; ++x) f(x);
  else for (int x = hi; x >= lo; --x) f(x);
}

// local periodic wrap of a single rank (what halo_exchange does for a rank that
// is its own neighbour): width-2 ghosts of u (left/right) and v (bottom/top)
void wrap_uv(Emu& e) {
  const Geo& g = e.g; const int kz = g.kz;
  double* u = P(e, MB_U); double* v = P(e, MB_V);
  if (e.cfg.bandflag && e.cfg.nbr_left == e.cfg.rank) {
    const int nj = g.jde2 - g.jde1 + 1;
    for (int k = 1; k <= kz; ++k) for (int i = g.ice1; i <= g.ice2; ++i) for (int x = 1; x <= 2; ++x) {
      u[gidx(g, g.jde1 - x, i, k)] = u[gidx(g, g.jde1 - x + nj, i, k)];
      u[gidx(g, g.jde2 + x, i, k)] = u[gidx(g, g.jde2 + x - nj, i, k)];
    }
  }
  if (e.cfg.crmflag && e.cfg.nbr_bottom == e.cfg.rank) {
    const int ni = g.ide2 - g.ide1 + 1;
    for (int k = 1; k <= kz; ++k) for (int x = 1; x <= 2; ++x) for (int j = g.jce1; j <= g.jce2; ++j) {
      v[gidx(g, j, g.ide1 - x, k)] = v[gidx(g, j, g.ide1 - x + ni, k)];
      v[gidx(g, j, g.ide2 + x, k)] = v[gidx(g, j, g.ide2 + x - ni, k)];
    }
  }
}

int do_diag(Emu& e, int which, bool diff) {   // = k_diag
  const Geo& g = e.g; const int o = e.order;
  DiagArgs a;
  std::memset(&a, 0, sizeof(a));
  a.g = g;
  a.idiag = e.cfg.idiag > 0; a.ichdiag = e.cfg.ichdiag > 0 && e.cfg.ichem && e.cfg.ntr > 0;
  if (!a.idiag && !a.ichdiag) return 0;
  a.t = P(e, MB_T); a.qv = P(e, MB_QX); a.trac = P(e, MB_TRAC);
  a.ten0 = P(e, MB_TEN0); a.qen0 = P(e, MB_QEN0); a.chiten0 = P(e, MB_CHITEN0);
  a.dt_out = P(e, which ? MB_TDIAG_BDY : MB_TDIAG_ADH); a.dq_out = P(e, which ? MB_QDIAG_BDY : MB_QDIAG_ADH);
  a.dc_out = P(e, which ? MB_CBDYDIAG : MB_CADVHDIAG);
  a.rdt = 1.0 / e.cfg.dtsec;
  walk(o, 1, g.kz, [&](int k) { walk(o, g.ici1, g.ici2, [&](int i) { walk(o, g.jci1, g.jci2, [&](int j) {
    if (diff) diag_diff_cell(a, j, i, k); else diag_snap_cell(a, j, i, k); }); }); });
  return 0;
}
int do_bdyval(Emu& e, double xbctime) {   // = k_bdyval
  const Geo& g = e.g; const int kz = g.kz, o = e.order;
  const BdyArgs a = bdy_args(e, xbctime);
  if (g.bl || g.br)
    walk(o, 0, 1, [&](int s) { walk(o, 1, kz, [&](int k) { walk(o, g.ide1, g.ide2, [&](int i) { bdyval_we_cell(a, s, i, k); }); }); });
  if (g.bb || g.bt)
    walk(o, 0, 1, [&](int s) { walk(o, 1, kz, [&](int k) { walk(o, g.jde1, g.jde2, [&](int j) { bdyval_sn_cell(a, s, j, k); }); }); });
  if (a.ichem) {
    if (g.bl || g.br)
      walk(o, 0, 1, [&](int s) { walk(o, 0, kz * g.ntr - 1, [&](int y) { walk(o, g.ice1, g.ice2, [&](int i) {
        chem_bdyval_we_cell(a, s, i, 1 + y % kz, y / kz); }); }); });
    if (g.bb || g.bt)
      walk(o, 0, 1, [&](int s) { walk(o, 0, kz * g.ntr - 1, [&](int y) { walk(o, g.jce1, g.jce2, [&](int j) {
        chem_bdyval_sn_cell(a, s, j, 1 + y % kz, y / kz); }); }); });
  }
  return 0;
}
int do_relax(Emu& e, double xbctime) {   // = k_bdy_relax
  const Geo& g = e.g; const int o = e.order;
  if (!(e.cfg.mo_top_nudge || e.cfg.nspgx > 0)) return 0;
  const BdyArgs a = bdy_args(e, xbctime);
  walk(o, 1, g.kz, [&](int k) { walk(o, g.ide1, g.ide2, [&](int i) { walk(o, g.jde1, g.jde2, [&](int j) { bdy_relax_cell(a, j, i, k); }); }); });
  return 0;
}
int do_finish(Emu& e) {   // = halo round + k_bdy_finish
  const Geo& g = e.g; const int o = e.order;
  wrap_uv(e);
  const BdyArgs a = bdy_args(e, 0.0);
  walk(o, 1, g.kz, [&](int k) { walk(o, g.ice1, g.ice2, [&](int i) { walk(o, g.jce1, g.jce2, [&](int j) { bdy_finish_cell(a, j, i, k); }); }); });
  return 0;
}
int do_spectral(Emu& e, double xbctime) {   // = k_spectral_nudge
  const Geo& g = e.g; const int o = e.order;
  const int kz = g.kz, km2 = 2 * e.cfg.km, lm2 = 2 * e.cfg.lm;
  const int ni = g.ide2 - g.ide1 + 1, nj = g.jde2 - g.jde1 + 1;
  if (e.sx.empty()) {
    e.zn.assign((size_t)kz * g.plane, 0.0); e.g1 = e.zn;
    e.sx.assign((size_t)kz * km2 * ni, 0.0); e.sy.assign((size_t)kz * lm2 * nj, 0.0);
    e.sx_stale.assign((size_t)km2 * ni, 0.0); e.sy_stale.assign((size_t)lm2 * nj, 0.0);
  }
  SpecArgs a;
  std::memset(&a, 0, sizeof(a));
  a.g = g;
  a.zn = e.zn.data(); a.g1 = e.g1.data(); a.sx = e.sx.data(); a.sy = e.sy.data();
  a.sx_stale = e.sx_stale.data(); a.sy_stale = e.sy_stale.data();
  a.bvx = e.tab[MB_TAB_BVX].data(); a.bvy = e.tab[MB_TAB_BVY].data(); a.cnudge = e.tab[MB_TAB_CNUDGE].data();
  const double rtb = 1.0 / e.cfg.dtbdys;
  a.x1 = (xbctime + e.cfg.dtsec) * rtb; a.x0 = 1.0 - a.x1;
  a.km2 = km2; a.lm2 = lm2; a.ni = ni; a.nj = nj;
  const int fid[3] = {MB_T, MB_U, MB_V}, b0[3] = {MB_XTB0, MB_DUB0, MB_DVB0}, b1[3] = {MB_XTB1, MB_DUB1, MB_DVB1};
  for (int var = 0; var < 3; ++var) {
    a.f = P(e, fid[var]); a.b0 = P(e, b0[var]); a.b1 = P(e, b1[var]);
    spec_ranges(g, var, a);
    walk(o, 1, kz, [&](int k) { walk(o, a.i1, a.i2, [&](int i) { walk(o, a.j1, a.j2, [&](int j) { spec_zn_cell(a, j, i, k); }); }); });
    walk(o, 1, kz, [&](int k) { walk(o, 1, km2, [&](int kk) { walk(o, a.i1, a.i2, [&](int i) { spec_sx_cell(a, i, kk, k); }); }); });
    if (a.count_x == km2 * ni) std::copy(e.sx.begin() + (size_t)(kz - 1) * km2 * ni, e.sx.begin() + (size_t)kz * km2 * ni, e.sx_stale.begin());
    walk(o, 1, kz, [&](int k) { walk(o, a.i1, a.i2, [&](int i) { walk(o, a.j1, a.j2, [&](int j) { spec_g1_cell(a, j, i, k); }); }); });
    walk(o, 1, kz, [&](int k) { walk(o, 1, lm2, [&](int l) { walk(o, a.j1, a.j2, [&](int j) { spec_sy_cell(a, j, l, k); }); }); });
    if (a.count_y == lm2 * nj) std::copy(e.sy.begin() + (size_t)(kz - 1) * lm2 * nj, e.sy.begin() + (size_t)kz * lm2 * nj, e.sy_stale.begin());
    walk(o, 1, kz, [&](int k) { walk(o, a.ii1, a.ii2, [&](int i) { walk(o, a.jj1, a.jj2, [&](int j) { spec_update_cell(a, j, i, k); }); }); });
  }
  return 0;
}
}  // namespace

extern "C" {
const char* emu_b200_last_error(void) { return g_err.c_str(); }

int emu_b200_create(const moloch_b200_config* cfg, void** out) {
  Emu* e = new Emu();
  e->cfg = *cfg;
  e->g = geo_from_cfg(*cfg);
  for (int id = 0; id < MB_NFIELDS; ++id) {
    field_shape(*cfg, id, e->nk[id], e->nspec[id]);
    if (id == MB_WZ || id == MB_P0) { e->nk[id] = 0; continue; }
    e->f[id].assign((size_t)e->g.plane * e->nk[id] * (e->nspec[id] > 0 ? e->nspec[id] : 0), 0.0);
  }
  for (int q = 0; q < MB_NTABLES; ++q) e->tab[q].assign(1, 0.0);
  *out = e;
  return 0;
}
int emu_b200_destroy(void* h) { delete (Emu*)h; return 0; }
int emu_b200_init(void*) { return 0; }
int emu_b200_set_order(void* h, int order) { ((Emu*)h)->order = order; return 0; }

static int xfer(Emu& e, int field, int n, double* host, int jlo, int jhi, int ilo, int ihi, int klo, int khi, bool put) {
  if (field < 0 || field >= MB_NFIELDS || e.f[field].empty()) return fail("emu: field not allocated");
  const Geo& g = e.g;
  const int spec = (e.nspec[field] > 1 || n > 0) ? n - 1 : 0;
  if (spec < 0 || spec >= (e.nspec[field] > 0 ? e.nspec[field] : 1)) return fail("emu: species out of range");
  double* dev = e.f[field].data() + (size_t)spec * e.nk[field] * g.plane;
  for (int k = klo; k <= khi; ++k) for (int i = ilo; i <= ihi; ++i) for (int j = jlo; j <= jhi; ++j) {
    if (j < g.j0 || j >= g.j0 + g.NJ || i < g.i0 || i >= g.i0 + g.NI || k < 1 || k > e.nk[field]) continue;
    double& h = host[((size_t)(k - klo) * (ihi - ilo + 1) + (i - ilo)) * (jhi - jlo + 1) + (j - jlo)];
    double& d = dev[gidx(g, j, i, k)];
    if (put) d = h; else h = d;
  }
  return 0;
}
int emu_b200_set_field(void* h, int field, int n, const double* host, int jlo, int jhi, int ilo, int ihi, int klo, int khi) {
  return xfer(*(Emu*)h, field, n, const_cast<double*>(host), jlo, jhi, ilo, ihi, klo, khi, true);
}
int emu_b200_get_field(void* h, int field, int n, double* host, int jlo, int jhi, int ilo, int ihi, int klo, int khi) {
  return xfer(*(Emu*)h, field, n, host, jlo, jhi, ilo, ihi, klo, khi, false);
}
int emu_b200_set_profile(void*, int, const double*, int) { return 0; }
int emu_b200_set_table(void* h, int which, const double* v, int n) {
  Emu& e = *(Emu*)h;
  if (which < 0 || which >= MB_NTABLES) return fail("emu: unknown table");
  e.tab[which].assign(v, v + n);
  return 0;
}
int emu_b200_set_ibnd(void* h, int which, const int32_t* ib, int jlo, int jhi, int ilo, int ihi) {
  Emu& e = *(Emu*)h; const Geo& g = e.g;
  e.ibnd[which].assign((size_t)g.plane, -1);
  for (int i = ilo; i <= ihi; ++i) for (int j = jlo; j <= jhi; ++j)
    e.ibnd[which][gidx2(g, j, i)] = ib[(size_t)(i - ilo) * (jhi - jlo + 1) + (j - jlo)];
  return 0;
}
int emu_b200_set_calday(void* h, double calday, double dayspy) { ((Emu*)h)->calday = calday; ((Emu*)h)->dayspy = dayspy; return 0; }
int emu_b200_set_xbctime(void* h, double t) { ((Emu*)h)->xbctime = t; return 0; }
double emu_b200_get_xbctime(void* h) { return ((Emu*)h)->xbctime; }

int emu_b200_bdyval(void* h) {
  Emu& e = *(Emu*)h;
  do_bdyval(e, e.xbctime);
  e.xbctime = e.xbctime + e.cfg.dtsec;
  return 0;
}
int emu_b200_boundary(void* h) {   // = do_boundary (capi.cu)
  Emu& e = *(Emu*)h;
  do_diag(e, 1, false);
  do_bdyval(e, e.xbctime);
  e.xbctime = e.xbctime + e.cfg.dtsec;
  do_relax(e, e.xbctime);
  if (e.cfg.mo_spectral_nudge) {
    e.tspectral = e.tspectral + e.cfg.dtsec;
    if ((int)std::fmod(e.tspectral, e.cfg.dtrad) == 0) do_spectral(e, e.xbctime);
  }
  do_diag(e, 1, true);
  return do_finish(e);
}
// the two halves of do_boundary around its u/v halo round, for runs on several ranks where the test
// moves the halos between the ranks' contexts (no mospectral_nudge: it is refused on > 1 rank)
int emu_b200_boundary_pre(void* h) {
  Emu& e = *(Emu*)h;
  do_diag(e, 1, false);
  do_bdyval(e, e.xbctime);
  e.xbctime = e.xbctime + e.cfg.dtsec;
  do_relax(e, e.xbctime);
  return do_diag(e, 1, true);
}
int emu_b200_boundary_post(void* h) { return do_finish(*(Emu*)h); }
int emu_b200_mkslice(void* h) {   // = k_mkslice
  Emu& e = *(Emu*)h; const Geo& g = e.g; const int o = e.order;
  SliceArgs a;
  std::memset(&a, 0, sizeof(a));
  a.g = g;
  a.pai = P(e, MB_PAI); a.t = P(e, MB_T); a.p = P(e, MB_P); a.rho = P(e, MB_RHO); a.qsat = P(e, MB_QSAT);
  a.w = P(e, MB_W); a.ps = P(e, MB_PS); a.zq = P(e, MB_ZETAF); a.qx = P(e, MB_QX); a.trac = P(e, MB_TRAC);
  a.pf3d = P(e, MB_PF3D); a.th3d = P(e, MB_TH3D); a.rhb3d = P(e, MB_RHB3D); a.wpx3d = P(e, MB_WPX3D);
  a.rhox2d = P(e, MB_RHOX2D); a.tp2d = P(e, MB_TP2D); a.th700 = P(e, MB_TH700);
  a.rhmin = e.cfg.rhmin; a.rhmax = e.cfg.rhmax;
  a.ichem = e.cfg.ichem && e.cfg.ntr > 0; a.icldmstrat = e.cfg.icldmstrat;
  a.xlat = P(e, MB_XLAT); a.za = P(e, MB_ZETA); a.ptrop = P(e, MB_PTROP); a.ktrop = P(e, MB_KTROP);
  a.kmxpbl = P(e, MB_KMXPBL); a.calday = e.calday; a.dayspy = e.dayspy; a.irceideal = e.cfg.irceideal;
  a.ibltyp = e.cfg.ibltyp;
  if (!a.pf3d) return fail("emu: do_slice not configured");
  walk(o, 1, g.kz, [&](int k) { walk(o, g.ice1, g.ice2, [&](int i) { walk(o, g.jce1, g.jce2, [&](int j) { mkslice_cell(a, j, i, k); }); }); });
  walk(o, g.ice1, g.ice2, [&](int i) { walk(o, g.jce1, g.jce2, [&](int j) { mkslice_col(a, j, i); mkslice_trop_col(a, j, i); }); });
  return 0;
}
// = k_massck: out7 = tdrym, tdadv, tqmass, tqadv, psmax, psmin, nonfinite
int emu_b200_massck7(void* h, int what, double* out7) {
  Emu& e = *(Emu*)h; const Geo& g = e.g; const int o = e.order;
  const int ni = g.ice2 - g.ice1 + 1, kz = g.kz;
  std::vector<double> work((size_t)2 * kz * ni + 4 * kz + 3 * ni + 8, 0.0);
  MassArgs a;
  std::memset(&a, 0, sizeof(a));
  a.g = g;
  a.rho = P(e, MB_RHO); a.zq = P(e, MB_ZETAF); a.qx = P(e, MB_QX); a.u = P(e, MB_U); a.v = P(e, MB_V); a.ps = P(e, MB_PS);
  a.rows = work.data(); a.lev = a.rows + (size_t)2 * kz * ni; a.psrow = a.lev + 4 * kz; a.out = a.psrow + 3 * ni;
  a.dxsq = e.cfg.dx * e.cfg.dx; a.dt = e.cfg.dtsec; a.dx = e.cfg.dx; a.ni = ni;
  if (what & 1) {
    if (!a.zq) return fail("emu: do_massck not configured");
    walk(o, 1, kz, [&](int k) { walk(o, g.ice1, g.ice2, [&](int i) { massck_row(a, i, k); }); });
    walk(o, 1, kz, [&](int k) { massck_bdy_level(a, k); });
  }
  if (what & 2) walk(o, g.ice1, g.ice2, [&](int i) { ps_row(a, i); });
  massck_final(a);
  std::copy(a.out, a.out + 7, out7);
  return 0;
}
int emu_b200_massck(void* h, double* out4) {
  double o7[7];
  if (emu_b200_massck7(h, 1, o7)) return 1;
  std::copy(o7, o7 + 4, out4);
  return 0;
}
int emu_b200_ps_check(void* h, double* maxmin, int32_t* nonfinite) {
  double o7[7];
  if (emu_b200_massck7(h, 2, o7)) return 1;
  maxmin[0] = o7[4]; maxmin[1] = o7[5]; *nonfinite = (int32_t)o7[6];
  return 0;
}
// TKE helpers = k_tke_destagger / k_tke_restagger / k_tke_update
int emu_b200_tke_destagger(void* h) {
  Emu& e = *(Emu*)h; const Geo& g = e.g; const int o = e.order;
  if (!P(e, MB_TKE)) return fail("emu: ibltyp != 2");
  walk(o, 1, g.kz, [&](int k) { walk(o, g.ice1, g.ice2, [&](int i) { walk(o, g.jce1, g.jce2, [&](int j) {
    zstagtoh_cell(g, P(e, MB_TKE), P(e, MB_TKEX), j, i, k); }); }); });
  return 0;
}
int emu_b200_tke_restagger(void* h) {
  Emu& e = *(Emu*)h; const Geo& g = e.g; const int o = e.order;
  if (!P(e, MB_TKE)) return fail("emu: ibltyp != 2");
  walk(o, 2, g.kz, [&](int k) { walk(o, g.ice1, g.ice2, [&](int i) { walk(o, g.jce1, g.jce2, [&](int j) {
    htozstag_cell(g, P(e, MB_TKEX), P(e, MB_TKE), j, i, k); }); }); });
  return 0;
}
int emu_b200_tke_update(void* h) {
  Emu& e = *(Emu*)h; const Geo& g = e.g; const int o = e.order;
  if (!P(e, MB_TKE)) return fail("emu: ibltyp != 2");
  walk(o, 1, g.kz + 1, [&](int k) { walk(o, g.ici1, g.ici2, [&](int i) { walk(o, g.jci1, g.jci2, [&](int j) {
    tke_update_cell(g, P(e, MB_TKE), P(e, MB_TKETEN), e.cfg.dtsec, e.cfg.tkemin, j, i, k); }); }); });
  return 0;
}
}  // extern "C"
