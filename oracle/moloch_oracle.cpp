/*
 * moloch_oracle.cpp -- CPU oracle for the MOLOCH dycore step (TEST INFRASTRUCTURE).
 *
 * Loop-for-loop FP64 restatement of /root/reference/Main/mod_moloch.F90
 * (RegCM 5.0.0) plus the set-up code it depends on.  Build with
 * -ffp-contract=off so that no FMA contraction happens: the CUDA product is
 * built with -fmad=false and keeps the reference's operation order, which
 * makes +,-,*,/ results bit-identical between the two.
 *
 * Pinned on the reference's own source, executed through the mechanical
 * translator of oracle/refrun (bit for bit; see moloch_oracle.h and
 * tests/test_reference_pin.py).
 *
 * Every routine cites the reference lines it restates as  [F90:a-b]  (lines of
 * Main/mod_moloch.F90) or with an explicit file name.
 */
#include "moloch_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------------------
// Constants: Share/mod_constants.F90 (non-RCEMIP branch) :105-233,273-282,347-386
// ---------------------------------------------------------------------------
constexpr double egrav = 9.80665;
constexpr double boltzk = 1.3806490e-23;
constexpr double navgdr = 6.02214076e23;
constexpr double amd = 28.96454;
constexpr double amw = 18.01528;
constexpr double rgasmol = navgdr * boltzk;
constexpr double rgas = (rgasmol / amd) * 1000.0;
constexpr double cpd = 3.5 * rgas;
constexpr double cvd = 2.5 * rgas;
constexpr double regrav = 1.0 / egrav;
constexpr double rcpd = 1.0 / cpd;
constexpr double rovcp = rgas * rcpd;
constexpr double rdrcv = rgas / cvd;
constexpr double cpovr = cpd / rgas;
constexpr double govr = egrav / rgas;
constexpr double govcp = egrav / cpd;
constexpr double p00 = 1.0e5;
constexpr double lrate = 0.00649;
constexpr double tzero = 273.15;
constexpr double ep1 = amd / amw - 1.0;
constexpr double ep2 = amw / amd;
constexpr double mathpi = 3.14159265358979323846;
constexpr double degrad = mathpi / 180.0;
constexpr double eomeg2 = 2.0 * 7.2921159e-5;
constexpr double rearthrad = 1.0 / 6.371229e6;
constexpr double mo_zfilt_fac = 0.8;  // Main/mod_init.F90:62
// Main/mpplib/mod_runparams.F90:184-192
const double qxcheckval[10] = {1.0e-8, 1.0e-16, 1.0e-16, 1.0e-16, 1.0e-16,
                               1.0e-16, 1.0e-16, 1.0e10, 100.0, 0.01};
const double qxzeroval[10] = {1.0e-8, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0e10, 100.0, 0.01};

thread_local std::string g_err;

// ---------------------------------------------------------------------------
// Arrays with Fortran lower bounds, (j,i,k) with j fastest
// ---------------------------------------------------------------------------
struct Arr {
  std::vector<double> d;
  int jlo = 0, jhi = -1, ilo = 0, ihi = -1, klo = 1, khi = 0;
  long sj = 0, sk = 0;  // row length, plane size
  void alloc(int jl, int jh, int il, int ih, int kl = 1, int kh = 1) {
    jlo = jl; jhi = jh; ilo = il; ihi = ih; klo = kl; khi = kh;
    sj = jh - jl + 1; sk = sj * (long)(ih - il + 1);
    d.assign((size_t)(sk * (kh - kl + 1)), 0.0);  // getmem zero-initialises (Share/mod_space.F90:304-311)
  }
  bool has(int j, int i) const { return j >= jlo && j <= jhi && i >= ilo && i <= ihi; }
  inline double& operator()(int j, int i, int k = 1) {
#ifdef ORACLE_BOUNDS_CHECK
    if (j < jlo || j > jhi || i < ilo || i > ihi || k < klo || k > khi) {
      std::fprintf(stderr, "oracle: out of bounds (%d,%d,%d) in [%d:%d,%d:%d,%d:%d]\n", j, i, k,
                   jlo, jhi, ilo, ihi, klo, khi);
      std::abort();
    }
#endif
    return d[(size_t)((long)(k - klo) * sk + (long)(i - ilo) * sj + (j - jlo))];
  }
  inline double operator()(int j, int i, int k = 1) const {
    return const_cast<Arr*>(this)->operator()(j, i, k);
  }
};

// ---------------------------------------------------------------------------
// Geometry of one subdomain: set_nproc (mod_mppparam.F90:1250-1641) +
// setup_model_indexes (mod_atm_interface.F90:182-382)
// ---------------------------------------------------------------------------
struct Geom {
  int jx, iy, kz, kzp1, kzm1;
  bool band, crm;
  int px, py, rank, locj, loci;
  int left, right, bottom, top;  // -1 == mpi_proc_null
  bool bl, br, bb, bt;           // has_bdyleft/right/bottom/top
  int jde1, jde2, ide1, ide2, jdi1, jdi2, idi1, idi2, jdii1, jdii2, idii1, idii2;
  int jce1, jce2, ice1, ice2, jci1, jci2, ici1, ici2;
  int gl, gr, gb, gt;  // ma%jbl1, jbr1, ibb1, ibt1 (1 if a neighbour exists there)
  int jmin, jmax, imin, imax;
  int jce1ga() const { return jce1 - gl; }
  int jce2ga() const { return jce2 + gr; }
  int ice1ga() const { return ice1 - gb; }
  int ice2ga() const { return ice2 + gt; }
  int jce1gb() const { return jce1 - 2 * gl; }
  int jce2gb() const { return jce2 + 2 * gr; }
  int ice1gb() const { return ice1 - 2 * gb; }
  int ice2gb() const { return ice2 + 2 * gt; }
  int jde1ga() const { return jde1 - gl; }
  int jde2ga() const { return jde2 + gr; }
  int ide1ga() const { return ide1 - gb; }
  int ide2ga() const { return ide2 + gt; }
  int jde1gb() const { return jde1 - 2 * gl; }
  int jde2gb() const { return jde2 + 2 * gr; }
  int ide1gb() const { return ide1 - 2 * gb; }
  int ide2gb() const { return ide2 + 2 * gt; }
};

bool make_geom(const oracle_config& c, int rank, Geom& g) {
  g.jx = c.jx; g.iy = c.iy; g.kz = c.kz; g.kzp1 = c.kz + 1; g.kzm1 = c.kz - 1;
  g.band = (c.i_band == 1 || c.i_crm == 1);
  g.crm = (c.i_crm == 1);
  g.px = c.px; g.py = c.py; g.rank = rank;
  const int nproc = c.px * c.py;
  int gdj1, gdj2, gdi1, gdi2, gcj2, gci2;
  if (nproc == 1) {
    g.locj = g.loci = 0;
    gdj1 = 1; gdi1 = 1; gdj2 = c.jx; gdi2 = c.iy;
    gcj2 = g.band ? c.jx : c.jx - 1;
    gci2 = g.crm ? c.iy : c.iy - 1;
    g.left = g.right = g.top = g.bottom = -1;
    g.bl = g.br = g.bb = g.bt = false;
    if (g.crm) {
      g.left = g.right = g.top = g.bottom = 0;
    } else {
      g.bt = g.bb = true;
      if (g.band) { g.left = g.right = 0; } else { g.bl = g.br = true; }
    }
  } else {
    // location(1) = rank / cpus_per_dim(2), location(2) = rank mod cpus_per_dim(2)
    g.locj = rank / c.py; g.loci = rank % c.py;
    auto cart = [&](int lj, int li) -> int {
      if (lj < 0 || lj >= c.px) { if (!g.band) return -1; lj = (lj + c.px) % c.px; }
      if (li < 0 || li >= c.py) { if (!g.crm) return -1; li = (li + c.py) % c.py; }
      return lj * c.py + li;
    };
    g.left = cart(g.locj - 1, g.loci); g.right = cart(g.locj + 1, g.loci);
    g.bottom = cart(g.locj, g.loci - 1); g.top = cart(g.locj, g.loci + 1);
    g.bt = (g.top < 0); g.bb = (g.bottom < 0); g.br = (g.right < 0); g.bl = (g.left < 0);
    int jxp = c.jx / c.px, iyp = c.iy / c.py;
    gdj1 = g.locj * jxp + 1; gdi1 = g.loci * iyp + 1;
    if (jxp * c.px < c.jx) {
      int imiss = c.jx - jxp * c.px;
      if (g.locj < imiss) { gdj1 += g.locj; jxp += 1; } else { gdj1 += imiss; }
    }
    if (iyp * c.py < c.iy) {
      int imiss = c.iy - iyp * c.py;
      if (g.loci < imiss) { gdi1 += g.loci; iyp += 1; } else { gdi1 += imiss; }
    }
    gdj2 = gdj1 + jxp - 1; gdi2 = gdi1 + iyp - 1;
    if (jxp < 3 || iyp < 3) { g_err = "Cannot have one processor with less than 3x3 points"; return false; }
    gci2 = gdi2; if (!g.crm && gdi2 == c.iy) gci2 -= 1;
    gcj2 = gdj2; if (!g.band && gdj2 == c.jx) gcj2 -= 1;
  }
  g.gl = g.bl ? 0 : 1; g.gr = g.br ? 0 : 1; g.gb = g.bb ? 0 : 1; g.gt = g.bt ? 0 : 1;
  g.jde1 = g.jdi1 = g.jdii1 = gdj1; g.jde2 = g.jdi2 = g.jdii2 = gdj2;
  g.ide1 = g.idi1 = g.idii1 = gdi1; g.ide2 = g.idi2 = g.idii2 = gdi2;
  if (g.bl) { g.jdi1 = g.jde1 + 1; g.jdii1 = g.jde1 + 2; }
  if (g.br) { g.jdi2 = g.jde2 - 1; g.jdii2 = g.jde2 - 2; }
  if (g.bb) { g.idi1 = g.ide1 + 1; g.idii1 = g.ide1 + 2; }
  if (g.bt) { g.idi2 = g.ide2 - 1; g.idii2 = g.ide2 - 2; }
  g.jce1 = g.jci1 = gdj1; g.jce2 = g.jci2 = gcj2;
  g.ice1 = g.ici1 = gdi1; g.ice2 = g.ici2 = gci2;
  if (g.bl) g.jci1 = g.jce1 + 1;
  if (g.br) g.jci2 = g.jce2 - 1;
  if (g.bb) g.ici1 = g.ice1 + 1;
  if (g.bt) g.ici2 = g.ice2 - 1;
  // init_moloch [F90:280-293]
  const int jcross2 = g.band ? c.jx : c.jx - 1, icross2 = g.crm ? c.iy : c.iy - 1;
  g.jmin = 1; g.jmax = jcross2; g.imin = 1; g.imax = icross2;
  if (g.band) { g.jmin = 1 - 2; g.jmax = jcross2 + 2; }
  if (g.crm) { g.jmin = 1 - 2; g.jmax = jcross2 + 2; g.imin = 1 - 2; g.imax = icross2 + 2; }
  return true;
}

enum Stag { S_CROSS, S_U, S_V, S_DOT };  // which points an array lives on

struct FieldInfo {
  Arr* a; Stag stag; int nk;  // nk levels starting at a->klo
};

struct Rank {
  Geom g;
  // state (mo_atm) -- Main/mod_atm_interface.F90:579-624
  Arr u, v, ux, vx, w, pai, tetav, t, tvirt, p, rho, qsat, zeta, zetaf, fmz, fmzf, rfmzu, rfmzv;
  std::vector<Arr> qx, trac, qxten, chiten;
  Arr tten, uten, vten;
  // moloch work arrays -- [F90:159-199]
  Arr s, wwkw, wz, tetavf, p0, wfw, zpby, zpbw, zdiv2, wx, laplacian, bdywtu, bdywtv, bdywtw, ud, vd;
  // 2-D
  Arr mx, mu, mv, mx2, rmx, rmu, rmv, hx, hy, coru, corv, ht, htu, htv, ulat, vlat, ps, ibnd_cr,
      ibnd_ud, ibnd_vd;
  // lateral boundary buffers: v3dbound/v2dbound b0, b1 (Main/mod_atm_interface.F90:547-577)
  Arr dub0, dub1, dvb0, dvb1, xtb0, xtb1, xpaib0, xpaib1, xqb0, xqb1, xlb0, xlb1, xib0, xib1, xpsb0, xpsb1;
  std::vector<Arr> chib0, chib1;
  // UW-PBL TKE (ibltyp == 2): Main/mod_atm_interface.F90:609-612, [F90:184]
  Arr tke, tketen, tkex;
  // mkslice outputs (Main/mod_atm_interface.F90:964-1012, :599)
  Arr pf3d, th3d, rhb3d, wpx3d, rhox2d, tp2d, th700, xlat, ptrop, ktrop, kmxpbl;
  // mospectral_nudge work arrays (Main/mod_bdycod.F90:422, :3868-3873), contiguous like the Fortran ones
  Arr zn1; std::vector<double> sx, sxg, sy, syg;
  // tendency diagnostics [F90:187-192], tdiag%adh/bdy, qdiag%adh/bdy, cadvhdiag, cbdydiag
  Arr ten0, qen0, tdiag_adh, qdiag_adh, tdiag_bdy, qdiag_bdy;
  std::vector<Arr> chiten0, cadvhdiag, cbdydiag;
  std::map<std::string, FieldInfo> reg;
};

struct World {
  oracle_config c;
  std::vector<Rank> r;
  // scalars / 1-D (identical on every rank)
  std::vector<double> zita, zitah, gzitak, gzitakh, xkdamp, xknu, ffilt, rlat, hefc;  // 1-based via [k]
  double mo_dzita, rdzita, dx, rdx, dtsec, dtstepa, dtsound;
  bool lrotllr, do_divdamp, do_divfilter;
  int nqx, ntr, iqfrst, ipptls;
  // boundary / slice extension (oracle_set_ext)
  oracle_ext_config x{};
  bool ext_set = false;
  double rtb = 0.0, xbctime = 0.0, tspectral = 0.0;
  int nztop = 0, km = 0, lm = 0;
  std::vector<double> gmeanz, tnudge, cnudge, fcx;  // 1-based via [k]
  std::vector<double> bvx, bvy;                      // lowpass basis, global: bvx[(k-1)*jx + (j-1)]
};

template <class F> void each(World& w, F f) { for (auto& r : w.r) f(r); }

#define PAR2 _Pragma("omp parallel for collapse(2) schedule(static)")
#define PAR1 _Pragma("omp parallel for schedule(static)")

// ---------------------------------------------------------------------------
// Halo exchange emulation: real8_3d_exchange_left_right_bottom_top and friends
// (Main/mpplib/mod_mppparam.F90:3809-3878, 4257-4309, 4661-4712).  Ghost
// (j1-iex) of a rank receives the neighbour's (j2-(iex-1)) and so on; no
// corners; a side is skipped when the neighbour is mpi_proc_null.
// ---------------------------------------------------------------------------
struct Box { int j1, j2, i1, i2; };
using BoxFn = std::function<Box(const Geom&)>;
using ArrFn = std::function<Arr&(Rank&)>;

void exchange(World& w, const ArrFn& get, int nex, bool lr, bool bt, const BoxFn& box) {
  for (auto& r : w.r) {
    const Geom& g = r.g;
    Arr& a = get(r);
    const Box b = box(g);
    const int k1 = a.klo, k2 = a.khi;
    if (lr) {
      if (g.left >= 0) {
        Rank& n = w.r[g.left]; Arr& na = get(n); const Box nb = box(n.g);
        for (int k = k1; k <= k2; ++k) for (int i = b.i1; i <= b.i2; ++i) for (int iex = 1; iex <= nex; ++iex)
          a(b.j1 - iex, i, k) = na(nb.j2 - (iex - 1), i, k);
      }
      if (g.right >= 0) {
        Rank& n = w.r[g.right]; Arr& na = get(n); const Box nb = box(n.g);
        for (int k = k1; k <= k2; ++k) for (int i = b.i1; i <= b.i2; ++i) for (int iex = 1; iex <= nex; ++iex)
          a(b.j2 + iex, i, k) = na(nb.j1 + (iex - 1), i, k);
      }
    }
    if (bt) {
      if (g.bottom >= 0) {
        Rank& n = w.r[g.bottom]; Arr& na = get(n); const Box nb = box(n.g);
        for (int k = k1; k <= k2; ++k) for (int iex = 1; iex <= nex; ++iex) for (int j = b.j1; j <= b.j2; ++j)
          a(j, b.i1 - iex, k) = na(j, nb.i2 - (iex - 1), k);
      }
      if (g.top >= 0) {
        Rank& n = w.r[g.top]; Arr& na = get(n); const Box nb = box(n.g);
        for (int k = k1; k <= k2; ++k) for (int iex = 1; iex <= nex; ++iex) for (int j = b.j1; j <= b.j2; ++j)
          a(j, b.i2 + iex, k) = na(j, nb.i1 + (iex - 1), k);
      }
    }
  }
}
const BoxFn BOX_CROSS = [](const Geom& g) { return Box{g.jce1, g.jce2, g.ice1, g.ice2}; };
const BoxFn BOX_U = [](const Geom& g) { return Box{g.jde1, g.jde2, g.ice1, g.ice2}; };
const BoxFn BOX_V = [](const Geom& g) { return Box{g.jce1, g.jce2, g.ide1, g.ide2}; };
const BoxFn BOX_DOT = [](const Geom& g) { return Box{g.jde1, g.jde2, g.ide1, g.ide2}; };
const BoxFn BOX_P0 = [](const Geom& g) { return Box{g.jce1, g.jce2, g.ici1, g.ici2}; };

// ---------------------------------------------------------------------------
// Share/mod_zita.F90:40-178
// ---------------------------------------------------------------------------
inline double zfz(double ztop, double zh) { return ztop / (std::exp(ztop / zh) - 1.0); }
inline double bzita(double zita, double ztop, double zh) { return zfz(ztop, zh) * (std::exp(zita / zh) - 1.0); }
inline double bzitap(double zita, double ztop, double zh) { return zfz(ztop, zh) * std::exp(zita / zh) / zh; }
inline double gzita(double zita, double ztop, double a0) {
  const double ratio = zita / ztop;
  return ((0.0 - 1.0 * a0) * ratio - (3.0 - 2.0 * a0) * (ratio * ratio) +
          (2.0 - 1.0 * a0) * (ratio * ratio * ratio)) + 1.0;
}
inline double gzitap(double zita, double ztop, double a0) {
  const double ratio = zita / ztop;
  return ((0.0 - 1.0 * a0) * 1.0 - (6.0 - 4.0 * a0) * ratio + (6.0 - 3.0 * a0) * (ratio * ratio)) / ztop;
}
inline double md_fmz_h(double zita, double orog, double ztop, double zh, double a0) {
  return 1.0 / (gzitap(zita, ztop, a0) * orog + bzitap(zita, ztop, zh));
}
inline double md_zeta_h(double zita, double orog, double ztop, double zh, double a0) {
  return orog * (gzita(zita, ztop, a0) - 1.0) + bzita(zita, ztop, zh);
}
// NB: in the reference md_fmz/md_zeta take the *geopotential* and multiply by
// regrav; mddom%ht holds geopotential for idynamic==3.  Here "ht" is passed as
// geopotential too (height*egrav), see oracle_setup_static.
inline double md_fmz(double zita, double geopot, double ztop, double zh, double a0) {
  return md_fmz_h(zita, geopot * regrav, ztop, zh, a0);
}
inline double md_zeta(double zita, double geopot, double ztop, double zh, double a0) {
  return md_zeta_h(zita, geopot * regrav, ztop, zh, a0);
}

// Share/pfwsat.inc
inline double pfwsat(double t, double p) {
  const double a0 = 0.611213476e+03, a1 = 0.444007856e+02, a2 = 0.143064234e+01, a3 = 0.264461437e-01,
               a4 = 0.305903558e-03, a5 = 0.196237241e-05, a6 = 0.892344772e-08, a7 = -0.373208410e-10,
               a8 = 0.209339997e-13;
  const double c0 = 0.611123516e+03, c1 = 0.503109514e+02, c2 = 0.188369801e+01, c3 = 0.420547422e-01,
               c4 = 0.614396778e-03, c5 = 0.602780717e-05, c6 = 0.387940929e-07, c7 = 0.149436277e-09,
               c8 = 0.262655803e-12;
  const double t_limit = t - tzero;
  const double td = std::min(std::max(t_limit, -75.0), 100.0);
  double es;
  if (td >= 0.0) {
    es = std::min(a0 + td * (a1 + td * (a2 + td * (a3 + td * (a4 + td * (a5 + td * (a6 + td * (a7 + td * a8))))))),
                  0.15 * p);
  } else {
    es = std::min(c0 + td * (c1 + td * (c2 + td * (c3 + td * (c4 + td * (c5 + td * (c6 + td * (c7 + td * c8))))))),
                  0.15 * p);
  }
  return ep2 * (es / (p - es));
}

// [F90:1571-1590]
inline double local_flow_param(double num, double den) {
  const double minden = 1.0e-30;
  const float minnum = (float)minden;  // real(rk4), parameter :: minnum = minden
  if (std::fabs(den) < minden) {
    if (std::fabs(num) < (double)minnum) return 1.0;
    return 0.0;
  }
  return num / den;
}

// ---------------------------------------------------------------------------
// allocation: allocate_atmosphere + allocate_moloch
// ---------------------------------------------------------------------------
void alloc_rank(World& w, Rank& r) {
  const Geom& g = r.g; const int kz = g.kz, kzp1 = g.kzp1;
  r.u.alloc(g.jde1gb(), g.jde2gb(), g.ice1ga(), g.ice2ga(), 1, kz);
  r.v.alloc(g.jce1ga(), g.jce2ga(), g.ide1gb(), g.ide2gb(), 1, kz);
  r.ux.alloc(g.jce1gb(), g.jce2gb(), g.ice1ga(), g.ice2ga(), 1, kz);
  r.vx.alloc(g.jce1ga(), g.jce2ga(), g.ice1gb(), g.ice2gb(), 1, kz);
  r.tetav.alloc(g.jce1ga(), g.jce2ga(), g.ice1ga(), g.ice2ga(), 1, kz);
  r.t.alloc(g.jce1ga(), g.jce2ga(), g.ice1ga(), g.ice2ga(), 1, kz);
  r.w.alloc(g.jce1, g.jce2, g.ice1, g.ice2, 1, kzp1);
  r.pai.alloc(g.jce1ga(), g.jce2ga(), g.ice1ga(), g.ice2ga(), 1, kz);
  r.qx.resize(w.nqx); for (auto& a : r.qx) a.alloc(g.jce1ga(), g.jce2ga(), g.ice1ga(), g.ice2ga(), 1, kz);
  r.trac.resize(w.ntr); for (auto& a : r.trac) a.alloc(g.jce1ga(), g.jce2ga(), g.ice1ga(), g.ice2ga(), 1, kz);
  r.zeta.alloc(g.jce1gb(), g.jce2gb(), g.ice1gb(), g.ice2gb(), 1, kz);
  r.rho.alloc(g.jce1, g.jce2, g.ice1, g.ice2, 1, kz);
  r.p.alloc(g.jce1, g.jce2, g.ice1, g.ice2, 1, kz);
  r.tvirt.alloc(g.jce1, g.jce2, g.ice1, g.ice2, 1, kz);
  r.zetaf.alloc(g.jce1, g.jce2, g.ice1, g.ice2, 1, kzp1);
  r.qsat.alloc(g.jce1, g.jce2, g.ice1, g.ice2, 1, kz);
  r.tten.alloc(g.jci1, g.jci2, g.ici1, g.ici2, 1, kz);
  r.uten.alloc(g.jci1, g.jci2, g.ici1, g.ici2, 1, kz);
  r.vten.alloc(g.jci1, g.jci2, g.ici1, g.ici2, 1, kz);
  r.qxten.resize(w.nqx); for (auto& a : r.qxten) a.alloc(g.jci1, g.jci2, g.ici1, g.ici2, 1, kz);
  r.chiten.resize(w.ntr); for (auto& a : r.chiten) a.alloc(g.jci1, g.jci2, g.ici1, g.ici2, 1, kz);
  r.fmz.alloc(g.jce1ga(), g.jce2ga(), g.ice1ga(), g.ice2ga(), 1, kz);
  r.rfmzu.alloc(g.jde1ga(), g.jde2ga(), g.ice1ga(), g.ice2ga(), 1, kz);
  r.rfmzv.alloc(g.jce1ga(), g.jce2ga(), g.ide1ga(), g.ide2ga(), 1, kz);
  r.fmzf.alloc(g.jce1, g.jce2, g.ice1, g.ice2, 1, kzp1);
  // allocate_moloch [F90:159-199]
  r.laplacian.alloc(g.jci1, g.jci2, g.ici1, g.ici2, 1, kz);
  r.bdywtu.alloc(g.jdi1, g.jdi2, g.ici1, g.ici2, 1, kz);
  r.bdywtv.alloc(g.jci1, g.jci2, g.idi1, g.idi2, 1, kz);
  r.bdywtw.alloc(g.jci1, g.jci2, g.ici1, g.ici2, 1, kz);
  r.wwkw.alloc(g.jce1, g.jce2, g.ice1, g.ice2, 2, kzp1);
  r.tetavf.alloc(g.jce1, g.jce2, g.ice1, g.ice2, 2, kz);
  r.s.alloc(g.jce1, g.jce2, g.ice1, g.ice2, 1, kzp1);
  r.zdiv2.alloc(g.jce1ga(), g.jce2ga(), g.ice1ga(), g.ice2ga(), 1, kz);
  r.wz.alloc(g.jce1gb(), g.jce2gb(), g.ice1gb(), g.ice2gb(), 1, kz);
  r.p0.alloc(g.jce1gb(), g.jce2gb(), g.ice1gb(), g.ice2gb(), 1, kz);
  r.wfw.alloc(g.jce1, g.jce2, g.ice1, g.ice2, 1, kzp1);
  r.wx.alloc(g.jce1ga(), g.jce2ga(), g.ice1ga(), g.ice2ga(), 1, kz);
  r.zpby.alloc(g.jce1, g.jce2, g.ici1, g.ice2ga(), 1, kz);
  r.zpbw.alloc(g.jci1, g.jce2ga(), g.ice1, g.ice2, 1, kz);
  r.ud.alloc(g.jde1, g.jde2, g.ice1, g.ice2, 1, kz);
  r.vd.alloc(g.jce1, g.jce2, g.ide1, g.ide2, 1, kz);
  // 2-D
  for (Arr* a : {&r.mx, &r.mu, &r.mv, &r.mx2, &r.rmx, &r.rmu, &r.rmv})
    a->alloc(g.jde1ga(), g.jde2ga(), g.ide1ga(), g.ide2ga());
  for (Arr* a : {&r.ht, &r.htu, &r.htv}) a->alloc(g.jde1gb(), g.jde2gb(), g.ide1gb(), g.ide2gb());
  r.hx.alloc(g.jde1ga(), g.jde2ga(), g.ice1, g.ice2);
  r.hy.alloc(g.jce1, g.jce2, g.ide1ga(), g.ide2ga());
  r.coru.alloc(g.jde1, g.jde2, g.ice1, g.ice2);
  r.corv.alloc(g.jce1, g.jce2, g.ide1, g.ide2);
  r.ulat.alloc(g.jde1, g.jde2, g.ide1, g.ide2);
  r.vlat.alloc(g.jde1, g.jde2, g.ide1, g.ide2);
  r.ps.alloc(g.jce1, g.jce2, g.ice1, g.ice2);
  for (Arr* a : {&r.ibnd_cr, &r.ibnd_ud, &r.ibnd_vd}) a->alloc(g.jde1, g.jde2, g.ide1, g.ide2);

  auto R = [&](const char* n, Arr* a, Stag s, int nk) { r.reg[n] = FieldInfo{a, s, nk}; };
  R("u", &r.u, S_U, kz); R("v", &r.v, S_V, kz); R("ux", &r.ux, S_CROSS, kz); R("vx", &r.vx, S_CROSS, kz);
  R("w", &r.w, S_CROSS, kzp1); R("pai", &r.pai, S_CROSS, kz); R("tetav", &r.tetav, S_CROSS, kz);
  R("t", &r.t, S_CROSS, kz); R("tvirt", &r.tvirt, S_CROSS, kz); R("p", &r.p, S_CROSS, kz);
  R("rho", &r.rho, S_CROSS, kz); R("qsat", &r.qsat, S_CROSS, kz); R("zeta", &r.zeta, S_CROSS, kz);
  R("zetaf", &r.zetaf, S_CROSS, kzp1); R("fmz", &r.fmz, S_CROSS, kz); R("fmzf", &r.fmzf, S_CROSS, kzp1);
  R("rfmzu", &r.rfmzu, S_U, kz); R("rfmzv", &r.rfmzv, S_V, kz);
  R("tten", &r.tten, S_CROSS, kz); R("uten", &r.uten, S_CROSS, kz); R("vten", &r.vten, S_CROSS, kz);
  R("s", &r.s, S_CROSS, kzp1); R("zdiv2", &r.zdiv2, S_CROSS, kz); R("wx", &r.wx, S_CROSS, kz);
  R("wz", &r.wz, S_CROSS, kz); R("p0", &r.p0, S_CROSS, kz);
  R("bdywtu", &r.bdywtu, S_U, kz); R("bdywtv", &r.bdywtv, S_V, kz); R("bdywtw", &r.bdywtw, S_CROSS, kz);
  R("msfx", &r.mx, S_DOT, 1); R("msfu", &r.mu, S_DOT, 1); R("msfv", &r.mv, S_DOT, 1);
  R("ht", &r.ht, S_DOT, 1); R("htu", &r.htu, S_DOT, 1); R("htv", &r.htv, S_DOT, 1);
  R("hx", &r.hx, S_U, 1); R("hy", &r.hy, S_V, 1); R("coru", &r.coru, S_U, 1); R("corv", &r.corv, S_V, 1);
  R("ulat", &r.ulat, S_DOT, 1); R("vlat", &r.vlat, S_DOT, 1); R("ps", &r.ps, S_CROSS, 1);
  R("mx2", &r.mx2, S_DOT, 1); R("rmx", &r.rmx, S_DOT, 1); R("rmu", &r.rmu, S_DOT, 1); R("rmv", &r.rmv, S_DOT, 1);
}

// arrays of the boundary / slice / TKE extension (oracle_set_ext)
void alloc_ext(World& w, Rank& r) {
  const Geom& g = r.g; const int kz = g.kz, kzp1 = g.kzp1;
  auto R = [&](const char* n, Arr* a, Stag s, int nk) { r.reg[n] = FieldInfo{a, s, nk}; };
  if (w.x.do_bdy) {
    // allocate_v3dbound / allocate_v2dbound (Main/mod_atm_interface.F90:547-577)
    for (Arr* a : {&r.dub0, &r.dub1, &r.dvb0, &r.dvb1}) a->alloc(g.jde1gb(), g.jde2gb(), g.ide1gb(), g.ide2gb(), 1, kz);
    for (Arr* a : {&r.xtb0, &r.xtb1, &r.xpaib0, &r.xpaib1, &r.xqb0, &r.xqb1, &r.xlb0, &r.xlb1, &r.xib0, &r.xib1})
      a->alloc(g.jce1ga(), g.jce2ga(), g.ice1ga(), g.ice2ga(), 1, kz);
    for (Arr* a : {&r.xpsb0, &r.xpsb1}) a->alloc(g.jce1ga(), g.jce2ga(), g.ice1ga(), g.ice2ga());
    if (w.x.ichem == 1 && w.x.ichebdy != 0) {
      r.chib0.resize(w.ntr); r.chib1.resize(w.ntr);
      for (auto& a : r.chib0) a.alloc(g.jce1ga(), g.jce2ga(), g.ice1ga(), g.ice2ga(), 1, kz);
      for (auto& a : r.chib1) a.alloc(g.jce1ga(), g.jce2ga(), g.ice1ga(), g.ice2ga(), 1, kz);
    }
    R("dub0", &r.dub0, S_U, kz); R("dub1", &r.dub1, S_U, kz); R("dvb0", &r.dvb0, S_V, kz); R("dvb1", &r.dvb1, S_V, kz);
    R("xtb0", &r.xtb0, S_CROSS, kz); R("xtb1", &r.xtb1, S_CROSS, kz);
    R("xpaib0", &r.xpaib0, S_CROSS, kz); R("xpaib1", &r.xpaib1, S_CROSS, kz);
    R("xqb0", &r.xqb0, S_CROSS, kz); R("xqb1", &r.xqb1, S_CROSS, kz);
    R("xlb0", &r.xlb0, S_CROSS, kz); R("xlb1", &r.xlb1, S_CROSS, kz);
    R("xib0", &r.xib0, S_CROSS, kz); R("xib1", &r.xib1, S_CROSS, kz);
    R("xpsb0", &r.xpsb0, S_CROSS, 1); R("xpsb1", &r.xpsb1, S_CROSS, 1);
    r.zn1.alloc(g.jde1, g.jde2, g.ide1, g.ide2);
  }
  if (w.x.ibltyp == 2) {
    r.tke.alloc(g.jce1, g.jce2, g.ice1, g.ice2, 1, kzp1);
    r.tketen.alloc(g.jci1, g.jci2, g.ici1, g.ici2, 1, kzp1);
    r.tkex.alloc(g.jce1, g.jce2, g.ice1, g.ice2, 1, kz);
    R("tke", &r.tke, S_CROSS, kzp1); R("tketen", &r.tketen, S_CROSS, kzp1); R("tkex", &r.tkex, S_CROSS, kz);
  }
  r.pf3d.alloc(g.jce1, g.jce2, g.ice1, g.ice2, 1, kzp1);
  r.th3d.alloc(g.jce1, g.jce2, g.ice1, g.ice2, 1, kz);
  r.rhb3d.alloc(g.jci1, g.jci2, g.ici1, g.ici2, 1, kz);
  r.wpx3d.alloc(g.jci1, g.jci2, g.ici1, g.ici2, 1, kz);
  r.rhox2d.alloc(g.jci1, g.jci2, g.ici1, g.ici2);
  r.tp2d.alloc(g.jci1, g.jci2, g.ici1, g.ici2);
  r.th700.alloc(g.jci1, g.jci2, g.ici1, g.ici2);
  R("pf3d", &r.pf3d, S_CROSS, kzp1); R("th3d", &r.th3d, S_CROSS, kz); R("rhb3d", &r.rhb3d, S_CROSS, kz);
  R("wpx3d", &r.wpx3d, S_CROSS, kz); R("rhox2d", &r.rhox2d, S_CROSS, 1); R("tp2d", &r.tp2d, S_CROSS, 1);
  R("th700", &r.th700, S_CROSS, 1);
  if (w.x.idiag > 0) {
    for (Arr* a : {&r.ten0, &r.qen0, &r.tdiag_adh, &r.qdiag_adh, &r.tdiag_bdy, &r.qdiag_bdy})
      a->alloc(g.jci1, g.jci2, g.ici1, g.ici2, 1, kz);
    R("ten0", &r.ten0, S_CROSS, kz); R("qen0", &r.qen0, S_CROSS, kz);
    R("tdiag_adh", &r.tdiag_adh, S_CROSS, kz); R("qdiag_adh", &r.qdiag_adh, S_CROSS, kz);
    R("tdiag_bdy", &r.tdiag_bdy, S_CROSS, kz); R("qdiag_bdy", &r.qdiag_bdy, S_CROSS, kz);
  }
  if (w.x.ichdiag > 0 && w.x.ichem == 1 && w.ntr > 0)
    for (auto* v : {&r.chiten0, &r.cadvhdiag, &r.cbdydiag}) {
      v->resize(w.ntr);
      for (auto& a : *v) a.alloc(g.jci1, g.jci2, g.ici1, g.ici2, 1, kz);
    }
  r.xlat.alloc(g.jde1, g.jde2, g.ide1, g.ide2);
  for (Arr* a : {&r.ptrop, &r.ktrop, &r.kmxpbl}) a->alloc(g.jci1, g.jci2, g.ici1, g.ici2);
  R("xlat", &r.xlat, S_CROSS, 1); R("ptrop", &r.ptrop, S_CROSS, 1); R("ktrop", &r.ktrop, S_CROSS, 1);
  R("kmxpbl", &r.kmxpbl, S_CROSS, 1);
}

// setup_bdycon, idynamic == 3 branch (Main/mod_bdycod.F90:478-568): rtb, the
// global mean level heights, nztop, tnudge, lowpass_init.  hefc itself is an
// input table (exponential_nudging / relax_coefficients are host-side set-up).
void lowpass_init(World& w);
void setup_bdycon(World& w) {
  const oracle_config& c = w.c; const int kz = c.kz;
  w.rtb = 1.0 / w.x.dtbdys;
  w.gmeanz.assign(kz + 1, 0.0); w.tnudge.assign(kz + 1, 0.0); w.cnudge.assign(kz + 1, 0.0);
  if (w.fcx.empty()) w.fcx.assign(std::max(c.nspgx, 1) + 1, 0.0);
  w.nztop = 0;
  const double zztop = 18000.0;
  if (w.x.mo_top_nudge || w.x.mo_spectral_nudge) {
    const int njcross = (w.r[0].g.band ? c.jx : c.jx - 1), nicross = (w.r[0].g.crm ? c.iy : c.iy - 1);
    const double np = (double)(njcross * nicross);
    for (int k = 1; k <= kz; ++k) {
      double mpmeanz = 0.0;
      for (auto& r : w.r) {   // sumall over ranks
        double meanz = 0.0;
        for (int i = r.g.ice1; i <= r.g.ice2; ++i) for (int j = r.g.jce1; j <= r.g.jce2; ++j)
          meanz = meanz + r.zeta(j, i, k) / np;
        mpmeanz += meanz;
      }
      w.gmeanz[k] = mpmeanz;
      if (w.gmeanz[k] > zztop) w.nztop = w.nztop + 1;
    }
  }
  if (w.x.mo_spectral_nudge) {
    lowpass_init(w);
    for (auto& r : w.r) {
      const Geom& g = r.g;
      r.sx.assign((size_t)(g.ide2 - g.ide1 + 1) * 2 * w.km, 0.0); r.sxg = r.sx;
      r.sy.assign((size_t)(g.jde2 - g.jde1 + 1) * 2 * w.lm, 0.0); r.syg = r.sy;
    }
  }
  if (w.x.mo_top_nudge) {
    for (int k = 1; k <= kz; ++k) {
      if (k <= w.nztop) {
        const double sn = std::sin(0.5 * mathpi * (w.gmeanz[k] - zztop) / (c.mo_h - zztop));
        w.tnudge[k] = sn * sn;
      } else w.tnudge[k] = 0.0;
    }
  }
}

// resolve "qx"/"trac"/"qxten"/"chiten" (4-D, all species) or plain 3-D/2-D names
bool lookup(World& w, Rank& r, const std::string& name, std::vector<FieldInfo>& out) {
  out.clear();
  const int kz = r.g.kz;
  if (name == "qx") { for (auto& a : r.qx) out.push_back({&a, S_CROSS, kz}); return true; }
  if (name == "trac") { for (auto& a : r.trac) out.push_back({&a, S_CROSS, kz}); return true; }
  if (name == "qxten") { for (auto& a : r.qxten) out.push_back({&a, S_CROSS, kz}); return true; }
  if (name == "chiten") { for (auto& a : r.chiten) out.push_back({&a, S_CROSS, kz}); return true; }
  if (name == "chiten0") { for (auto& a : r.chiten0) out.push_back({&a, S_CROSS, kz}); return !r.chiten0.empty(); }
  if (name == "cadvhdiag") { for (auto& a : r.cadvhdiag) out.push_back({&a, S_CROSS, kz}); return !r.cadvhdiag.empty(); }
  if (name == "cbdydiag") { for (auto& a : r.cbdydiag) out.push_back({&a, S_CROSS, kz}); return !r.cbdydiag.empty(); }
  if (name == "chib0") { for (auto& a : r.chib0) out.push_back({&a, S_CROSS, kz}); return !r.chib0.empty(); }
  if (name == "chib1") { for (auto& a : r.chib1) out.push_back({&a, S_CROSS, kz}); return !r.chib1.empty(); }
  auto it = r.reg.find(name);
  if (it == r.reg.end()) return false;
  out.push_back(it->second);
  (void)w;
  return true;
}

void owned_box(const Geom& g, Stag s, int& j1, int& j2, int& i1, int& i2) {
  switch (s) {
    case S_CROSS: j1 = g.jce1; j2 = g.jce2; i1 = g.ice1; i2 = g.ice2; break;
    case S_U: j1 = g.jde1; j2 = g.jde2; i1 = g.ice1; i2 = g.ice2; break;
    case S_V: j1 = g.jce1; j2 = g.jce2; i1 = g.ide1; i2 = g.ide2; break;
    default: j1 = g.jde1; j2 = g.jde2; i1 = g.ide1; i2 = g.ide2; break;
  }
}

// ---------------------------------------------------------------------------
// compute_moloch_static (Main/mod_params.F90:3316-3395), zita levels
// (Main/mod_params.F90:2461-2463, Share/mod_zita.F90:54-84), init_moloch
// [F90:201-308], setup_bdywt (Main/mod_bdycod.F90:4033-4047) with ba%ibnd from
// setup_boundaries (Main/mod_atm_interface.F90:384-532), ffilt
// (Main/mod_init.F90:1008-1026).
// ---------------------------------------------------------------------------
void setup_boundaries(const World& w, Rank& r, bool ldotx, bool ldoty, Arr& ibnd) {
  const Geom& g = r.g; const int jx = g.jx, iy = g.iy;
  const int nsp = w.c.nspgx;  // nspgd == nspgx here
  const int jcx = ldotx ? 0 : 1, icy = ldoty ? 0 : 1;
  for (auto& x : ibnd.d) x = -1.0;
  if (nsp <= 0 || g.crm) return;
  const int igbb1 = 2, igbb2 = nsp - 1;
  int jgbl1 = 2; const int jgbl2 = nsp - 1;
  const int igbt1 = iy - icy - nsp + 2, igbt2 = iy - 1 - icy;
  const int jgbr1 = jx - jcx - nsp + 2; int jgbr2 = jx - 1 - jcx;
  if (g.band) { jgbl1 = 1; jgbr2 = jx - jcx; }
  std::vector<char> bs((size_t)ibnd.d.size(), 0), bn((size_t)ibnd.d.size(), 0);
  auto idx = [&](int j, int i) { return (size_t)((i - ibnd.ilo) * ibnd.sj + (j - ibnd.jlo)); };
  const int i1 = g.ide1, i2 = g.ide2, j1 = g.jde1, j2 = g.jde2;
  if (g.band) {
    for (int i = i1; i <= i2; ++i) if (i >= igbb1 && i <= igbb2) for (int j = j1; j <= j2; ++j) {
      if (j < jgbl1 && j > jgbr2) continue;
      ibnd(j, i) = i - igbb1 + 2; bs[idx(j, i)] = 1;
    }
    for (int i = i1; i <= i2; ++i) if (i >= igbt1 && i <= igbt2) for (int j = j1; j <= j2; ++j) {
      if (j < jgbl1 && j > jgbr2) continue;
      ibnd(j, i) = igbt2 - i + 2; bn[idx(j, i)] = 1;
    }
  } else {
    for (int i = i1; i <= i2; ++i) if (i >= igbb1 && i <= igbb2) for (int j = j1; j <= j2; ++j) {
      if (j >= jgbl1 && j <= jgbr2) {
        if (j <= jgbl2 && i >= j) continue;
        if (j >= jgbr1 && i >= (jgbr2 - j + 2)) continue;
        ibnd(j, i) = i - igbb1 + 2; bs[idx(j, i)] = 1;
      }
    }
    for (int i = i1; i <= i2; ++i) if (i >= igbt1 && i <= igbt2) for (int j = j1; j <= j2; ++j) {
      if (j >= jgbl1 && j <= jgbr2) {
        if (j <= jgbl2 && j <= (igbt2 - i + 2)) continue;
        if (j >= jgbr1 && (igbt2 - i) >= (jgbr2 - j)) continue;
        ibnd(j, i) = igbt2 - i + 2; bn[idx(j, i)] = 1;
      }
    }
    for (int i = i1; i <= i2; ++i) for (int j = j1; j <= j2; ++j) {
      if (bn[idx(j, i)] || bs[idx(j, i)]) continue;
      if (i < igbb1 || i > igbt2) continue;
      if (j >= jgbl1 && j <= jgbl2) ibnd(j, i) = j - jgbl1 + 2;
    }
    for (int i = i1; i <= i2; ++i) for (int j = j1; j <= j2; ++j) {
      if (bn[idx(j, i)] || bs[idx(j, i)]) continue;
      if (i < igbb1 || i > igbt2) continue;
      if (j >= jgbr1 && j <= jgbr2) ibnd(j, i) = jgbr2 - j + 2;
    }
  }
}

void setup_bdywt(const World& w, Rank& r, int j1, int j2, int i1, int i2, Arr& mask, const Arr& ibnd) {
  const int kz = r.g.kz, nsp = w.c.nspgx;
  for (int k = 1; k <= kz; ++k) for (int i = i1; i <= i2; ++i) for (int j = j1; j <= j2; ++j) {
    const int ib = (int)ibnd(j, i);
    if (ib > 0 && nsp > 0) mask(j, i, k) = 1.0 - w.hefc[(size_t)((k - 1) * nsp + (ib - 1))];
    else mask(j, i, k) = 1.0;
  }
}

int setup_static(World& w) {
  const oracle_config& c = w.c; const int kz = c.kz, kzp1 = kz + 1;
  // model_zitaf / model_zitah (Share/mod_zita.F90:54-84); mo_dzita = zita(kz)
  w.zita.assign(kzp1 + 1, 0.0); w.zitah.assign(kz + 1, 0.0);
  { const double dz = c.mo_ztop / (double)kz;
    w.zita[kzp1] = 0.0; w.zita[1] = c.mo_ztop;
    for (int k = kz; k >= 2; --k) w.zita[k] = w.zita[k + 1] + dz;
    w.zitah[kz] = dz * 0.5; w.zitah[1] = c.mo_ztop - dz * 0.5;
    for (int k = kz - 1; k >= 2; --k) w.zitah[k] = w.zitah[k + 1] + dz; }
  w.mo_dzita = w.zita[kz]; w.rdzita = 1.0 / w.mo_dzita;
  w.dx = c.dx; w.rdx = 1.0 / c.dx;  // Main/mod_params.F90:2193-2200
  const double rdx = w.rdx;
  // msf*, ht* ghosts were filled by oracle_set_global (= the exchanges at :3319-3323)
  each(w, [&](Rank& r) {
    const Geom& g = r.g;
    for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jde1; j <= g.jde2; ++j) {
      if (j == 1) r.hx(j, i) = 2.0 * rdx * regrav * r.mu(j, i) * (r.ht(j, i) - r.htu(j, i));
      else r.hx(j, i) = rdx * regrav * r.mu(j, i) * (r.ht(j, i) - r.ht(j - 1, i));
    }
    for (int i = g.ide1; i <= g.ide2; ++i) for (int j = g.jce1; j <= g.jce2; ++j) {
      if (w.lrotllr) {
        if (i == 1) r.hy(j, i) = 2.0 * rdx * regrav * (r.ht(j, i) - r.htv(j, i));
        else r.hy(j, i) = rdx * regrav * (r.ht(j, i) - r.ht(j, i - 1));
      } else {
        if (i == 1) r.hy(j, i) = 2.0 * rdx * regrav * r.mv(j, i) * (r.ht(j, i) - r.htv(j, i));
        else r.hy(j, i) = rdx * regrav * r.mv(j, i) * (r.ht(j, i) - r.ht(j, i - 1));
      }
    }
  });
  exchange(w, [](Rank& r) -> Arr& { return r.hx; }, 1, true, false, BOX_U);
  exchange(w, [](Rank& r) -> Arr& { return r.hy; }, 1, false, true, BOX_V);
  each(w, [&](Rank& r) {
    const Geom& g = r.g;
    for (int k = 1; k <= kz; ++k) for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jce1; j <= g.jce2; ++j) {
      r.zeta(j, i, k) = md_zeta(w.zitah[k], r.ht(j, i), c.mo_ztop, c.mo_h, c.mo_a0);
      r.fmz(j, i, k) = md_fmz(w.zitah[k], r.ht(j, i), c.mo_ztop, c.mo_h, c.mo_a0);
    }
    for (int k = 1; k <= kz; ++k) for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jde1; j <= g.jde2; ++j)
      r.rfmzu(j, i, k) = 1.0 / md_fmz(w.zitah[k], r.htu(j, i), c.mo_ztop, c.mo_h, c.mo_a0);
    for (int k = 1; k <= kz; ++k) for (int i = g.ide1; i <= g.ide2; ++i) for (int j = g.jce1; j <= g.jce2; ++j)
      r.rfmzv(j, i, k) = 1.0 / md_fmz(w.zitah[k], r.htv(j, i), c.mo_ztop, c.mo_h, c.mo_a0);
    for (int k = 1; k <= kzp1; ++k) for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jce1; j <= g.jce2; ++j) {
      r.fmzf(j, i, k) = md_fmz(w.zita[k], r.ht(j, i), c.mo_ztop, c.mo_h, c.mo_a0);
      r.zetaf(j, i, k) = md_zeta(w.zita[k], r.ht(j, i), c.mo_ztop, c.mo_h, c.mo_a0);
    }
  });
  exchange(w, [](Rank& r) -> Arr& { return r.fmz; }, 1, true, true, BOX_CROSS);
  exchange(w, [](Rank& r) -> Arr& { return r.rfmzu; }, 1, true, true, BOX_U);
  exchange(w, [](Rank& r) -> Arr& { return r.rfmzv; }, 1, true, true, BOX_V);
  // init_moloch [F90:256-308]
  each(w, [&](Rank& r) {
    const Geom& g = r.g;
    for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jde1; j <= g.jde2; ++j)
      r.coru(j, i) = eomeg2 * std::sin(r.ulat(j, i) * degrad);
    for (int i = g.ide1; i <= g.ide2; ++i) for (int j = g.jce1; j <= g.jce2; ++j)
      r.corv(j, i) = eomeg2 * std::sin(r.vlat(j, i) * degrad);
    for (int i = g.ide1; i <= g.ide2; ++i) for (int j = g.jde1; j <= g.jde2; ++j) {
      r.mx2(j, i) = r.mx(j, i) * r.mx(j, i);
      r.rmx(j, i) = 1.0 / r.mx(j, i);
      r.rmu(j, i) = 1.0 / r.mu(j, i);
      r.rmv(j, i) = 1.0 / r.mv(j, i);
    }
    for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jce1; j <= g.jce2; ++j) r.w(j, i, 1) = 0.0;
  });
  for (auto f : {&Rank::mx2, &Rank::rmx, &Rank::rmu, &Rank::rmv})
    exchange(w, [f](Rank& r) -> Arr& { return r.*f; }, 1, true, true, BOX_DOT);
  w.gzitak.assign(kzp1 + 1, 0.0); w.gzitakh.assign(kz + 1, 0.0);
  for (int k = 1; k <= kzp1; ++k) w.gzitak[k] = gzita(w.zita[k], c.mo_ztop, c.mo_a0);
  for (int k = 1; k <= kz; ++k) w.gzitakh[k] = gzita(w.zitah[k], c.mo_ztop, c.mo_a0);
  w.xkdamp.assign(kz + 1, 0.0); w.xknu.assign(kz + 1, 0.0);
  { const double numax = 0.125, ddamp = 0.850;
    for (int k = 1; k <= kz; ++k) {
      w.xkdamp[k] = numax * ddamp * (1.0 / ((double)k + 1.0) - 1.0 / ((double)kz + 2.0));
      w.xknu[k] = numax * (0.55 + 0.45 * ((double)(kz - k + 1) - 1.0) / ((double)kz - 1.0));
    } }
  each(w, [&](Rank& r) {
    const Geom& g = r.g;
    setup_boundaries(w, r, false, false, r.ibnd_cr);  // ba_cr
    setup_boundaries(w, r, true, false, r.ibnd_ud);   // ba_ud (Main/mod_atm_interface.F90 allocate: ldotx only)
    setup_boundaries(w, r, false, true, r.ibnd_vd);   // ba_vd
    setup_bdywt(w, r, g.jdi1, g.jdi2, g.ici1, g.ici2, r.bdywtu, r.ibnd_ud);
    setup_bdywt(w, r, g.jci1, g.jci2, g.idi1, g.idi2, r.bdywtv, r.ibnd_vd);
    setup_bdywt(w, r, g.jci1, g.jci2, g.ici1, g.ici2, r.bdywtw, r.ibnd_cr);
  });
  w.dtsec = c.dtsec;
  w.dtstepa = c.dtsec / (double)c.mo_nadv;
  w.dtsound = w.dtstepa / (double)c.mo_nsound;
  // ffilt (Main/mod_init.F90:1008-1026): global mean height of each level
  w.ffilt.assign(kz + 1, 0.0);
  { const int njcross = (w.r[0].g.band ? c.jx : c.jx - 1), nicross = (w.r[0].g.crm ? c.iy : c.iy - 1);
    const double np = (double)(njcross * nicross);
    for (int k = 1; k <= kz; ++k) {
      double gmeanz = 0.0;
      for (auto& r : w.r) {  // sumall over ranks
        double meanz = 0.0;
        for (int i = r.g.ice1; i <= r.g.ice2; ++i) for (int j = r.g.jce1; j <= r.g.jce2; ++j)
          meanz = meanz + r.zeta(j, i, k) / np;
        gmeanz += meanz;
      }
      if (gmeanz < 18000.0) w.ffilt[k] = 0.0;
      else {
        const double zzi = (gmeanz - 18000.0) / (c.mo_ztop - 18000.0);
        const double sn = std::sin(0.5 * mathpi * zzi);
        w.ffilt[k] = mo_zfilt_fac * (sn * sn);
      }
    } }
  if (w.ext_set && w.x.do_bdy) setup_bdycon(w);
  return 0;
}

// paicompute (Main/mod_bdycod.F90:3762-3796) and Main/mod_init.F90:941-953
void temp_to_tvirt(World& w);
int init_state(World& w) {
  const int kz = w.c.kz;
  each(w, [&](Rank& r) {
    const Geom& g = r.g; Arr& q = r.qx[0];
    for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jce1; j <= g.jce2; ++j) {
      double zdelta = r.zeta(j, i, kz) * egrav;
      double tv1 = r.t(j, i, kz) * (1.0 + ep1 * q(j, i, kz));
      double tv2 = r.t(j, i, kz - 1) * (1.0 + ep1 * q(j, i, kz - 1));
      double lrt = (tv2 - tv1) / (r.zeta(j, i, kz - 1) - r.zeta(j, i, kz));
      if (lrt > govcp) lrt = govcp;
      else if (lrt < -0.005) lrt = 0.5 * lrt - 0.5 * lrate;
      const double tv = tv1 - 0.5 * r.zeta(j, i, kz) * lrt;
      const double zz = 1.0 / (rgas * tv);
      const double p = r.ps(j, i) * std::exp(-zdelta * zz);
      double paikp1 = std::pow(p / p00, rovcp);
      r.pai(j, i, kz) = paikp1;
      for (int k = kz - 1; k >= 1; --k) {
        tv1 = r.t(j, i, k) * (1.0 + ep1 * q(j, i, k));
        tv2 = r.t(j, i, k + 1) * (1.0 + ep1 * q(j, i, k + 1));
        const double zb = 2.0 * egrav * w.mo_dzita / (r.fmzf(j, i, k + 1) * cpd) + tv1 - tv2;
        zdelta = std::sqrt(zb * zb + 4.0 * tv2 * tv1);
        paikp1 = -paikp1 / (2.0 * tv2) * (zb - zdelta);
        r.pai(j, i, k) = paikp1;
      }
    }
    for (int k = 1; k <= kz; ++k) for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jce1; j <= g.jce2; ++j) {
      r.p(j, i, k) = std::pow(r.pai(j, i, k), cpovr) * p00;
      r.qsat(j, i, k) = pfwsat(r.t(j, i, k), r.p(j, i, k));
      r.rho(j, i, k) = r.p(j, i, k) / (rgas * r.t(j, i, k));
      r.tvirt(j, i, k) = r.t(j, i, k) * (1.0 + ep1 * q(j, i, k));
      r.tetav(j, i, k) = r.tvirt(j, i, k) / r.pai(j, i, k);
    }
    for (auto& x : r.w.d) x = 0.0;
  });
  exchange(w, [](Rank& r) -> Arr& { return r.pai; }, 1, true, true, BOX_CROSS);
  return 0;
}

// ---------------------------------------------------------------------------
// The hot path
// ---------------------------------------------------------------------------

// reset_tendencies [F90:1044-1083]
void reset_tendencies(World& w) {
  each(w, [&](Rank& r) {
    const Geom& g = r.g; const int kz = g.kz, kzp1 = g.kzp1;
    PAR2 for (int k = 1; k <= kzp1; ++k) for (int i = g.ice1; i <= g.ice2; ++i)
      for (int j = g.jce1; j <= g.jce2; ++j) r.s(j, i, k) = 0.0;
    // jci1ga:jci2ga == interior range extended by one ghost where a neighbour exists
    const int j1 = g.jci1 - g.gl, j2 = g.jci2 + g.gr, i1 = g.ici1 - g.gb, i2 = g.ici2 + g.gt;
    PAR2 for (int k = 1; k <= kz; ++k) for (int i = i1; i <= i2; ++i)
      for (int j = j1; j <= j2; ++j) r.zdiv2(j, i, k) = 0.0;
    PAR2 for (int k = 2; k <= kzp1; ++k) for (int i = g.ice1; i <= g.ice2; ++i)
      for (int j = g.jce1; j <= g.jce2; ++j) r.wwkw(j, i, k) = 0.0;
    PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i)
      for (int j = g.jci1; j <= g.jci2; ++j) { r.tten(j, i, k) = 0.0; r.uten(j, i, k) = 0.0; r.vten(j, i, k) = 0.0; }
    for (auto& a : r.qxten) { PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i)
      for (int j = g.jci1; j <= g.jci2; ++j) a(j, i, k) = 0.0; }
    for (auto& a : r.chiten) { PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i)
      for (int j = g.jci1; j <= g.jci2; ++j) a(j, i, k) = 0.0; }
    if (w.ext_set && w.x.ibltyp == 2) {   // [F90:1071-1075]
      PAR2 for (int k = 1; k <= kzp1; ++k) for (int i = g.ici1; i <= g.ici2; ++i)
        for (int j = g.jci1; j <= g.jci2; ++j) r.tketen(j, i, k) = 0.0;
    }
  });
}

// divergence_damping [F90:738-765]
void divergence_damping(World& w, double dts) {
  const double dxrdt = w.dx / dts;
  exchange(w, [](Rank& r) -> Arr& { return r.zdiv2; }, 1, true, true, BOX_CROSS);
  each(w, [&](Rank& r) {
    const Geom& g = r.g; const int kz = g.kz;
    PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jdi1; j <= g.jdi2; ++j) {
      const double xdam = dxrdt * w.xkdamp[k] * r.mu(j, i);
      r.u(j, i, k) = r.u(j, i, k) + xdam * (r.zdiv2(j, i, k) - r.zdiv2(j - 1, i, k));
    }
    if (w.lrotllr) {
      PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.idi1; i <= g.idi2; ++i) for (int j = g.jci1; j <= g.jci2; ++j) {
        const double xdam = dxrdt * w.xkdamp[k];
        r.v(j, i, k) = r.v(j, i, k) + xdam * (r.zdiv2(j, i, k) - r.zdiv2(j, i - 1, k));
      }
    } else {
      PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.idi1; i <= g.idi2; ++i) for (int j = g.jci1; j <= g.jci2; ++j) {
        const double xdam = dxrdt * w.xkdamp[k] * r.mv(j, i);
        r.v(j, i, k) = r.v(j, i, k) + xdam * (r.zdiv2(j, i, k) - r.zdiv2(j, i - 1, k));
      }
    }
  });
}

// divergence_diffusion [F90:531-543]
void divergence_diffusion(World& w) {
  exchange(w, [](Rank& r) -> Arr& { return r.zdiv2; }, 1, true, true, BOX_CROSS);
  each(w, [&](Rank& r) {
    const Geom& g = r.g; const int kz = g.kz;
    PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j)
      r.laplacian(j, i, k) = (r.zdiv2(j - 1, i, k) + r.zdiv2(j + 1, i, k) + r.zdiv2(j, i - 1, k) +
                              r.zdiv2(j, i + 1, k) - 4.0 * r.zdiv2(j, i, k));
    PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j)
      r.zdiv2(j, i, k) = r.zdiv2(j, i, k) + w.xknu[k] * r.laplacian(j, i, k);
  });
}

// sound [F90:545-736]
void sound(World& w, double dts) {
  const double dtrdx = dts * w.rdx, dtrdy = dts * w.rdx, dtrdz = dts * w.rdzita;
  const double zcs2 = (dtrdz * dtrdz) * rdrcv;
  const int kz = w.c.kz, kzp1 = kz + 1;
  exchange(w, [](Rank& r) -> Arr& { return r.tetav; }, 1, true, true, BOX_CROSS);
  each(w, [&](Rank& r) {
    const Geom& g = r.g;
    PAR2 for (int k = 2; k <= kz; ++k) for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jce1; j <= g.jce2; ++j)
      r.tetavf(j, i, k) = 0.5 * (r.tetav(j, i, k - 1) + r.tetav(j, i, k));
  });
  for (int nsound = 1; nsound <= w.c.mo_nsound; ++nsound) {
    exchange(w, [](Rank& r) -> Arr& { return r.u; }, 1, true, false, BOX_U);
    exchange(w, [](Rank& r) -> Arr& { return r.v; }, 1, false, true, BOX_V);
    each(w, [&](Rank& r) {
      const Geom& g = r.g;
      PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ice1; i <= g.ice2; ++i)
        for (int j = g.jde1; j <= g.jde2; ++j) r.ud(j, i, k) = r.u(j, i, k);
      PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ide1; i <= g.ide2; ++i)
        for (int j = g.jce1; j <= g.jce2; ++j) r.vd(j, i, k) = r.v(j, i, k);
      // partial definition of the generalized vertical velocity [F90:582-597]
      PAR1 for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j) {
        const double zuh = r.u(j, i, kz) * r.hx(j, i) + r.u(j + 1, i, kz) * r.hx(j + 1, i);
        const double zvh = r.v(j, i, kz) * r.hy(j, i) + r.v(j, i + 1, kz) * r.hy(j, i + 1);
        r.s(j, i, kzp1) = -0.5 * (zuh + zvh);
        r.w(j, i, kzp1) = -r.s(j, i, kzp1);
      }
      PAR2 for (int k = 2; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j) {
        const double zuh = (r.u(j, i, k) + r.u(j, i, k - 1)) * r.hx(j, i) +
                           (r.u(j + 1, i, k) + r.u(j + 1, i, k - 1)) * r.hx(j + 1, i);
        const double zvh = (r.v(j, i, k) + r.v(j, i, k - 1)) * r.hy(j, i) +
                           (r.v(j, i + 1, k) + r.v(j, i + 1, k - 1)) * r.hy(j, i + 1);
        r.s(j, i, k) = -0.25 * (zuh + zvh) * w.gzitak[k];
      }
      // Equation 16 [F90:602-618]
      if (w.lrotllr) {
        PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jce1; j <= g.jce2; ++j) {
          const double zum = dtrdx * r.u(j, i, k) * r.rfmzu(j, i, k);
          const double zup = dtrdx * r.u(j + 1, i, k) * r.rfmzu(j + 1, i, k);
          const double zvm = dtrdy * r.v(j, i, k) * r.rfmzv(j, i, k) * r.rmv(j, i);
          const double zvp = dtrdy * r.v(j, i + 1, k) * r.rfmzv(j, i + 1, k) * r.rmv(j, i + 1);
          r.zdiv2(j, i, k) = r.fmz(j, i, k) * r.mx(j, i) * ((zup - zum) + (zvp - zvm));
        }
      } else {
        PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jce1; j <= g.jce2; ++j) {
          const double zum = dtrdx * r.u(j, i, k) * r.rfmzu(j, i, k) * r.rmu(j, i);
          const double zup = dtrdx * r.u(j + 1, i, k) * r.rfmzu(j + 1, i, k) * r.rmu(j + 1, i);
          const double zvm = dtrdy * r.v(j, i, k) * r.rfmzv(j, i, k) * r.rmv(j, i);
          const double zvp = dtrdy * r.v(j, i + 1, k) * r.rfmzv(j, i + 1, k) * r.rmv(j, i + 1);
          r.zdiv2(j, i, k) = r.fmz(j, i, k) * r.mx2(j, i) * ((zup - zum) + (zvp - zvm));
        }
      }
    });
    if (w.do_divdamp) divergence_damping(w, dts);
    if (w.do_divfilter) divergence_diffusion(w);
    each(w, [&](Rank& r) {
      const Geom& g = r.g;
      PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j)
        r.zdiv2(j, i, k) = r.zdiv2(j, i, k) +
                           r.bdywtw(j, i, k) * dtrdz * r.fmz(j, i, k) * (r.s(j, i, k) - r.s(j, i, k + 1));
      // new w (implicit scheme) from Equation 19 [F90:634-664]
      PAR1 for (int i = g.ici1; i <= g.ici2; ++i) {
        for (int k = kz; k >= 2; --k) for (int j = g.jci1; j <= g.jci2; ++j) {
          r.tetavf(j, i, k) = r.tetavf(j, i, k) -
              r.w(j, i, k) * r.fmzf(j, i, k) * dtrdz * (r.tetav(j, i, k - 1) - r.tetav(j, i, k));
          const double zrom1w = cpd * r.tetavf(j, i, k) * r.fmzf(j, i, k);
          double zwexpl = r.w(j, i, k) - zrom1w * dtrdz * (r.pai(j, i, k - 1) - r.pai(j, i, k)) - egrav * dts;
          zwexpl = zwexpl + rdrcv * zrom1w * dtrdz *
                                (r.pai(j, i, k - 1) * r.zdiv2(j, i, k - 1) - r.pai(j, i, k) * r.zdiv2(j, i, k));
          const double zu = zcs2 * r.fmz(j, i, k - 1) * zrom1w * r.pai(j, i, k - 1) + w.ffilt[k];
          const double zd = zcs2 * r.fmz(j, i, k) * zrom1w * r.pai(j, i, k) + w.ffilt[k];
          const double zrapp = 1.0 / (1.0 + zd + zu - zd * r.wwkw(j, i, k + 1));
          r.w(j, i, k) = zrapp * (zwexpl + zd * r.w(j, i, k + 1));
          r.wwkw(j, i, k) = zrapp * zu;
        }
        for (int k = 2; k <= kz; ++k) for (int j = g.jci1; j <= g.jci2; ++j)
          r.w(j, i, k) = r.w(j, i, k) + r.wwkw(j, i, k) * r.w(j, i, k - 1);
      }
      // new Exner function [F90:668-671]
      PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j)
        r.pai(j, i, k) = r.pai(j, i, k) *
            (1.0 - rdrcv * (r.zdiv2(j, i, k) + (dtrdz * r.fmz(j, i, k) * (r.w(j, i, k) - r.w(j, i, k + 1)))));
    });
    exchange(w, [](Rank& r) -> Arr& { return r.pai; }, 1, true, true, BOX_CROSS);
    // horizontal momentum equations [F90:677-721]
    each(w, [&](Rank& r) {
      const Geom& g = r.g;
      PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jdi1; j <= g.jdi2; ++j) {
        const double zcx = dtrdx * r.mu(j, i);
        const double zfz = egrav * dts;
        const double zrom1u = 0.5 * cpd * (r.tetav(j - 1, i, k) + r.tetav(j, i, k));
        const double zcor1u = r.coru(j, i) * dts * r.vd(j, i, k);
        r.u(j, i, k) = r.u(j, i, k) + r.bdywtu(j, i, k) *
            (zcor1u - zfz * r.hx(j, i) * w.gzitakh[k] - zcx * zrom1u * (r.pai(j, i, k) - r.pai(j - 1, i, k)));
      }
      PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.idi1; i <= g.idi2; ++i) for (int j = g.jci1; j <= g.jci2; ++j) {
        const double zcy = w.lrotllr ? dtrdy : dtrdy * r.mv(j, i);
        const double zfz = egrav * dts;
        const double zrom1v = 0.5 * cpd * (r.tetav(j, i - 1, k) + r.tetav(j, i, k));
        const double zcor1v = r.corv(j, i) * dts * r.ud(j, i, k);
        r.v(j, i, k) = r.v(j, i, k) + r.bdywtv(j, i, k) *
            (-zcor1v - zfz * r.hy(j, i) * w.gzitakh[k] - zcy * zrom1v * (r.pai(j, i, k) - r.pai(j, i - 1, k)));
      }
    });
  }  // sound loop
  // complete Equation 10 [F90:728-734]
  each(w, [&](Rank& r) {
    const Geom& g = r.g;
    PAR2 for (int k = 2; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j)
      r.s(j, i, k) = (r.w(j, i, k) + r.s(j, i, k)) * r.fmzf(j, i, k);
    PAR1 for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j) {
      r.s(j, i, 1) = 0.0; r.s(j, i, kzp1) = 0.0;
    }
  });
}

// one vertical WAF pass [F90:868-892 / 896-920]: q -> out (out may alias q only
// through the second-pass form, which reads and writes wz)
void waf_vertical_pass(World& w, Rank& r, Arr& q, double dtrdz) {
  const Geom& g = r.g; const int kz = g.kz, kzm1 = g.kzm1;
  (void)w;
  PAR2 for (int k = 1; k <= kzm1; ++k) for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jce1; j <= g.jce2; ++j) {
    const double zamu = r.s(j, i, k + 1) * dtrdz;
    double is; int k1, k1p1;
    if (zamu >= 0.0) { is = 1.0; k1 = k + 1; k1p1 = k1 + 1; if (k1p1 > kz) k1p1 = kz; }
    else { is = -1.0; k1 = k - 1; k1p1 = k; if (k1 < 1) k1 = 1; }
    const double rr = local_flow_param(q(j, i, k1) - q(j, i, k1p1), q(j, i, k) - q(j, i, k + 1));
    const double b = std::max(0.0, std::min(2.0, std::max(rr, std::min(2.0 * rr, 1.0))));
    const double zphi = is + zamu * b - is * b;
    r.wfw(j, i, k + 1) = 0.5 * r.s(j, i, k + 1) * ((1.0 + zphi) * q(j, i, k + 1) + (1.0 - zphi) * q(j, i, k));
  }
  PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jce1; j <= g.jce2; ++j) {
    const double zrfmu = dtrdz * r.fmz(j, i, k) / r.fmzf(j, i, k);
    const double zrfmd = dtrdz * r.fmz(j, i, k) / r.fmzf(j, i, k + 1);
    const double zdv = (r.s(j, i, k) * zrfmu - r.s(j, i, k + 1) * zrfmd) * q(j, i, k);
    r.wz(j, i, k) = q(j, i, k) - r.wfw(j, i, k) * zrfmu + r.wfw(j, i, k + 1) * zrfmd + zdv;
  }
}

// wafone [F90:838-1042]
void wafone(World& w, const ArrFn& getpp, double dta) {
  const double dtrdx = dta * w.rdx, dtrdy = dta * w.rdx;
  double dtrdz = dta * w.rdzita;
  const bool do_vadvtwice = true;
  if (do_vadvtwice) dtrdz = 0.5 * dtrdz;
  const int kz = w.c.kz, kzp1 = kz + 1;
  each(w, [&](Rank& r) {
    const Geom& g = r.g; Arr& pp = getpp(r);
    for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jce1; j <= g.jce2; ++j) {
      r.wfw(j, i, 1) = 0.0; r.wfw(j, i, kzp1) = 0.0;
    }
    waf_vertical_pass(w, r, pp, dtrdz);
    if (do_vadvtwice) waf_vertical_pass(w, r, r.wz, dtrdz);
  });
  exchange(w, [](Rank& r) -> Arr& { return r.wz; }, 2, false, true, BOX_CROSS);
  const bool rot = w.lrotllr;
  // Meridional advection [F90:929-953 / 987-1010]
  each(w, [&](Rank& r) {
    const Geom& g = r.g; Arr& pp = getpp(r);
    PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ice2ga(); ++i) for (int j = g.jce1; j <= g.jce2; ++j) {
      const double zamu = rot ? r.v(j, i, k) * dtrdy : r.v(j, i, k) * r.mv(j, i) * dtrdy;
      double is; int ih;
      if (zamu > 0.0) { is = 1.0; ih = i - 1; } else { is = -1.0; ih = std::min(i + 1, g.imax); }
      const int ihm1 = std::max(ih - 1, g.imin);
      const double rr = local_flow_param(r.wz(j, ih, k) - r.wz(j, ihm1, k), r.wz(j, i, k) - r.wz(j, i - 1, k));
      const double b = std::max(0.0, std::min(2.0, std::max(rr, std::min(2.0 * rr, 1.0))));
      const double zphi = is + zamu * b - is * b;
      r.zpby(j, i, k) = 0.5 * r.v(j, i, k) * ((1.0 + zphi) * r.wz(j, i - 1, k) + (1.0 - zphi) * r.wz(j, i, k));
    }
    if (rot) {
      PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jce1; j <= g.jce2; ++j) {
        const double zhxvtn = dtrdy * r.rmv(j, i + 1) * r.mx(j, i);
        const double zhxvts = dtrdy * r.rmv(j, i) * r.mx(j, i);
        const double zrfmn = zhxvtn * r.fmz(j, i, k) * r.rfmzv(j, i + 1, k);
        const double zrfms = zhxvts * r.fmz(j, i, k) * r.rfmzv(j, i, k);
        const double zdv = (r.v(j, i + 1, k) * zrfmn - r.v(j, i, k) * zrfms) * pp(j, i, k);
        r.p0(j, i, k) = r.wz(j, i, k) + (r.zpby(j, i, k) * zrfms - r.zpby(j, i + 1, k) * zrfmn + zdv);
      }
    } else {
      PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jce1; j <= g.jce2; ++j) {
        const double zrfmn = dtrdy * r.fmz(j, i, k) * r.rfmzu(j, i + 1, k);  // sic: rfmzu [F90:1004-1005]
        const double zrfms = dtrdy * r.fmz(j, i, k) * r.rfmzu(j, i, k);
        const double zdv = (r.v(j, i + 1, k) * r.rmv(j, i + 1) * zrfmn - r.v(j, i, k) * r.rmv(j, i) * zrfms) * pp(j, i, k);
        r.p0(j, i, k) = r.wz(j, i, k) + r.mx2(j, i) * (r.zpby(j, i, k) * zrfms - r.zpby(j, i + 1, k) * zrfmn + zdv);
      }
    }
  });
  exchange(w, [](Rank& r) -> Arr& { return r.p0; }, 2, true, false, BOX_P0);
  // Zonal advection [F90:959-982 / 1015-1038]
  each(w, [&](Rank& r) {
    const Geom& g = r.g; Arr& pp = getpp(r);
    PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jce2ga(); ++j) {
      const double zamu = r.u(j, i, k) * r.mu(j, i) * dtrdx;
      double is; int jh;
      if (zamu > 0.0) { is = 1.0; jh = j - 1; } else { is = -1.0; jh = std::min(j + 1, g.jmax); }
      const int jhm1 = std::max(jh - 1, g.jmin);
      const double rr = local_flow_param(r.p0(jh, i, k) - r.p0(jhm1, i, k), r.p0(j, i, k) - r.p0(j - 1, i, k));
      const double b = std::max(0.0, std::min(2.0, std::max(rr, std::min(2.0 * rr, 1.0))));
      const double zphi = is + zamu * b - is * b;
      r.zpbw(j, i, k) = 0.5 * r.u(j, i, k) * ((1.0 + zphi) * r.p0(j - 1, i, k) + (1.0 - zphi) * r.p0(j, i, k));
    }
    if (rot) {
      PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j) {
        const double zcostx = dtrdx * r.mx(j, i);
        const double zrfmw = zcostx * r.fmz(j, i, k) * r.rfmzu(j, i, k);
        const double zrfme = zcostx * r.fmz(j, i, k) * r.rfmzu(j + 1, i, k);
        const double zdv = (r.u(j + 1, i, k) * zrfme - r.u(j, i, k) * zrfmw) * pp(j, i, k);
        pp(j, i, k) = r.p0(j, i, k) + r.zpbw(j, i, k) * zrfmw - r.zpbw(j + 1, i, k) * zrfme + zdv;
      }
    } else {
      PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j) {
        const double zrfmw = dtrdx * r.fmz(j, i, k) * r.rfmzu(j, i, k);
        const double zrfme = dtrdx * r.fmz(j, i, k) * r.rfmzu(j + 1, i, k);
        const double zdv = (r.u(j + 1, i, k) * r.rmu(j + 1, i) * zrfme - r.u(j, i, k) * r.rmu(j, i) * zrfmw) * pp(j, i, k);
        pp(j, i, k) = r.p0(j, i, k) + r.mx2(j, i) * (r.zpbw(j, i, k) * zrfmw - r.zpbw(j + 1, i, k) * zrfme + zdv);
      }
    }
  });
}

// uvstagtouvx [F90:1524-1569]
void uvstagtouvx(World& w) {
  const int kz = w.c.kz;
  exchange(w, [](Rank& r) -> Arr& { return r.u; }, 2, true, false, BOX_U);
  exchange(w, [](Rank& r) -> Arr& { return r.v; }, 2, false, true, BOX_V);
  each(w, [&](Rank& r) {
    const Geom& g = r.g;
    PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jci1; j <= g.jci2; ++j)
      r.ux(j, i, k) = 0.5625 * (r.u(j + 1, i, k) + r.u(j, i, k)) - 0.0625 * (r.u(j + 2, i, k) + r.u(j - 1, i, k));
    if (g.bl) for (int k = 1; k <= kz; ++k) for (int i = g.ice1; i <= g.ice2; ++i)
      r.ux(g.jce1, i, k) = 0.5 * (r.u(g.jde1, i, k) + r.u(g.jdi1, i, k));
    if (g.br) for (int k = 1; k <= kz; ++k) for (int i = g.ice1; i <= g.ice2; ++i)
      r.ux(g.jce2, i, k) = 0.5 * (r.u(g.jde2, i, k) + r.u(g.jdi2, i, k));
    PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jce1; j <= g.jce2; ++j)
      r.vx(j, i, k) = 0.5625 * (r.v(j, i + 1, k) + r.v(j, i, k)) - 0.0625 * (r.v(j, i + 2, k) + r.v(j, i - 1, k));
    if (g.bb) for (int k = 1; k <= kz; ++k) for (int j = g.jce1; j <= g.jce2; ++j)
      r.vx(j, g.ice1, k) = 0.5 * (r.v(j, g.ide1, k) + r.v(j, g.idi1, k));
    if (g.bt) for (int k = 1; k <= kz; ++k) for (int j = g.jce1; j <= g.jce2; ++j)
      r.vx(j, g.ice2, k) = 0.5 * (r.v(j, g.ide2, k) + r.v(j, g.idi2, k));
  });
}

// uvxtouvstag [F90:1477-1522]
void uvxtouvstag(World& w) {
  const int kz = w.c.kz;
  exchange(w, [](Rank& r) -> Arr& { return r.ux; }, 2, true, false, BOX_CROSS);
  exchange(w, [](Rank& r) -> Arr& { return r.vx; }, 2, false, true, BOX_CROSS);
  each(w, [&](Rank& r) {
    const Geom& g = r.g;
    PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jdii1; j <= g.jdii2; ++j)
      r.u(j, i, k) = 0.5625 * (r.ux(j, i, k) + r.ux(j - 1, i, k)) - 0.0625 * (r.ux(j + 1, i, k) + r.ux(j - 2, i, k));
    if (g.br) for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i)
      r.u(g.jdi2, i, k) = 0.5 * (r.ux(g.jci2, i, k) + r.ux(g.jce2, i, k));
    if (g.bl) for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i)
      r.u(g.jdi1, i, k) = 0.5 * (r.ux(g.jci1, i, k) + r.ux(g.jce1, i, k));
    PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.idii1; i <= g.idii2; ++i) for (int j = g.jci1; j <= g.jci2; ++j)
      r.v(j, i, k) = 0.5625 * (r.vx(j, i, k) + r.vx(j, i - 1, k)) - 0.0625 * (r.vx(j, i + 1, k) + r.vx(j, i - 2, k));
    if (g.bt) for (int k = 1; k <= kz; ++k) for (int j = g.jci1; j <= g.jci2; ++j)
      r.v(j, g.idi2, k) = 0.5 * (r.vx(j, g.ici2, k) + r.vx(j, g.ice2, k));
    if (g.bb) for (int k = 1; k <= kz; ++k) for (int j = g.jci1; j <= g.jci2; ++j)
      r.v(j, g.idi1, k) = 0.5 * (r.vx(j, g.ici1, k) + r.vx(j, g.ice1, k));
  });
}

// zstagtoh [F90:1445-1459], htozstag [F90:1461-1475]
void zstagtoh(World& w, Arr Rank::*fl, Arr Rank::*hl) {
  const int kz = w.c.kz, kzm1 = kz - 1, kzp1 = kz + 1;
  each(w, [&](Rank& r) {
    const Geom& g = r.g; Arr& f = r.*fl; Arr& h = r.*hl;
    PAR2 for (int k = 2; k <= kzm1; ++k) for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jce1; j <= g.jce2; ++j)
      h(j, i, k) = 0.5625 * (f(j, i, k + 1) + f(j, i, k)) - 0.0625 * (f(j, i, k + 2) + f(j, i, k - 1));
    PAR1 for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jce1; j <= g.jce2; ++j) {
      h(j, i, 1) = 0.5 * (f(j, i, 2) + f(j, i, 1));
      h(j, i, kz) = 0.5 * (f(j, i, kzp1) + f(j, i, kz));
    }
  });
}
void htozstag(World& w, Arr Rank::*hl, Arr Rank::*fl) {
  const int kz = w.c.kz, kzm1 = kz - 1;
  each(w, [&](Rank& r) {
    const Geom& g = r.g; Arr& f = r.*fl; Arr& h = r.*hl;
    PAR2 for (int k = 3; k <= kzm1; ++k) for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jce1; j <= g.jce2; ++j)
      f(j, i, k) = 0.5625 * (h(j, i, k) + h(j, i, k - 1)) - 0.0625 * (h(j, i, k + 1) + h(j, i, k - 2));
    PAR1 for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jce1; j <= g.jce2; ++j) {
      f(j, i, 2) = 0.5 * (h(j, i, 2) + h(j, i, 1));
      f(j, i, kz) = 0.5 * (h(j, i, kz) + h(j, i, kzm1));
    }
  });
}

// advection [F90:767-836]
void advection(World& w, double dta) {
  const int kz = w.c.kz;
  const bool tke = w.ext_set && w.x.ibltyp == 2;
  uvstagtouvx(w);
  zstagtoh(w, &Rank::w, &Rank::wx);
  if (tke) zstagtoh(w, &Rank::tke, &Rank::tkex);   // [F90:782-784]
  wafone(w, [](Rank& r) -> Arr& { return r.tetav; }, dta);
  wafone(w, [](Rank& r) -> Arr& { return r.pai; }, dta);
  wafone(w, [](Rank& r) -> Arr& { return r.ux; }, dta);
  wafone(w, [](Rank& r) -> Arr& { return r.vx; }, dta);
  wafone(w, [](Rank& r) -> Arr& { return r.wx; }, dta);
  wafone(w, [](Rank& r) -> Arr& { return r.qx[0]; }, dta);
  if (w.ipptls > 0)
    for (int n = w.iqfrst; n <= w.nqx; ++n) wafone(w, [n](Rank& r) -> Arr& { return r.qx[n - 1]; }, dta);
  if (tke) wafone(w, [](Rank& r) -> Arr& { return r.tkex; }, dta);   // [F90:799-801]
  for (int n = 1; n <= w.ntr; ++n) wafone(w, [n](Rank& r) -> Arr& { return r.trac[n - 1]; }, dta);
  // curvature terms [F90:811-825]
  each(w, [&](Rank& r) {
    const Geom& g = r.g;
    if (w.lrotllr) {
      PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j) {
        const double dlat = degrad * 0.5 * (w.rlat[i] + w.rlat[i + 1]);
        const double tanx = std::sin(dlat) * r.mx(j, i) * rearthrad;
        r.ux(j, i, k) = r.ux(j, i, k) + r.ux(j, i, k) * r.vx(j, i, k) * tanx * dta;
        r.vx(j, i, k) = r.vx(j, i, k) - r.ux(j, i, k) * r.ux(j, i, k) * tanx * dta;
      }
    } else {
      PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j) {
        const double tanx = (r.mu(j - 1, i) - r.mu(j, i)) * w.rdx;
        const double tany = (r.mv(j, i - 1) - r.mv(j, i)) * w.rdx;
        r.ux(j, i, k) = r.ux(j, i, k) + r.ux(j, i, k) * r.vx(j, i, k) * tanx * dta;
        r.vx(j, i, k) = r.vx(j, i, k) - r.ux(j, i, k) * r.ux(j, i, k) * tany * dta;
      }
    }
  });
  uvxtouvstag(w);
  htozstag(w, &Rank::wx, &Rank::w);
  if (tke) htozstag(w, &Rank::tkex, &Rank::tke);   // [F90:832-834]
}

// temp_to_tvirt [F90:1608-1627], tvirt_to_temp [F90:1629-1648]
inline double moist_factor(const World& w, Rank& r, int j, int i, int k) {
  if (w.ipptls > 0) {
    if (w.ipptls > 1)
      return 1.0 + ep1 * r.qx[0](j, i, k) - r.qx[1](j, i, k) - r.qx[2](j, i, k) - r.qx[3](j, i, k) - r.qx[4](j, i, k);
    return 1.0 + ep1 * r.qx[0](j, i, k) - r.qx[1](j, i, k);
  }
  return 1.0 + ep1 * r.qx[0](j, i, k);
}
void temp_to_tvirt(World& w) {
  const int kz = w.c.kz;
  each(w, [&](Rank& r) {
    const Geom& g = r.g;
    PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jce1; j <= g.jce2; ++j)
      r.tvirt(j, i, k) = r.t(j, i, k) * moist_factor(w, r, j, i, k);
  });
}
void tvirt_to_temp(World& w) {
  const int kz = w.c.kz;
  each(w, [&](Rank& r) {
    const Geom& g = r.g;
    PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j)
      r.t(j, i, k) = r.tvirt(j, i, k) / moist_factor(w, r, j, i, k);
  });
}

// tendency diagnostics: snapshots [F90:1092-1103, 455-466] and differences [F90:1127-1139, 508-519]
void diag_snapshot(World& w) {
  if (!w.ext_set) return;
  const int kz = w.c.kz;
  each(w, [&](Rank& r) {
    const Geom& g = r.g;
    if (w.x.idiag > 0) {
      PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j) {
        r.ten0(j, i, k) = r.t(j, i, k); r.qen0(j, i, k) = r.qx[0](j, i, k);
      }
    }
    for (size_t n = 0; n < r.chiten0.size(); ++n)
      for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j)
        r.chiten0[n](j, i, k) = r.trac[n](j, i, k);
  });
}
void diag_difference(World& w, bool bdy) {
  if (!w.ext_set) return;
  const int kz = w.c.kz; const double rdt = 1.0 / w.dtsec;
  each(w, [&](Rank& r) {
    const Geom& g = r.g;
    if (w.x.idiag > 0) {
      Arr& dt_ = bdy ? r.tdiag_bdy : r.tdiag_adh; Arr& dq_ = bdy ? r.qdiag_bdy : r.qdiag_adh;
      PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j) {
        dt_(j, i, k) = (r.t(j, i, k) - r.ten0(j, i, k)) * rdt;
        dq_(j, i, k) = (r.qx[0](j, i, k) - r.qen0(j, i, k)) * rdt;
      }
    }
    for (size_t n = 0; n < r.chiten0.size(); ++n) {
      Arr& dc = bdy ? r.cbdydiag[n] : r.cadvhdiag[n];
      for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j)
        dc(j, i, k) = (r.trac[n](j, i, k) - r.chiten0[n](j, i, k)) * rdt;
    }
  });
}

// dynamical_core [F90:1085-1141]
void dynamical_core(World& w) {
  diag_snapshot(w);
  const int kz = w.c.kz;
  for (int nadv = 1; nadv <= w.c.mo_nadv; ++nadv) {
    sound(w, w.dtsound);
    advection(w, w.dtstepa);
  }
  each(w, [&](Rank& r) {
    const Geom& g = r.g;
    PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jce1; j <= g.jce2; ++j)
      r.tvirt(j, i, k) = r.tetav(j, i, k) * r.pai(j, i, k);
  });
  tvirt_to_temp(w);
  diag_difference(w, false);
}

// moloch [F90:348-354] + extrapolate_surface_pressure [F90:1592-1606]
void diagnostics(World& w) {
  const int kz = w.c.kz, kzm1 = kz - 1;
  each(w, [&](Rank& r) {
    const Geom& g = r.g;
    PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jce1; j <= g.jce2; ++j) {
      r.p(j, i, k) = std::pow(r.pai(j, i, k), cpovr) * p00;
      r.rho(j, i, k) = r.p(j, i, k) / (rgas * r.t(j, i, k));
      r.qsat(j, i, k) = pfwsat(r.t(j, i, k), r.p(j, i, k));
    }
    PAR1 for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j) {
      double lrt = (r.tvirt(j, i, kzm1) - r.tvirt(j, i, kz)) / (r.zeta(j, i, kzm1) - r.zeta(j, i, kz));
      if (lrt < -govcp) lrt = -govcp;
      else if (lrt > -0.005) lrt = 0.65 * lrt - 0.35 * lrate;
      const double tv = r.tvirt(j, i, kz) - lrt * 0.5 * r.zeta(j, i, kz);
      r.ps(j, i) = r.p(j, i, kz) * std::exp(govr * r.zeta(j, i, kz) / tv);
    }
  });
}

// status_update [F90:1403-1443]
void status_update(World& w, double dtinc) {
  const int kz = w.c.kz;
  each(w, [&](Rank& r) {
    const Geom& g = r.g;
    PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j) {
      r.t(j, i, k) = r.t(j, i, k) + dtinc * r.tten(j, i, k);
      r.ux(j, i, k) = r.ux(j, i, k) + dtinc * r.uten(j, i, k);
      r.vx(j, i, k) = r.vx(j, i, k) + dtinc * r.vten(j, i, k);
    }
    for (int n = 0; n < w.nqx; ++n) {
      Arr& q = r.qx[n]; Arr& qt = r.qxten[n];
      PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j) {
        q(j, i, k) = q(j, i, k) + dtinc * qt(j, i, k);
        if (q(j, i, k) < qxcheckval[n]) q(j, i, k) = qxzeroval[n];
      }
    }
    if (w.ext_set && w.x.ibltyp == 2) {   // [F90:1419-1424]
      const double tkemin = w.x.tkemin;
      PAR2 for (int k = 1; k <= kz + 1; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j) {
        r.tke(j, i, k) = r.tke(j, i, k) + dtinc * r.tketen(j, i, k);
        if (r.tke(j, i, k) < tkemin) r.tke(j, i, k) = tkemin;
      }
    }
    for (int n = 0; n < w.ntr; ++n) {
      Arr& q = r.trac[n]; Arr& qt = r.chiten[n];
      PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j) {
        q(j, i, k) = q(j, i, k) + dtinc * qt(j, i, k);
        if (q(j, i, k) < 0.0) q(j, i, k) = 0.0;
      }
    }
  });
  temp_to_tvirt(w);
  each(w, [&](Rank& r) {
    const Geom& g = r.g;
    PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jce1; j <= g.jce2; ++j) {
      r.tetav(j, i, k) = r.tvirt(j, i, k) / r.pai(j, i, k);
      r.rho(j, i, k) = r.p(j, i, k) / (rgas * r.t(j, i, k));
      r.qsat(j, i, k) = pfwsat(r.t(j, i, k), r.p(j, i, k));
    }
  });
  uvxtouvstag(w);
}

// ---------------------------------------------------------------------------
// Lateral boundary: `boundary` [F90:448-529] and what it calls
// ---------------------------------------------------------------------------

// bdyval, MOLOCH branch (Main/mod_bdycod.F90:1618-1875) and the closing
// xbctime = xbctime + dtsec (:2653).  The SST update at :2620-2651 acts on
// sfs%tg (surface model) and is outside the dycore path.
void chem_bdyval(World& w);
void bdyval(World& w) {
  const double x1 = (w.xbctime + w.dtsec) * w.rtb;
  const double x0 = 1.0 - x1;
  const int kz = w.c.kz, nqx = w.nqx, iqfrst = w.iqfrst;
  const bool pqc = w.x.present_qc != 0, pqi = w.x.present_qi != 0;
  const bool tke = w.x.ibltyp == 2;
  const double tkemin = w.x.tkemin;
  auto skip = [&](int n) { return (pqc && n == 2) || (pqi && n == 3); };
  each(w, [&](Rank& r) {
    const Geom& g = r.g;
    Arr& qv = r.qx[0];
    auto lin = [&](const Arr& b0, const Arr& b1, int j, int i, int k) { return x0 * b0(j, i, k) + x1 * b1(j, i, k); };
    // west / east: corners excluded (:1641-1702, :1706-1767)
    for (int side = 0; side < 2; ++side) {
      if (side == 0 ? !g.bl : !g.br) continue;
      const int jd = side == 0 ? g.jde1 : g.jde2, jc = side == 0 ? g.jce1 : g.jce2;
      const int jin = side == 0 ? g.jci1 : g.jci2;
      const double sgn = side == 0 ? 1.0 : -1.0;   // inflow: u > 0 (west), u < 0 (east)
      for (int i = g.ici1; i <= g.ici2; ++i) r.ps(jc, i) = x0 * r.xpsb0(jc, i) + x1 * r.xpsb1(jc, i);
      for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) r.u(jd, i, k) = lin(r.dub0, r.dub1, jd, i, k);
      for (int k = 1; k <= kz; ++k) for (int i = g.idi1; i <= g.idi2; ++i) r.v(jc, i, k) = lin(r.dvb0, r.dvb1, jc, i, k);
      for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) {
        r.t(jc, i, k) = lin(r.xtb0, r.xtb1, jc, i, k);
        r.pai(jc, i, k) = lin(r.xpaib0, r.xpaib1, jc, i, k);
        qv(jc, i, k) = lin(r.xqb0, r.xqb1, jc, i, k);
      }
      if (pqc) for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i)
        r.qx[1](jc, i, k) = lin(r.xlb0, r.xlb1, jc, i, k);
      if (pqi && w.ipptls > 1) for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i)
        r.qx[2](jc, i, k) = lin(r.xib0, r.xib1, jc, i, k);
      for (int n = iqfrst; n <= nqx; ++n) {
        if (skip(n)) continue;
        Arr& q = r.qx[n - 1];
        for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) {
          const double qxint = q(jin, i, k);
          if (sgn * r.u(jd, i, k) > 0.0) q(jc, i, k) = qxzeroval[n - 1]; else q(jc, i, k) = qxint;
        }
      }
      for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) {
        if (sgn * r.u(jd, i, k) > 0.0) r.w(jc, i, k) = 0.0; else r.w(jc, i, k) = r.w(jin, i, k);
      }
      if (tke) {
        for (int i = g.ici1; i <= g.ici2; ++i) r.tke(jc, i, 1) = tkemin;
        for (int k = 2; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) {
          if (sgn * (r.u(jd, i, k) + r.u(jd, i, k - 1)) > 0.0) r.tke(jc, i, k + 1) = tkemin;
          else r.tke(jc, i, k + 1) = r.tke(jin, i, k + 1);
        }
      }
    }
    // south / north: corners included (:1771-1832, :1836-1897)
    for (int side = 0; side < 2; ++side) {
      if (side == 0 ? !g.bb : !g.bt) continue;
      const int id = side == 0 ? g.ide1 : g.ide2, ic = side == 0 ? g.ice1 : g.ice2;
      const int iin = side == 0 ? g.ici1 : g.ici2;
      const double sgn = side == 0 ? 1.0 : -1.0;
      for (int j = g.jce1; j <= g.jce2; ++j) r.ps(j, ic) = x0 * r.xpsb0(j, ic) + x1 * r.xpsb1(j, ic);
      for (int k = 1; k <= kz; ++k) for (int j = g.jde1; j <= g.jde2; ++j) r.u(j, ic, k) = lin(r.dub0, r.dub1, j, ic, k);
      for (int k = 1; k <= kz; ++k) for (int j = g.jce1; j <= g.jce2; ++j) r.v(j, id, k) = lin(r.dvb0, r.dvb1, j, id, k);
      for (int k = 1; k <= kz; ++k) for (int j = g.jce1; j <= g.jce2; ++j) {
        r.t(j, ic, k) = lin(r.xtb0, r.xtb1, j, ic, k);
        r.pai(j, ic, k) = lin(r.xpaib0, r.xpaib1, j, ic, k);
        qv(j, ic, k) = lin(r.xqb0, r.xqb1, j, ic, k);
      }
      if (pqc) for (int k = 1; k <= kz; ++k) for (int j = g.jce1; j <= g.jce2; ++j)
        r.qx[1](j, ic, k) = lin(r.xlb0, r.xlb1, j, ic, k);
      if (pqi && w.ipptls > 1) for (int k = 1; k <= kz; ++k) for (int j = g.jce1; j <= g.jce2; ++j)
        r.qx[2](j, ic, k) = lin(r.xib0, r.xib1, j, ic, k);
      for (int n = iqfrst; n <= nqx; ++n) {
        if (skip(n)) continue;
        Arr& q = r.qx[n - 1];
        for (int k = 1; k <= kz; ++k) for (int j = g.jce1; j <= g.jce2; ++j) {
          const double qxint = q(j, iin, k);
          if (sgn * r.v(j, id, k) > 0.0) q(j, ic, k) = qxzeroval[n - 1]; else q(j, ic, k) = qxint;
        }
      }
      for (int k = 1; k <= kz; ++k) for (int j = g.jce1; j <= g.jce2; ++j) {
        if (sgn * r.v(j, id, k) > 0.0) r.w(j, ic, k) = 0.0; else r.w(j, ic, k) = r.w(j, iin, k);
      }
      if (tke) {
        for (int j = g.jce1; j <= g.jce2; ++j) r.tke(j, ic, 1) = tkemin;
        for (int k = 2; k <= kz; ++k) for (int j = g.jce1; j <= g.jce2; ++j) {
          if (sgn * (r.v(j, id, k) + r.v(j, id, k - 1)) > 0.0) r.tke(j, ic, k + 1) = tkemin;
          else r.tke(j, ic, k + 1) = r.tke(j, iin, k + 1);
        }
      }
    }
  });
  if (w.x.ichem == 1 && w.ntr > 0) chem_bdyval(w);
  w.xbctime = w.xbctime + w.dtsec;   // :2653
}

// chem_bdyval_uncoupled (Main/chemlib/mod_che_bdyco.F90:391-535)
void chem_bdyval(World& w) {
  const int kz = w.c.kz;
  if (w.x.ichebdy == 0) {
    each(w, [&](Rank& r) {
      const Geom& g = r.g;
      for (int n = 0; n < w.ntr; ++n) {
        Arr& c = r.trac[n];
        if (g.bl) for (int k = 1; k <= kz; ++k) for (int i = g.ice1; i <= g.ice2; ++i) {
          const double trint = c(g.jci1, i, k), windavg = r.u(g.jde1, i, k) - r.u(g.jdi1, i, k);
          c(g.jce1, i, k) = (windavg < 0.0) ? trint : 0.0;
        }
        if (g.br) for (int k = 1; k <= kz; ++k) for (int i = g.ice1; i <= g.ice2; ++i) {
          const double trint = c(g.jci2, i, k), windavg = r.u(g.jde2, i, k) - r.u(g.jdi2, i, k);
          c(g.jce2, i, k) = (windavg > 0.0) ? trint : 0.0;
        }
      }
      for (int n = 0; n < w.ntr; ++n) {
        Arr& c = r.trac[n];
        if (g.bb) for (int k = 1; k <= kz; ++k) for (int j = g.jci1; j <= g.jci2; ++j) {
          const double trint = c(j, g.ici1, k), windavg = r.v(j, g.ide1, k) - r.v(j, g.idi1, k);
          c(j, g.ice1, k) = (windavg < 0.0) ? trint : 0.0;
        }
        if (g.bt) for (int k = 1; k <= kz; ++k) for (int j = g.jci1; j <= g.jci2; ++j) {
          const double trint = c(j, g.ici2, k), windavg = r.v(j, g.ide2, k) - r.v(j, g.idi2, k);
          c(j, g.ice2, k) = (windavg > 0.0) ? trint : 0.0;
        }
      }
    });
    return;
  }
  // time-dependent boundary values (:507-531); note the division by dtbdys
  const double x1 = (w.xbctime + w.dtsec) / w.x.dtbdys;
  const double x0 = 1.0 - x1;
  each(w, [&](Rank& r) {
    const Geom& g = r.g;
    for (int n = 0; n < w.ntr; ++n) {
      Arr& c = r.trac[n]; const Arr& b0 = r.chib0[n]; const Arr& b1 = r.chib1[n];
      if (g.bl) for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i)
        c(g.jce1, i, k) = x0 * b0(g.jce1, i, k) + x1 * b1(g.jce1, i, k);
      if (g.br) for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i)
        c(g.jce2, i, k) = x0 * b0(g.jce2, i, k) + x1 * b1(g.jce2, i, k);
      if (g.bb) for (int k = 1; k <= kz; ++k) for (int j = g.jce1; j <= g.jce2; ++j)
        c(j, g.ice1, k) = x0 * b0(j, g.ice1, k) + x1 * b1(j, g.ice1, k);
      if (g.bt) for (int k = 1; k <= kz; ++k) for (int j = g.jce1; j <= g.jce2; ++j)
        c(j, g.ice2, k) = x0 * b0(j, g.ice2, k) + x1 * b1(j, g.ice2, k);
    }
  });
}

// motopnudge (Main/mod_bdycod.F90:4049-4081), called with wfac = 0, dta = dtsec
void motopnudge(World& w, double wfac, double dta) {
  const double trtau = 1.0 / (2.0 * 3600.0), wrtau = 1.0 / (3600.0 / 4.0);
  const double x1 = (w.xbctime + w.dtsec) * w.rtb, x0 = 1.0 - x1;
  const int nztop = w.nztop;
  each(w, [&](Rank& r) {
    const Geom& g = r.g;
    for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j) {
      const double fext = wfac * r.w(j, i, 2);
      const double xf = wrtau * dta;
      r.w(j, i, 2) = (1.0 - xf) * r.w(j, i, 2) + xf * fext;
    }
    PAR2 for (int k = 1; k <= nztop; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j) {
      const double fext = x0 * r.xtb0(j, i, k) + x1 * r.xtb1(j, i, k);
      const double xf = w.tnudge[k] * trtau * dta;
      r.t(j, i, k) = (1.0 - xf) * r.t(j, i, k) + xf * fext;
    }
    PAR2 for (int k = 1; k <= nztop; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jdi1; j <= g.jdi2; ++j) {
      const double fext = x0 * r.dub0(j, i, k) + x1 * r.dub1(j, i, k);
      const double xf = w.tnudge[k] * trtau * dta;
      r.u(j, i, k) = (1.0 - xf) * r.u(j, i, k) + xf * fext;
    }
    PAR2 for (int k = 1; k <= nztop; ++k) for (int i = g.idi1; i <= g.idi2; ++i) for (int j = g.jci1; j <= g.jci2; ++j) {
      const double fext = x0 * r.dvb0(j, i, k) + x1 * r.dvb1(j, i, k);
      const double xf = w.tnudge[k] * trtau * dta;
      r.v(j, i, k) = (1.0 - xf) * r.v(j, i, k) + xf * fext;
    }
  });
}

// morelax_external (Main/mod_bdycod.F90:3995-4031)
void morelax_external(World& w, int stag, const ArrFn& getf, const ArrFn& getb0, const ArrFn& getb1) {
  const double x1 = (w.xbctime + w.dtsec) * w.rtb, x0 = 1.0 - x1;
  const int kz = w.c.kz, nsp = w.c.nspgx;
  if (nsp <= 0) return;   // ba%havebound is false everywhere
  each(w, [&](Rank& r) {
    const Geom& g = r.g;
    int j1 = g.jci1, j2 = g.jci2, i1 = g.ici1, i2 = g.ici2; const Arr* ibnd = &r.ibnd_cr;
    if (stag == S_U) { j1 = g.jdi1; j2 = g.jdi2; ibnd = &r.ibnd_ud; }
    if (stag == S_V) { i1 = g.idi1; i2 = g.idi2; ibnd = &r.ibnd_vd; }
    Arr& f = getf(r); const Arr& b0 = getb0(r); const Arr& b1 = getb1(r);
    PAR2 for (int k = 1; k <= kz; ++k) for (int i = i1; i <= i2; ++i) for (int j = j1; j <= j2; ++j) {
      const int ib = (int)(*ibnd)(j, i);
      if (ib > 0) {
        const double xf = w.hefc[(size_t)((k - 1) * nsp + (ib - 1))];
        const double fext = (x0 * b0(j, i, k) + x1 * b1(j, i, k));
        f(j, i, k) = (1.0 - xf) * f(j, i, k) + xf * fext;
      }
    }
  });
}

// morelax_fraction (Main/mod_bdycod.F90:3962-3993); only called for w with frac = 0
void morelax_fraction(World& w, const ArrFn& getf, double frac) {
  const int kz = w.c.kz, nsp = w.c.nspgx;
  if (nsp <= 0) return;
  each(w, [&](Rank& r) {
    const Geom& g = r.g; Arr& f = getf(r);
    PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j) {
      const int ib = (int)r.ibnd_cr(j, i);
      if (ib > 0) {
        const double xf = w.hefc[(size_t)((k - 1) * nsp + (ib - 1))];
        f(j, i, k) = (1.0 - xf) * f(j, i, k) + xf * f(j, i, k) * frac;
      }
    }
  });
}

// morelax_chiten (Main/chemlib/mod_che_bdyco.F90:965-1026).  cba%ibnd and its
// south/north/west/east flags partition the same cells as ba_cr%ibnd > 0, so
// the four masked loops visit every sponge cell exactly once.
void morelax_chiten(World& w) {
  const double x1 = (w.xbctime + w.dtsec) / w.x.dtbdys, x0 = 1.0 - x1;
  const int kz = w.c.kz, nsp = w.c.nspgx;
  if (nsp <= 0) return;
  each(w, [&](Rank& r) {
    const Geom& g = r.g;
    for (int n = 0; n < w.ntr; ++n) {
      Arr& f = r.trac[n]; const Arr& b0 = r.chib0[n]; const Arr& b1 = r.chib1[n];
      PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j) {
        const int ib = (int)r.ibnd_cr(j, i);
        if (ib > 0) {
          const double xf = w.fcx[ib];
          const double fext = (x0 * b0(j, i, k) + x1 * b1(j, i, k));
          f(j, i, k) = (1.0 - xf) * f(j, i, k) + xf * fext;
        }
      }
    }
  });
}

// lowpass_init (Main/mod_bdycod.F90:3844-3896) on the global index space
void lowpass_init(World& w) {
  const oracle_config& c = w.c;
  const int jx = c.jx, iy = c.iy, kz = c.kz;
  const bool band = w.r[0].g.band, crm = w.r[0].g.crm;
  const int njcross = band ? jx : jx - 1, nicross = crm ? iy : iy - 1;
  const double ds = c.dx / 1000.0;
  w.km = std::max((int)std::lround((njcross * ds) / 1500.0), 1);
  w.lm = std::max((int)std::lround((nicross * ds) / 750.0), 1);
  const double dx = mathpi / (double)(njcross - 1), dy = mathpi / (double)(nicross - 1);
  const int km = w.km, lm = w.lm;
  std::vector<double> px(2 * km + 1), py(2 * lm + 1);
  for (int k = 1; k <= 2 * km; ++k) { const double q = (double)k / (double)km; px[k] = std::exp(-(q * q)); }
  for (int l = 1; l <= 2 * lm; ++l) { const double q = (double)l / (double)lm; py[l] = std::exp(-(q * q)); }
  w.bvx.assign((size_t)2 * km * jx, 0.0); w.bvy.assign((size_t)2 * lm * iy, 0.0);
  for (int k = 1; k <= 2 * km; ++k) for (int j = 1; j <= jx; ++j)
    w.bvx[(size_t)(k - 1) * jx + (j - 1)] = std::sqrt(2.0 / (double)(jx - 1) * px[k]) * std::sin((double)(k * (j - 2)) * dx);
  for (int l = 1; l <= 2 * lm; ++l) for (int i = 1; i <= iy; ++i)
    w.bvy[(size_t)(l - 1) * iy + (i - 1)] = std::sqrt(2.0 / (double)(iy - 1) * py[l]) * std::sin((double)(l * (i - 2)) * dy);
  const double cn0 = w.x.dtrad * w.rtb;
  w.cnudge.assign(kz + 1, 0.0);
  for (int k = 1; k <= kz; ++k) { const double q = w.gmeanz[k] / c.mo_h; w.cnudge[k] = cn0 * std::min(q * q, 1.0); }
}

// mospectral_nudge + lowpass_filter (Main/mod_bdycod.F90:3898-3960).
// row_reduce/column_reduce are MPI_Allreduce(SUM) over the ranks of a row /
// column (mod_mppparam.F90:20620-20664) with count = nk*(i2-i1+1) ELEMENTS OF
// THE CONTIGUOUS sx(ide1:ide2,1:2km) array: when i1:i2 is shorter than
// ide1:ide2 (the top row of ranks: ice2 = ide2-1) the tail of the array is not
// reduced and sxg keeps what an earlier call left there.  Restated as is.
struct SpecRange { int j1, j2, i1, i2, jj1, jj2, ii1, ii2; };
using RangeFn = std::function<SpecRange(const Geom&)>;
void mospectral_nudge(World& w, const RangeFn& range, const ArrFn& getf, const ArrFn& getb0, const ArrFn& getb1) {
  const double x1 = (w.xbctime + w.dtsec) * w.rtb, x0 = 1.0 - x1;
  const int kz = w.c.kz, jx = w.c.jx, iy = w.c.iy, km2 = 2 * w.km, lm2 = 2 * w.lm;
  const int nr = (int)w.r.size(), py = w.c.py;
  auto bvx = [&](int j, int k) { return w.bvx[(size_t)(k - 1) * jx + (j - 1)]; };
  auto bvy = [&](int i, int l) { return w.bvy[(size_t)(l - 1) * iy + (i - 1)]; };
  for (int k = 1; k <= kz; ++k) {
    for (auto& r : w.r) {
      const Geom& g = r.g; const SpecRange q = range(g);
      Arr& f = getf(r); const Arr& b0 = getb0(r); const Arr& b1 = getb1(r); Arr& zn1 = r.zn1;
      for (int i = q.i1; i <= q.i2; ++i) for (int j = q.j1; j <= q.j2; ++j)
        zn1(j, i) = (x0 * b0(j, i, k) + x1 * b1(j, i, k)) - f(j, i, k);
      const int ni = g.ide2 - g.ide1 + 1;
      for (int kk = 1; kk <= km2; ++kk) for (int i = q.i1; i <= q.i2; ++i) {
        double acc = 0.0;
        for (int j = q.jj1; j <= q.jj2; ++j) acc = acc + zn1(j, i) * bvx(j, kk);
        r.sx[(size_t)(kk - 1) * ni + (i - g.ide1)] = acc;
      }
    }
    for (int a = 0; a < nr; ++a) {   // row_reduce: ranks with the same loci, summed in rank order
      Rank& r = w.r[a]; const SpecRange q = range(r.g);
      const size_t count = (size_t)km2 * (size_t)(q.i2 - q.i1 + 1);
      for (size_t e = 0; e < count; ++e) {
        double acc = 0.0; bool first = true;
        for (int b = 0; b < nr; ++b) if (b % py == a % py) { acc = first ? w.r[b].sx[e] : acc + w.r[b].sx[e]; first = false; }
        r.sxg[e] = acc;
      }
    }
    for (auto& r : w.r) {
      const Geom& g = r.g; const SpecRange q = range(g); Arr& zn1 = r.zn1;
      const int ni = g.ide2 - g.ide1 + 1, nj = g.jde2 - g.jde1 + 1;
      for (auto& v : zn1.d) v = 0.0;
      for (int kk = 1; kk <= km2; ++kk) for (int i = q.i1; i <= q.i2; ++i) for (int j = q.j1; j <= q.j2; ++j)
        zn1(j, i) = zn1(j, i) + r.sxg[(size_t)(kk - 1) * ni + (i - g.ide1)] * bvx(j, kk);
      for (int l = 1; l <= lm2; ++l) for (int j = q.j1; j <= q.j2; ++j) {
        double acc = 0.0;
        for (int i = q.ii1; i <= q.ii2; ++i) acc = acc + zn1(j, i) * bvy(i, l);
        r.sy[(size_t)(l - 1) * nj + (j - g.jde1)] = acc;
      }
    }
    for (int a = 0; a < nr; ++a) {   // column_reduce: ranks with the same locj
      Rank& r = w.r[a]; const SpecRange q = range(r.g);
      const size_t count = (size_t)lm2 * (size_t)(q.j2 - q.j1 + 1);
      for (size_t e = 0; e < count; ++e) {
        double acc = 0.0; bool first = true;
        for (int b = 0; b < nr; ++b) if (b / py == a / py) { acc = first ? w.r[b].sy[e] : acc + w.r[b].sy[e]; first = false; }
        r.syg[e] = acc;
      }
    }
    for (auto& r : w.r) {
      const Geom& g = r.g; const SpecRange q = range(g); Arr& zn1 = r.zn1; Arr& f = getf(r);
      const int nj = g.jde2 - g.jde1 + 1;
      for (auto& v : zn1.d) v = 0.0;
      for (int l = 1; l <= lm2; ++l) for (int i = q.i1; i <= q.i2; ++i) for (int j = q.j1; j <= q.j2; ++j)
        zn1(j, i) = zn1(j, i) + r.syg[(size_t)(l - 1) * nj + (j - g.jde1)] * bvy(i, l);
      for (int i = q.ii1; i <= q.ii2; ++i) for (int j = q.jj1; j <= q.jj2; ++j)
        f(j, i, k) = f(j, i, k) + w.cnudge[k] * zn1(j, i);
    }
  }
}

// boundary [F90:448-529]
void uvstagtouvx(World& w);
void temp_to_tvirt(World& w);
void diag_snapshot(World& w);
void diag_difference(World& w, bool bdy);
void boundary(World& w) {
  const int kz = w.c.kz;
  diag_snapshot(w);
  bdyval(w);
  if (w.x.mo_top_nudge) motopnudge(w, 0.0, w.dtsec);
  morelax_external(w, S_U, [](Rank& r) -> Arr& { return r.u; }, [](Rank& r) -> Arr& { return r.dub0; },
                   [](Rank& r) -> Arr& { return r.dub1; });
  morelax_external(w, S_V, [](Rank& r) -> Arr& { return r.v; }, [](Rank& r) -> Arr& { return r.dvb0; },
                   [](Rank& r) -> Arr& { return r.dvb1; });
  morelax_external(w, S_CROSS, [](Rank& r) -> Arr& { return r.t; }, [](Rank& r) -> Arr& { return r.xtb0; },
                   [](Rank& r) -> Arr& { return r.xtb1; });
  morelax_external(w, S_CROSS, [](Rank& r) -> Arr& { return r.pai; }, [](Rank& r) -> Arr& { return r.xpaib0; },
                   [](Rank& r) -> Arr& { return r.xpaib1; });
  morelax_external(w, S_CROSS, [](Rank& r) -> Arr& { return r.qx[0]; }, [](Rank& r) -> Arr& { return r.xqb0; },
                   [](Rank& r) -> Arr& { return r.xqb1; });
  morelax_fraction(w, [](Rank& r) -> Arr& { return r.w; }, 0.0);
  if (w.ipptls > 0) {
    if (w.x.present_qc)
      morelax_external(w, S_CROSS, [](Rank& r) -> Arr& { return r.qx[1]; }, [](Rank& r) -> Arr& { return r.xlb0; },
                       [](Rank& r) -> Arr& { return r.xlb1; });
    if (w.ipptls > 1 && w.x.present_qi)
      morelax_external(w, S_CROSS, [](Rank& r) -> Arr& { return r.qx[2]; }, [](Rank& r) -> Arr& { return r.xib0; },
                       [](Rank& r) -> Arr& { return r.xib1; });
  }
  if (w.x.ichem == 1 && w.ntr > 0 && w.x.ichebdy != 0) morelax_chiten(w);
  if (w.x.mo_spectral_nudge) {
    w.tspectral = w.tspectral + w.dtsec;
    if ((int)std::fmod(w.tspectral, w.x.dtrad) == 0) {
      // NB the reference passes jci1,jci1 as the j update range of t [F90:502]
      mospectral_nudge(w, [](const Geom& g) { return SpecRange{g.jce1, g.jce2, g.ice1, g.ice2, g.jci1, g.jci1, g.ici1, g.ici2}; },
                       [](Rank& r) -> Arr& { return r.t; }, [](Rank& r) -> Arr& { return r.xtb0; },
                       [](Rank& r) -> Arr& { return r.xtb1; });
      mospectral_nudge(w, [](const Geom& g) { return SpecRange{g.jde1, g.jde2, g.ice1, g.ice2, g.jdi1, g.jdi2, g.ici1, g.ici2}; },
                       [](Rank& r) -> Arr& { return r.u; }, [](Rank& r) -> Arr& { return r.dub0; },
                       [](Rank& r) -> Arr& { return r.dub1; });
      mospectral_nudge(w, [](const Geom& g) { return SpecRange{g.jce1, g.jce2, g.ide1, g.ide2, g.jci1, g.jci2, g.idi1, g.idi2}; },
                       [](Rank& r) -> Arr& { return r.v; }, [](Rank& r) -> Arr& { return r.dvb0; },
                       [](Rank& r) -> Arr& { return r.dvb1; });
    }
  }
  diag_difference(w, true);
  uvstagtouvx(w);
  temp_to_tvirt(w);
  each(w, [&](Rank& r) {
    const Geom& g = r.g;
    PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jce1; j <= g.jce2; ++j)
      r.tetav(j, i, k) = r.tvirt(j, i, k) / r.pai(j, i, k);
  });
}

// mkslice, idynamic == 3 branch (Main/mod_slice.F90:115-173).  atms%ps2d is
// sfs%psb; in a MOLOCH run psa and psb are the same surface pressure the
// dycore extrapolates [F90:354], so `ps` is used.
void mkslice(World& w) {
  const int kz = w.c.kz, kzp1 = kz + 1;
  const double rhmin = w.x.rhmin, rhmax = w.x.rhmax;
  each(w, [&](Rank& r) {
    const Geom& g = r.g;
    for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jce1; j <= g.jce2; ++j) r.pf3d(j, i, kzp1) = r.ps(j, i);
    PAR2 for (int k = 2; k <= kz; ++k) for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jce1; j <= g.jce2; ++j)
      r.pf3d(j, i, k) = p00 * std::pow(0.5 * (r.pai(j, i, k) + r.pai(j, i, k - 1)), cpovr);
    for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jce1; j <= g.jce2; ++j)
      r.pf3d(j, i, 1) = r.pf3d(j, i, 2) - egrav * r.rho(j, i, 1) * (r.zetaf(j, i, 1) - r.zetaf(j, i, 2));
    for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j) {
      r.rhox2d(j, i) = r.ps(j, i) / (rgas * r.t(j, i, kz));
      r.tp2d(j, i) = r.t(j, i, kz) * std::pow(r.ps(j, i) / r.p(j, i, kz), rovcp);
    }
    PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ice1; i <= g.ice2; ++i) for (int j = g.jce1; j <= g.jce2; ++j)
      r.th3d(j, i, k) = r.t(j, i, k) * std::pow(p00 / r.p(j, i, k), rovcp);
    for (int n = 0; n < w.nqx; ++n) {
      Arr& q = r.qx[n];
      PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j)
        if (q(j, i, k) < qxcheckval[n]) q(j, i, k) = qxzeroval[n];
    }
    if (w.x.ichem == 1) for (int n = 0; n < w.ntr; ++n) {
      Arr& q = r.trac[n];
      PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j)
        if (q(j, i, k) < 1.0e-50) q(j, i, k) = 0.0;
    }
    PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j)
      r.rhb3d(j, i, k) = std::min(std::max(r.qx[0](j, i, k) / r.qsat(j, i, k), rhmin), rhmax);
    PAR2 for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j)
      r.wpx3d(j, i, k) = -egrav * r.rho(j, i, k) * 0.5 * (r.w(j, i, k + 1) + r.w(j, i, k));
    if (w.x.icldmstrat == 1) {
      for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j) {
        r.th700(j, i) = r.th3d(j, i, kz);
        for (int k = 2; k <= kz - 1; ++k) {
          if (r.p(j, i, k) > 70000.0) {
            const double w1 = (r.p(j, i, k) - 70000.0) / (r.p(j, i, k) - r.p(j, i, k - 1));
            const double w2 = 1.0 - w1;
            r.th700(j, i) = r.th3d(j, i, k - 1) * w1 + r.th3d(j, i, k) * w2;
            break;
          }
        }
      }
    }
    // common tail (Main/mod_slice.F90:342-384): tropopause pressure, its level, highest PBL level
    const double twopi = mathpi * 2.0;
    const double anorth[6] = {7.9925, 8.3329, 24.1731, -1.8069, 0.1082, -0.1493};
    const double asouth[6] = {8.1797, 8.1455, -23.4839, 1.1464, 0.0798, -0.1491};
    if (w.x.irceideal != 1) {
      for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j) {
        double ztrop;
        if (r.xlat(j, i) > 0.0)
          ztrop = anorth[0] + anorth[1] / std::pow(1.0 + std::exp(-(r.xlat(j, i) - anorth[2]) / anorth[3]), anorth[4]) +
                  anorth[5] * std::cos((twopi * (w.x.calday - 28.0)) / w.x.dayspy);
        else
          ztrop = asouth[0] + asouth[1] / std::pow(1.0 + std::exp(-(r.xlat(j, i) - asouth[2]) / asouth[3]), asouth[4]) +
                  asouth[5] * std::cos((twopi * (w.x.calday - 28.0)) / w.x.dayspy);
        r.ptrop(j, i) = p00 * std::exp(-ztrop / 8.4);
      }
    }
    for (auto& v : r.ktrop.d) v = kz;
    for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j)
      for (int k = kz - 1; k >= 2; --k) {
        r.ktrop(j, i) = k;
        if (r.p(j, i, k) < r.ptrop(j, i)) break;
      }
    if (w.x.ibltyp == 1) {
      for (auto& v : r.kmxpbl.d) v = kz;
      for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j)
        for (int k = kz - 1; k >= 2; --k) {
          if (r.zeta(j, i, k) > 5000.0) break;
          r.kmxpbl(j, i) = k;
        }
    }
  });
}

// massck, idynamic == 3 branch (Main/mod_massck.F90:77-185): the sums over the
// atmosphere (the surface terms crrate, ncrrate, qfx belong to the physics) in
// the reference's single running sum per rank, then sumall over ranks.
// dz = zetaf(k) - zetaf(k+1) (Main/mod_params.F90:3389).
void massck(World& w, double out[4]) {
  const int kz = w.c.kz;
  const double dxsq = w.dx * w.dx, dt = w.dtsec, dx = w.dx;
  double drymass = 0.0, dryadv = 0.0, qmass = 0.0, qadv = 0.0;
  for (auto& r : w.r) {
    const Geom& g = r.g;
    auto dz = [&](int j, int i, int k) { return r.zetaf(j, i, k) - r.zetaf(j, i, k + 1); };
    double tdrym = 0.0, tdadv = 0.0, tqmass = 0.0, tqadv = 0.0;
    for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j)
      tdrym = tdrym + dxsq * dz(j, i, k) * r.rho(j, i, k);
    if (g.bl) for (int k = 1; k <= kz; ++k) for (int i = g.ice1; i <= g.ice2; ++i)
      tdadv = tdadv + r.u(g.jde1, i, k) * dt * dx * dz(g.jce1, i, k) * r.rho(g.jce1, i, k);
    if (g.br) for (int k = 1; k <= kz; ++k) for (int i = g.ice1; i <= g.ice2; ++i)
      tdadv = tdadv - r.u(g.jde2, i, k) * dt * dx * dz(g.jce2, i, k) * r.rho(g.jce2, i, k);
    if (g.bb) for (int k = 1; k <= kz; ++k) for (int j = g.jci1; j <= g.jci2; ++j)
      tdadv = tdadv + r.v(j, g.ide1, k) * dt * dx * dz(j, g.ice1, k) * r.rho(j, g.ice1, k);
    if (g.bt) for (int k = 1; k <= kz; ++k) for (int j = g.jci1; j <= g.jci2; ++j)
      tdadv = tdadv - r.v(j, g.ide2, k) * dt * dx * dz(j, g.ice2, k) * r.rho(j, g.ice2, k);
    for (int n = 0; n < w.nqx; ++n) {
      Arr& q = r.qx[n];
      for (int k = 1; k <= kz; ++k) for (int i = g.ici1; i <= g.ici2; ++i) for (int j = g.jci1; j <= g.jci2; ++j)
        tqmass = tqmass + q(j, i, k) * dxsq * dz(j, i, k) * r.rho(j, i, k);
      if (g.bl) for (int k = 1; k <= kz; ++k) for (int i = g.ice1; i <= g.ice2; ++i)
        tqadv = tqadv + q(g.jce1, i, k) * r.u(g.jde1, i, k) * dt * dx * dz(g.jce1, i, k) * r.rho(g.jce1, i, k);
      if (g.br) for (int k = 1; k <= kz; ++k) for (int i = g.ice1; i <= g.ice2; ++i)
        tqadv = tqadv - q(g.jce2, i, k) * r.u(g.jde2, i, k) * dt * dx * dz(g.jce2, i, k) * r.rho(g.jce2, i, k);
      if (g.bb) for (int k = 1; k <= kz; ++k) for (int j = g.jci1; j <= g.jci2; ++j)
        tqadv = tqadv + q(j, g.ice1, k) * r.v(j, g.ide1, k) * dt * dx * dz(j, g.ice1, k) * r.rho(j, g.ice1, k);
      if (g.bt) for (int k = 1; k <= kz; ++k) for (int j = g.jci1; j <= g.jci2; ++j)
        tqadv = tqadv - q(j, g.ice2, k) * r.v(j, g.ide2, k) * dt * dx * dz(j, g.ice2, k) * r.rho(j, g.ice2, k);
    }
    drymass += tdrym; dryadv += tdadv; qmass += tqmass; qadv += tqadv;   // sumall
  }
  out[0] = drymass; out[1] = dryadv; out[2] = qmass; out[3] = qadv;
}

// moloch [F90:312-446]: physics disabled ([F90:362]); massck/report skipped.
// Without oracle_set_ext: do_apply_bdy = .false. (irceideal / test modes,
// [F90:305]) and no mkslice.
void moloch_step(World& w) {
  reset_tendencies(w);
  dynamical_core(w);
  if (w.ext_set && w.x.do_bdy) boundary(w);
  diagnostics(w);
  if (w.ext_set && w.x.do_slice) mkslice(w);
  status_update(w, w.dtsec);
}

ArrFn field_getter(World& w, const std::string& name, int n, bool& ok) {
  ok = true;
  if (name == "qx") { if (n < 1 || n > w.nqx) ok = false; return [n](Rank& r) -> Arr& { return r.qx[n - 1]; }; }
  if (name == "trac") { if (n < 1 || n > w.ntr) ok = false; return [n](Rank& r) -> Arr& { return r.trac[n - 1]; }; }
  if (w.r[0].reg.find(name) == w.r[0].reg.end()) { ok = false; }
  return [name](Rank& r) -> Arr& { return *r.reg[name].a; };
}

}  // namespace

// ---------------------------------------------------------------------------
// C API
// ---------------------------------------------------------------------------
extern "C" {

const char* oracle_last_error(void) { return g_err.c_str(); }

void* oracle_create(const oracle_config* cfg) {
  if (!cfg || cfg->jx < 4 || cfg->iy < 4 || cfg->kz < 4 || cfg->px < 1 || cfg->py < 1 || cfg->nqx < 1 ||
      cfg->nqx > 10) { g_err = "bad config"; return nullptr; }
  if (cfg->ipptls > 1 && cfg->nqx < 5) { g_err = "ipptls=2 needs nqx>=5"; return nullptr; }
  if (cfg->ipptls == 1 && cfg->nqx < 2) { g_err = "ipptls=1 needs nqx>=2"; return nullptr; }
  World* w = new World();
  w->c = *cfg;
  w->nqx = cfg->nqx; w->ntr = cfg->ntr; w->iqfrst = 2; w->ipptls = cfg->ipptls;
  w->lrotllr = cfg->lrotllr != 0; w->do_divdamp = cfg->mo_divdamp != 0; w->do_divfilter = cfg->mo_divfilter != 0;
  w->dx = cfg->dx; w->rdx = 1.0 / cfg->dx; w->dtsec = cfg->dtsec;
  const int n = cfg->px * cfg->py;
  w->r.resize(n);
  for (int i = 0; i < n; ++i) {
    if (!make_geom(*cfg, i, w->r[i].g)) { delete w; return nullptr; }
    alloc_rank(*w, w->r[i]);
  }
  w->rlat.assign(cfg->iy + 3, 0.0);
  if (cfg->nspgx > 0) w->hefc.assign((size_t)cfg->nspgx * cfg->kz, 0.0);
  return w;
}

void oracle_destroy(void* h) { delete (World*)h; }

int oracle_set_ext(void* h, const oracle_ext_config* x) {
  World& w = *(World*)h;
  if (!x) { g_err = "null ext config"; return 1; }
  if (w.ext_set) { g_err = "oracle_set_ext called twice"; return 1; }
  if (x->do_bdy && !(x->dtbdys > 0.0)) { g_err = "dtbdys must be > 0"; return 1; }
  if (x->do_bdy && x->mo_spectral_nudge && !(x->dtrad > 0.0)) { g_err = "dtrad must be > 0"; return 1; }
  w.x = *x; w.ext_set = true;
  for (auto& r : w.r) alloc_ext(w, r);
  w.fcx.assign(std::max(w.c.nspgx, 1) + 1, 0.0);
  return 0;
}

long oracle_global_size(void* h, const char* name) {
  World& w = *(World*)h; std::string s(name);
  const long plane = (long)w.c.jx * w.c.iy;
  if (s == "rlat") return w.c.iy + 1;
  if (s == "fcx") return w.c.nspgx;
  if (s == "bvx") return (long)w.bvx.size();
  if (s == "bvy") return (long)w.bvy.size();
  if (s == "tnudge" || s == "cnudge" || s == "gmeanz") return w.c.kz;
  if (s == "hefc") return (long)w.c.nspgx * w.c.kz;
  if (s == "ffilt" || s == "xkdamp" || s == "xknu" || s == "gzitakh" || s == "zitah") return w.c.kz;
  if (s == "gzitak" || s == "zita") return w.c.kz + 1;
  std::vector<FieldInfo> f;
  if (!lookup(w, w.r[0], s, f)) return 0;
  long tot = 0; for (auto& x : f) tot += plane * x.nk;
  return tot;
}

int oracle_set_global(void* h, const char* name, const double* src) {
  World& w = *(World*)h; std::string s(name);
  const int jx = w.c.jx, iy = w.c.iy;
  if (s == "rlat") { for (int i = 1; i <= iy + 1; ++i) w.rlat[i] = src[i - 1]; return 0; }
  if (s == "hefc") { std::copy(src, src + w.hefc.size(), w.hefc.begin()); return 0; }
  if (s == "tnudge" || s == "cnudge") {   // override the sumall-order dependent profiles (decomposition tests)
    std::vector<double>& v = (s == "tnudge") ? w.tnudge : w.cnudge;
    v.assign(w.c.kz + 1, 0.0); for (int k = 1; k <= w.c.kz; ++k) v[k] = src[k - 1]; return 0; }
  if (s == "fcx") { w.fcx.assign(w.c.nspgx + 1, 0.0); for (int n = 1; n <= w.c.nspgx; ++n) w.fcx[n] = src[n - 1]; return 0; }
  if (s == "ffilt") { w.ffilt.assign(w.c.kz + 1, 0.0); for (int k = 1; k <= w.c.kz; ++k) w.ffilt[k] = src[k - 1]; return 0; }
  // init_moloch's 1-D tables (Main/mod_moloch.F90:273-274, 294-299): settable so that a run can start from the
  // host model's own tables (bench.py's parity fixture starts the oracle from exactly the arrays the library gets)
  if (s == "xkdamp" || s == "xknu" || s == "gzitakh" || s == "gzitak") {
    std::vector<double>& v = (s == "xkdamp") ? w.xkdamp : (s == "xknu") ? w.xknu : (s == "gzitakh") ? w.gzitakh : w.gzitak;
    const int n = (s == "gzitak") ? w.c.kz + 1 : w.c.kz;
    if ((int)v.size() < n + 1) v.assign(n + 1, 0.0);
    for (int k = 1; k <= n; ++k) v[k] = src[k - 1];
    return 0;
  }
  const bool perj = w.r[0].g.band, peri = w.r[0].g.crm;
  for (auto& r : w.r) {
    std::vector<FieldInfo> f;
    if (!lookup(w, r, s, f)) { g_err = "unknown field " + s; return 1; }
    long off = 0;
    for (auto& fi : f) {
      Arr& a = *fi.a;
      for (int kk = 0; kk < fi.nk; ++kk) for (int i = a.ilo; i <= a.ihi; ++i) for (int j = a.jlo; j <= a.jhi; ++j) {
        int gj = j, gi = i;
        if (perj) gj = ((j - 1) % jx + jx) % jx + 1;
        if (peri) gi = ((i - 1) % iy + iy) % iy + 1;
        if (gj < 1 || gj > jx || gi < 1 || gi > iy) continue;
        a(j, i, a.klo + kk) = src[off + ((long)kk * iy + (gi - 1)) * jx + (gj - 1)];
      }
      off += (long)fi.nk * iy * jx;
    }
  }
  return 0;
}

int oracle_get_global(void* h, const char* name, double* dst) {
  World& w = *(World*)h; std::string s(name);
  const int jx = w.c.jx, iy = w.c.iy, kz = w.c.kz;
  auto cp = [&](const std::vector<double>& v, int n) { for (int k = 1; k <= n; ++k) dst[k - 1] = v[k]; return 0; };
  if (s == "ffilt") return cp(w.ffilt, kz);
  if (s == "xkdamp") return cp(w.xkdamp, kz);
  if (s == "xknu") return cp(w.xknu, kz);
  if (s == "gzitak") return cp(w.gzitak, kz + 1);
  if (s == "gzitakh") return cp(w.gzitakh, kz);
  if (s == "zita") return cp(w.zita, kz + 1);
  if (s == "zitah") return cp(w.zitah, kz);
  if (s == "tnudge" && !w.tnudge.empty()) return cp(w.tnudge, kz);
  if (s == "cnudge" && !w.cnudge.empty()) return cp(w.cnudge, kz);
  if (s == "gmeanz" && !w.gmeanz.empty()) return cp(w.gmeanz, kz);
  if (s == "hefc") { std::copy(w.hefc.begin(), w.hefc.end(), dst); return 0; }
  if (s == "bvx") { std::copy(w.bvx.begin(), w.bvx.end(), dst); return 0; }   // [2km][jx]
  if (s == "bvy") { std::copy(w.bvy.begin(), w.bvy.end(), dst); return 0; }   // [2lm][iy]
  for (auto& r : w.r) {
    std::vector<FieldInfo> f;
    if (!lookup(w, r, s, f)) { g_err = "unknown field " + s; return 1; }
    long off = 0;
    for (auto& fi : f) {
      Arr& a = *fi.a;
      int j1, j2, i1, i2; owned_box(r.g, fi.stag, j1, j2, i1, i2);
      j1 = std::max(j1, a.jlo); j2 = std::min(j2, a.jhi); i1 = std::max(i1, a.ilo); i2 = std::min(i2, a.ihi);
      for (int kk = 0; kk < fi.nk; ++kk) for (int i = i1; i <= i2; ++i) for (int j = j1; j <= j2; ++j)
        dst[off + ((long)kk * iy + (i - 1)) * jx + (j - 1)] = a(j, i, a.klo + kk);
      off += (long)fi.nk * iy * jx;
    }
  }
  return 0;
}

int oracle_setup_static(void* h) { return setup_static(*(World*)h); }
int oracle_init_state(void* h) { return init_state(*(World*)h); }
int oracle_step(void* h, int nsteps) { World& w = *(World*)h; for (int n = 0; n < nsteps; ++n) moloch_step(w); return 0; }
int oracle_reset_tendencies(void* h) { reset_tendencies(*(World*)h); return 0; }
int oracle_sound(void* h) { World& w = *(World*)h; sound(w, w.dtsound); return 0; }
int oracle_advection(void* h) { World& w = *(World*)h; advection(w, w.dtstepa); return 0; }
int oracle_wafone(void* h, const char* field, int n) {
  World& w = *(World*)h; bool ok;
  ArrFn f = field_getter(w, field, n, ok);
  if (!ok) { g_err = std::string("unknown field ") + field; return 1; }
  wafone(w, f, w.dtstepa); return 0;
}
int oracle_dynamical_core(void* h) { dynamical_core(*(World*)h); return 0; }
int oracle_diagnostics(void* h) { diagnostics(*(World*)h); return 0; }
int oracle_status_update(void* h) { World& w = *(World*)h; status_update(w, w.dtsec); return 0; }
static int need_bdy(World& w) {
  if (!(w.ext_set && w.x.do_bdy)) { g_err = "boundary not configured (oracle_set_ext with do_bdy)"; return 1; }
  if (w.rtb == 0.0) { g_err = "oracle_setup_static has not run after oracle_set_ext"; return 1; }
  return 0;
}
int oracle_boundary(void* h) { World& w = *(World*)h; if (need_bdy(w)) return 1; boundary(w); return 0; }
int oracle_bdyval(void* h) { World& w = *(World*)h; if (need_bdy(w)) return 1; bdyval(w); return 0; }
int oracle_mkslice(void* h) {
  World& w = *(World*)h;
  if (!w.ext_set) { g_err = "oracle_set_ext has not been called"; return 1; }
  mkslice(w); return 0;
}
int oracle_massck(void* h, double* out4) { massck(*(World*)h, out4); return 0; }
// maxval/minval of ps over the interior + maxall/minall [F90:408-411]; non-finite values are counted
int oracle_ps_check(void* h, double* maxmin, int* nonfinite) {
  World& w = *(World*)h;
  double mx = -1.0e300, mn = 1.0e300; int bad = 0;
  for (auto& r : w.r) for (int i = r.g.ici1; i <= r.g.ici2; ++i) for (int j = r.g.jci1; j <= r.g.jci2; ++j) {
    const double p = r.ps(j, i);
    if (!std::isfinite(p)) { ++bad; continue; }
    mx = std::max(mx, p); mn = std::min(mn, p);
  }
  maxmin[0] = mx; maxmin[1] = mn; *nonfinite = bad;
  return 0;
}
int oracle_set_xbctime(void* h, double t) { ((World*)h)->xbctime = t; return 0; }
double oracle_get_xbctime(void* h) { return ((World*)h)->xbctime; }
int oracle_get_int(void* h, const char* name) {
  World& w = *(World*)h; std::string s(name);
  if (s == "nztop") return w.nztop;
  if (s == "km") return w.km;
  if (s == "lm") return w.lm;
  return -1;
}

void oracle_set_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int oracle_get_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
