// geo.h -- geometry of one rank, index helpers and physical constants.
//
// Host/device-neutral (no CUDA runtime calls): included by the CUDA library
// through common.cuh and by the host-compiled instantiation of the boundary
// cell functions that the CPU tests use (tests/emu).
#pragma once
#include <stdint.h>
#include "../../include/moloch_b200.h"

#if defined(__CUDACC__)
#define MB_HD __host__ __device__ __forceinline__
#else
#define MB_HD inline
#endif

namespace mb {

constexpr int HJ = 4;
constexpr int HI = 3;

// physical constants: Share/mod_constants.F90:105-233 (non-RCEMIP branch)
constexpr double egrav = 9.80665;
constexpr double boltzk = 1.3806490e-23;
constexpr double navgdr = 6.02214076e23;
constexpr double amd = 28.96454;
constexpr double amw = 18.01528;
constexpr double rgasmol = navgdr * boltzk;
constexpr double rgas = (rgasmol / amd) * 1000.0;
constexpr double cpd = 3.5 * rgas;
constexpr double cvd = 2.5 * rgas;
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
constexpr double rearthrad = 1.0 / 6.371229e6;

// Geometry of one rank, passed by value to every kernel.  Index ranges follow
// setup_model_indexes (Main/mod_atm_interface.F90:182-382).
struct Geo {
  int NJ, NI, j0, i0;
  long long plane;  // NJ*NI
  int kz;
  int jde1, jde2, ide1, ide2, jdi1, jdi2, idi1, idi2, jdii1, jdii2, idii1, idii2;
  int jce1, jce2, ice1, ice2, jci1, jci2, ici1, ici2;
  int gl, gr, gb, gt;  // 1 where a neighbour exists (ma%jbl1 ...)
  int bl, br, bb, bt;  // ma%has_bdy*
  int jmin, jmax, imin, imax;  // Main/mod_moloch.F90:280-293
  int lrotllr, ipptls, nqx, ntr;
};

MB_HD long long gidx(const Geo& g, int j, int i, int k) {
  return (long long)(k - 1) * g.plane + (long long)(i - g.i0) * g.NJ + (j - g.j0);
}
MB_HD long long gidx2(const Geo& g, int j, int i) {
  return (long long)(i - g.i0) * g.NJ + (j - g.j0);
}


// setup_model_indexes (Main/mod_atm_interface.F90:182-382) + init_moloch's
// clamp limits (Main/mod_moloch.F90:280-293) + the padded device box
inline Geo geo_from_cfg(const moloch_b200_config& f) {
  Geo g;
  g.kz = f.kz;
  g.jde1 = f.jde1; g.jde2 = f.jde2; g.ide1 = f.ide1; g.ide2 = f.ide2;
  g.jce1 = f.jce1; g.jce2 = f.jce2; g.ice1 = f.ice1; g.ice2 = f.ice2;
  g.bl = f.has_bdy_left != 0; g.br = f.has_bdy_right != 0; g.bb = f.has_bdy_bottom != 0; g.bt = f.has_bdy_top != 0;
  g.gl = g.bl ? 0 : 1; g.gr = g.br ? 0 : 1; g.gb = g.bb ? 0 : 1; g.gt = g.bt ? 0 : 1;
  // setup_model_indexes, Main/mod_atm_interface.F90:182-382
  g.jdi1 = g.jde1 + (g.bl ? 1 : 0); g.jdii1 = g.jde1 + (g.bl ? 2 : 0);
  g.jdi2 = g.jde2 - (g.br ? 1 : 0); g.jdii2 = g.jde2 - (g.br ? 2 : 0);
  g.idi1 = g.ide1 + (g.bb ? 1 : 0); g.idii1 = g.ide1 + (g.bb ? 2 : 0);
  g.idi2 = g.ide2 - (g.bt ? 1 : 0); g.idii2 = g.ide2 - (g.bt ? 2 : 0);
  g.jci1 = g.jce1 + (g.bl ? 1 : 0); g.jci2 = g.jce2 - (g.br ? 1 : 0);
  g.ici1 = g.ice1 + (g.bb ? 1 : 0); g.ici2 = g.ice2 - (g.bt ? 1 : 0);
  // init_moloch, Main/mod_moloch.F90:280-293
  const int jcross2 = f.bandflag ? f.jx : f.jx - 1, icross2 = f.crmflag ? f.iy : f.iy - 1;
  g.jmin = 1; g.jmax = jcross2; g.imin = 1; g.imax = icross2;
  if (f.bandflag) { g.jmin = 1 - 2; g.jmax = jcross2 + 2; }
  if (f.crmflag) { g.jmin = 1 - 2; g.jmax = jcross2 + 2; g.imin = 1 - 2; g.imax = icross2 + 2; }
  g.lrotllr = f.lrotllr; g.ipptls = f.ipptls; g.nqx = f.nqx; g.ntr = f.ntr;
  g.j0 = g.jde1 - HJ; g.i0 = g.ide1 - HI;
  int nj = (g.jde2 - g.jde1 + 1) + 2 * HJ;
  nj = (nj + 3) / 4 * 4;
  g.NJ = nj; g.NI = (g.ide2 - g.ide1 + 1) + 2 * HI;
  g.plane = (long long)g.NJ * g.NI;
  return g;
}

// levels and species count of an ABI field (0 levels: not allocated in this configuration)
inline void field_shape(const moloch_b200_config& f, int id, int& nk, int& nspec) {
  const int kz = f.kz;
  nspec = 1;
  switch (id) {
    case MB_W: case MB_FMZF: case MB_S: nk = kz + 1; break;
    case MB_QX: case MB_QXTEN: nk = kz; nspec = f.nqx; break;
    case MB_TRAC: case MB_CHITEN: nk = kz; nspec = f.ntr; break;
    case MB_PS: case MB_HX: case MB_HY: case MB_MSFX: case MB_MSFU: case MB_MSFV: case MB_CORU: case MB_CORV:
      nk = 1; break;
    // ---- ABI v2: allocated only when the feature is configured -------------------
    case MB_TKE: case MB_TKETEN: nk = (f.ibltyp == 2) ? kz + 1 : 0; break;
    case MB_TKEX: nk = (f.ibltyp == 2) ? kz : 0; break;
    case MB_DUB0: case MB_DUB1: case MB_DVB0: case MB_DVB1: case MB_XTB0: case MB_XTB1: case MB_XPAIB0:
    case MB_XPAIB1: case MB_XQB0: case MB_XQB1:
      nk = f.do_bdy ? kz : 0; break;
    case MB_XLB0: case MB_XLB1: nk = (f.do_bdy && f.present_qc) ? kz : 0; break;
    case MB_XIB0: case MB_XIB1: nk = (f.do_bdy && f.present_qi) ? kz : 0; break;
    case MB_XPSB0: case MB_XPSB1: nk = f.do_bdy ? 1 : 0; break;
    case MB_CHIB0: case MB_CHIB1:
      nk = (f.do_bdy && f.ichem && f.ichebdy != 0 && f.ntr > 0) ? kz : 0; nspec = f.ntr > 0 ? f.ntr : 1; break;
    case MB_PF3D: nk = f.do_slice ? kz + 1 : 0; break;
    case MB_ZETAF: nk = (f.do_slice || f.do_massck) ? kz + 1 : 0; break;
    case MB_TH3D: case MB_RHB3D: case MB_WPX3D: nk = f.do_slice ? kz : 0; break;
    case MB_RHOX2D: case MB_TP2D: case MB_TH700: case MB_XLAT: case MB_PTROP: case MB_KTROP: case MB_KMXPBL:
      nk = f.do_slice ? 1 : 0; break;
    case MB_TEN0: case MB_QEN0: case MB_TDIAG_ADH: case MB_QDIAG_ADH: case MB_TDIAG_BDY: case MB_QDIAG_BDY:
      nk = f.idiag > 0 ? kz : 0; break;
    case MB_CHITEN0: case MB_CADVHDIAG: case MB_CBDYDIAG:
      nk = (f.ichdiag > 0 && f.ichem && f.ntr > 0) ? kz : 0; nspec = f.ntr > 0 ? f.ntr : 1; break;
    default: nk = kz; break;
  }
}

}  // namespace mb
