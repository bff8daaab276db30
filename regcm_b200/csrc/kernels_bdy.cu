// kernels_bdy.cu -- lateral boundary (`boundary`, Main/mod_moloch.F90:448-529),
// mkslice (Main/mod_slice.F90:115-173) and the TKE helpers of the UW-PBL path.
//
// All of this is HBM-bound pointwise work on boundary strips, the sponge ring
// or whole fields; the per-cell arithmetic lives in bdy_cells.h (shared with
// the host-compiled instantiation the CPU tests use), this file maps threads
// onto cells: 32 consecutive j per warp row (coalesced 256-B rows), one level
// per blockIdx.z.  Compiled -fmad=false like the rest of the library.
#include <cstring>
#include <vector>
#include "bdy_cells.h"
#include "common.cuh"

namespace mb {

constexpr int BX = 32, BY = 8;
static inline dim3 grid3(int nj, int ni, int nk) {
  return dim3((unsigned)((nj + BX - 1) / BX), (unsigned)((ni + BY - 1) / BY), (unsigned)nk);
}

// ---- bdyval ----------------------------------------------------------------------
// boundary lines: x = position along the line, y = level, z = side
__global__ void moloch_bdyval_we(BdyArgs a) {
  const int i = a.g.ide1 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i > a.g.ide2) return;
  bdyval_we_cell(a, (int)blockIdx.z, i, 1 + (int)blockIdx.y);
}
__global__ void moloch_bdyval_sn(BdyArgs a) {
  const int j = a.g.jde1 + blockIdx.x * blockDim.x + threadIdx.x;
  if (j > a.g.jde2) return;
  bdyval_sn_cell(a, (int)blockIdx.z, j, 1 + (int)blockIdx.y);
}
__global__ void moloch_chem_bdyval_we(BdyArgs a) {
  const int i = a.g.ice1 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i > a.g.ice2) return;
  const int k = 1 + (int)(blockIdx.y % a.g.kz), n = (int)(blockIdx.y / a.g.kz);
  chem_bdyval_we_cell(a, (int)blockIdx.z, i, k, n);
}
__global__ void moloch_chem_bdyval_sn(BdyArgs a) {
  const int j = a.g.jce1 + blockIdx.x * blockDim.x + threadIdx.x;
  if (j > a.g.jce2) return;
  const int k = 1 + (int)(blockIdx.y % a.g.kz), n = (int)(blockIdx.y / a.g.kz);
  chem_bdyval_sn_cell(a, (int)blockIdx.z, j, k, n);
}

static BdyArgs bdy_args(Ctx& c, double xbctime) {
  BdyArgs a;
  memset(&a, 0, sizeof(a));
  a.g = c.g;
  a.u = c.f[MB_U].p; a.v = c.f[MB_V].p; a.w = c.f[MB_W].p; a.t = c.f[MB_T].p; a.pai = c.f[MB_PAI].p;
  a.qx = c.f[MB_QX].p; a.trac = c.f[MB_TRAC].p; a.ps = c.f[MB_PS].p; a.tke = c.f[MB_TKE].p;
  a.ux = c.f[MB_UX].p; a.vx = c.f[MB_VX].p; a.tvirt = c.f[MB_TVIRT].p; a.tetav = c.f[MB_TETAV].p;
  a.dub0 = c.f[MB_DUB0].p; a.dub1 = c.f[MB_DUB1].p; a.dvb0 = c.f[MB_DVB0].p; a.dvb1 = c.f[MB_DVB1].p;
  a.xtb0 = c.f[MB_XTB0].p; a.xtb1 = c.f[MB_XTB1].p; a.xpaib0 = c.f[MB_XPAIB0].p; a.xpaib1 = c.f[MB_XPAIB1].p;
  a.xqb0 = c.f[MB_XQB0].p; a.xqb1 = c.f[MB_XQB1].p; a.xlb0 = c.f[MB_XLB0].p; a.xlb1 = c.f[MB_XLB1].p;
  a.xib0 = c.f[MB_XIB0].p; a.xib1 = c.f[MB_XIB1].p; a.xpsb0 = c.f[MB_XPSB0].p; a.xpsb1 = c.f[MB_XPSB1].p;
  a.chib0 = c.f[MB_CHIB0].p; a.chib1 = c.f[MB_CHIB1].p;
  a.ib_cr = c.ibnd[MB_IBND_CR]; a.ib_ud = c.ibnd[MB_IBND_UD]; a.ib_vd = c.ibnd[MB_IBND_VD];
  a.hefc = c.tab[MB_TAB_HEFC]; a.tnudge = c.tab[MB_TAB_TNUDGE]; a.fcx = c.tab[MB_TAB_FCX];
  // x1 = (xbctime + dt)*rtb, rtb = d_one/dtbdys (Main/mod_bdycod.F90:504, :1631)
  const double rtb = 1.0 / c.cfg.dtbdys;
  a.x1 = (xbctime + c.cfg.dtsec) * rtb; a.x0 = 1.0 - a.x1;
  a.xc1 = (xbctime + c.cfg.dtsec) / c.cfg.dtbdys; a.xc0 = 1.0 - a.xc1;   // mod_che_bdyco.F90:507
  a.dtsec = c.cfg.dtsec; a.tkemin = c.cfg.tkemin;
  a.nspgx = c.cfg.nspgx; a.iqfrst = c.cfg.iqfrst; a.present_qc = c.cfg.present_qc; a.present_qi = c.cfg.present_qi;
  a.tke_on = c.cfg.ibltyp == 2; a.nztop = c.cfg.nztop; a.top_nudge = c.cfg.mo_top_nudge;
  a.ichem = c.cfg.ichem && c.cfg.ntr > 0; a.ichebdy = c.cfg.ichebdy;
  return a;
}

// bdyval (MOLOCH branch) with the time weights of `xbctime`; the caller
// advances xbctime afterwards (:2653)
int k_bdyval(Ctx& c, double xbctime) {
  const Geo& g = c.g;
  const BdyArgs a = bdy_args(c, xbctime);
  const int kz = g.kz;
  if (g.bl || g.br) {
    LaunchScope ls(c, KID_BDYVAL);
    const int n = g.ide2 - g.ide1 + 1;
    moloch_bdyval_we<<<dim3((unsigned)((n + 127) / 128), (unsigned)kz, 2), 128, 0, c.stream>>>(a);
    MB_CUDA(cudaGetLastError());
  }
  if (g.bb || g.bt) {
    LaunchScope ls(c, KID_BDYVAL);
    const int n = g.jde2 - g.jde1 + 1;
    moloch_bdyval_sn<<<dim3((unsigned)((n + 127) / 128), (unsigned)kz, 2), 128, 0, c.stream>>>(a);
    MB_CUDA(cudaGetLastError());
  }
  if (a.ichem) {   // chem_bdyval_uncoupled reads the u, v the four sides have just set
    if (g.bl || g.br) {
      LaunchScope ls(c, KID_BDYVAL);
      const int n = g.ice2 - g.ice1 + 1;
      moloch_chem_bdyval_we<<<dim3((unsigned)((n + 127) / 128), (unsigned)(kz * g.ntr), 2), 128, 0, c.stream>>>(a);
      MB_CUDA(cudaGetLastError());
    }
    if (g.bb || g.bt) {
      LaunchScope ls(c, KID_BDYVAL);
      const int n = g.jce2 - g.jce1 + 1;
      moloch_chem_bdyval_sn<<<dim3((unsigned)((n + 127) / 128), (unsigned)(kz * g.ntr), 2), 128, 0, c.stream>>>(a);
      MB_CUDA(cudaGetLastError());
    }
  }
  return 0;
}

// ---- motopnudge + morelax of all variables -----------------------------------------
__global__ void moloch_bdy_relax(BdyArgs a) {
  const int j = a.g.jde1 + blockIdx.x * BX + threadIdx.x;
  const int i = a.g.ide1 + blockIdx.y * BY + threadIdx.y;
  if (j > a.g.jde2 || i > a.g.ide2) return;
  bdy_relax_cell(a, j, i, 1 + (int)blockIdx.z);
}
int k_bdy_relax(Ctx& c, double xbctime) {
  const Geo& g = c.g;
  if (!(c.cfg.mo_top_nudge || c.cfg.nspgx > 0)) return 0;
  const BdyArgs a = bdy_args(c, xbctime);
  LaunchScope ls(c, KID_BDYRELAX);
  moloch_bdy_relax<<<grid3(g.jde2 - g.jde1 + 1, g.ide2 - g.ide1 + 1, g.kz), dim3(BX, BY), 0, c.stream>>>(a);
  MB_CUDA(cudaGetLastError());
  return 0;
}

// ---- uvstagtouvx + temp_to_tvirt + tetav ---------------------------------------------
__global__ void moloch_bdy_finish(BdyArgs a) {
  const int j = a.g.jce1 + blockIdx.x * BX + threadIdx.x;
  const int i = a.g.ice1 + blockIdx.y * BY + threadIdx.y;
  if (j > a.g.jce2 || i > a.g.ice2) return;
  bdy_finish_cell(a, j, i, 1 + (int)blockIdx.z);
}
int k_bdy_finish(Ctx& c) {
  const Geo& g = c.g;
  const BdyArgs a = bdy_args(c, 0.0);
  LaunchScope ls(c, KID_BDYFINISH);
  moloch_bdy_finish<<<grid3(g.jce2 - g.jce1 + 1, g.ice2 - g.ice1 + 1, g.kz), dim3(BX, BY), 0, c.stream>>>(a);
  MB_CUDA(cudaGetLastError());
  return 0;
}

// ---- mkslice -----------------------------------------------------------------------------
__global__ void moloch_mkslice(SliceArgs a) {
  const int j = a.g.jce1 + blockIdx.x * BX + threadIdx.x;
  const int i = a.g.ice1 + blockIdx.y * BY + threadIdx.y;
  if (j > a.g.jce2 || i > a.g.ice2) return;
  mkslice_cell(a, j, i, 1 + (int)blockIdx.z);
}
__global__ void moloch_mkslice_col(SliceArgs a) {
  const int j = a.g.jce1 + blockIdx.x * BX + threadIdx.x;
  const int i = a.g.ice1 + blockIdx.y * BY + threadIdx.y;
  if (j > a.g.jce2 || i > a.g.ice2) return;
  mkslice_col(a, j, i);
  mkslice_trop_col(a, j, i);
}
int k_mkslice(Ctx& c) {
  const Geo& g = c.g;
  SliceArgs a;
  memset(&a, 0, sizeof(a));
  a.g = g;
  a.pai = c.f[MB_PAI].p; a.t = c.f[MB_T].p; a.p = c.f[MB_P].p; a.rho = c.f[MB_RHO].p; a.qsat = c.f[MB_QSAT].p;
  a.w = c.f[MB_W].p; a.ps = c.f[MB_PS].p; a.zq = c.f[MB_ZETAF].p; a.qx = c.f[MB_QX].p; a.trac = c.f[MB_TRAC].p;
  a.pf3d = c.f[MB_PF3D].p; a.th3d = c.f[MB_TH3D].p; a.rhb3d = c.f[MB_RHB3D].p; a.wpx3d = c.f[MB_WPX3D].p;
  a.rhox2d = c.f[MB_RHOX2D].p; a.tp2d = c.f[MB_TP2D].p; a.th700 = c.f[MB_TH700].p;
  a.rhmin = c.cfg.rhmin; a.rhmax = c.cfg.rhmax;
  a.ichem = c.cfg.ichem && c.cfg.ntr > 0; a.icldmstrat = c.cfg.icldmstrat;
  a.xlat = c.f[MB_XLAT].p; a.za = c.f[MB_ZETA].p; a.ptrop = c.f[MB_PTROP].p; a.ktrop = c.f[MB_KTROP].p;
  a.kmxpbl = c.f[MB_KMXPBL].p; a.calday = c.calday; a.dayspy = c.dayspy; a.irceideal = c.cfg.irceideal;
  a.ibltyp = c.cfg.ibltyp;
  {
    LaunchScope ls(c, KID_MKSLICE);
    moloch_mkslice<<<grid3(g.jce2 - g.jce1 + 1, g.ice2 - g.ice1 + 1, g.kz), dim3(BX, BY), 0, c.stream>>>(a);
    MB_CUDA(cudaGetLastError());
  }
  {
    LaunchScope ls(c, KID_MKSLICE);
    moloch_mkslice_col<<<grid3(g.jce2 - g.jce1 + 1, g.ice2 - g.ice1 + 1, 1), dim3(BX, BY), 0, c.stream>>>(a);
    MB_CUDA(cudaGetLastError());
  }
  return 0;
}

// ---- massck partial sums and the ps guard ------------------------------------------------------
// A diagnostic that runs when RegCM's debug_level > 0 / every syncro_rep: one
// thread per row walks along j in the reference's order (deterministic sums;
// the strided access costs a fraction of a millisecond at 400x400x41).
__global__ void moloch_massck_rows(MassArgs a) {
  const int i = a.g.ice1 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i > a.g.ice2) return;
  massck_row(a, i, 1 + (int)blockIdx.y);
}
__global__ void moloch_massck_bdy(MassArgs a) {
  const int k = 1 + blockIdx.x * blockDim.x + threadIdx.x;
  if (k > a.g.kz) return;
  massck_bdy_level(a, k);
}
__global__ void moloch_ps_rows(MassArgs a) {
  const int i = a.g.ice1 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i > a.g.ice2) return;
  ps_row(a, i);
}
__global__ void moloch_massck_final(MassArgs a) { massck_final(a); }
// what: 1 massck sums, 2 ps guard, 3 both; out7 (host): tdrym, tdadv, tqmass, tqadv, psmax, psmin, nonfinite
int k_massck(Ctx& c, int what, double* out7) {
  const Geo& g = c.g;
  const int ni = g.ice2 - g.ice1 + 1, kz = g.kz;
  const size_t need = (size_t)2 * kz * ni + 4 * kz + 3 * ni + 8;
  if (c.mass_work_doubles < need) {
    if (c.mass_work) cudaFree(c.mass_work);
    MB_CUDA(cudaMalloc(&c.mass_work, need * sizeof(double)));
    c.mass_work_doubles = need;
  }
  MB_CUDA(cudaMemsetAsync(c.mass_work, 0, need * sizeof(double), c.stream));
  MassArgs a;
  memset(&a, 0, sizeof(a));
  a.g = g;
  a.rho = c.f[MB_RHO].p; a.zq = c.f[MB_ZETAF].p; a.qx = c.f[MB_QX].p; a.u = c.f[MB_U].p; a.v = c.f[MB_V].p;
  a.ps = c.f[MB_PS].p;
  a.rows = c.mass_work; a.lev = a.rows + (size_t)2 * kz * ni; a.psrow = a.lev + 4 * kz; a.out = a.psrow + 3 * ni;
  a.dxsq = c.cfg.dx * c.cfg.dx; a.dt = c.cfg.dtsec; a.dx = c.cfg.dx; a.ni = ni;
  if (what & 1) {
    if (!a.zq) return fail("massck: zq (MB_ZETAF) is not on the device (moloch_b200_config.do_massck)");
    {
      LaunchScope ls(c, KID_MASSCK);
      moloch_massck_rows<<<dim3((unsigned)((ni + 63) / 64), (unsigned)kz), 64, 0, c.stream>>>(a);
      MB_CUDA(cudaGetLastError());
    }
    {
      LaunchScope ls(c, KID_MASSCK);
      moloch_massck_bdy<<<(unsigned)((kz + 63) / 64), 64, 0, c.stream>>>(a);
      MB_CUDA(cudaGetLastError());
    }
  }
  if (what & 2) {
    LaunchScope ls(c, KID_MASSCK);
    moloch_ps_rows<<<(unsigned)((ni + 63) / 64), 64, 0, c.stream>>>(a);
    MB_CUDA(cudaGetLastError());
  }
  {
    LaunchScope ls(c, KID_MASSCK);
    moloch_massck_final<<<1, 1, 0, c.stream>>>(a);
    MB_CUDA(cudaGetLastError());
  }
  MB_CUDA(cudaMemcpyAsync(out7, a.out, 7 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  if (sync_stream(c)) return 1;
  return 0;
}

// ---- tendency diagnostics ---------------------------------------------------------------------------
__global__ void moloch_diag_snap(DiagArgs a) {
  const int j = a.g.jci1 + blockIdx.x * BX + threadIdx.x, i = a.g.ici1 + blockIdx.y * BY + threadIdx.y;
  if (j > a.g.jci2 || i > a.g.ici2) return;
  diag_snap_cell(a, j, i, 1 + (int)blockIdx.z);
}
__global__ void moloch_diag_diff(DiagArgs a) {
  const int j = a.g.jci1 + blockIdx.x * BX + threadIdx.x, i = a.g.ici1 + blockIdx.y * BY + threadIdx.y;
  if (j > a.g.jci2 || i > a.g.ici2) return;
  diag_diff_cell(a, j, i, 1 + (int)blockIdx.z);
}
// which: 0 = dynamical_core (adh / cadvhdiag), 1 = boundary (bdy / cbdydiag); diff: false = snapshot
int k_diag(Ctx& c, int which, bool diff) {
  const Geo& g = c.g;
  DiagArgs a;
  memset(&a, 0, sizeof(a));
  a.g = g;
  a.idiag = c.cfg.idiag > 0; a.ichdiag = c.cfg.ichdiag > 0 && c.cfg.ichem && c.cfg.ntr > 0;
  if (!a.idiag && !a.ichdiag) return 0;
  a.t = c.f[MB_T].p; a.qv = c.f[MB_QX].p; a.trac = c.f[MB_TRAC].p;
  a.ten0 = c.f[MB_TEN0].p; a.qen0 = c.f[MB_QEN0].p; a.chiten0 = c.f[MB_CHITEN0].p;
  a.dt_out = c.f[which ? MB_TDIAG_BDY : MB_TDIAG_ADH].p; a.dq_out = c.f[which ? MB_QDIAG_BDY : MB_QDIAG_ADH].p;
  a.dc_out = c.f[which ? MB_CBDYDIAG : MB_CADVHDIAG].p;
  a.rdt = 1.0 / c.cfg.dtsec;       // rdt, Main/mpplib/mod_runparams.F90
  LaunchScope ls(c, KID_DIAGTEN);
  const dim3 grid = grid3(g.jci2 - g.jci1 + 1, g.ici2 - g.ici1 + 1, g.kz);
  if (diff) moloch_diag_diff<<<grid, dim3(BX, BY), 0, c.stream>>>(a);
  else moloch_diag_snap<<<grid, dim3(BX, BY), 0, c.stream>>>(a);
  MB_CUDA(cudaGetLastError());
  return 0;
}

// ---- TKE helpers (ibltyp == 2) ---------------------------------------------------------------
__global__ void moloch_zstagtoh(Geo g, const double* __restrict__ fl, double* __restrict__ hl) {
  const int j = g.jce1 + blockIdx.x * BX + threadIdx.x;
  const int i = g.ice1 + blockIdx.y * BY + threadIdx.y;
  if (j > g.jce2 || i > g.ice2) return;
  zstagtoh_cell(g, fl, hl, j, i, 1 + (int)blockIdx.z);
}
__global__ void moloch_htozstag(Geo g, const double* __restrict__ hl, double* __restrict__ fl) {
  const int j = g.jce1 + blockIdx.x * BX + threadIdx.x;
  const int i = g.ice1 + blockIdx.y * BY + threadIdx.y;
  if (j > g.jce2 || i > g.ice2) return;
  htozstag_cell(g, hl, fl, j, i, 2 + (int)blockIdx.z);
}
__global__ void moloch_tke_update(Geo g, double* __restrict__ tke, const double* __restrict__ tketen, double dtinc,
                                  double tkemin) {
  const int j = g.jci1 + blockIdx.x * BX + threadIdx.x;
  const int i = g.ici1 + blockIdx.y * BY + threadIdx.y;
  if (j > g.jci2 || i > g.ici2) return;
  tke_update_cell(g, tke, tketen, dtinc, tkemin, j, i, 1 + (int)blockIdx.z);
}
int k_tke_destagger(Ctx& c) {
  const Geo& g = c.g;
  LaunchScope ls(c, KID_TKE);
  moloch_zstagtoh<<<grid3(g.jce2 - g.jce1 + 1, g.ice2 - g.ice1 + 1, g.kz), dim3(BX, BY), 0, c.stream>>>(
      g, c.f[MB_TKE].p, c.f[MB_TKEX].p);
  MB_CUDA(cudaGetLastError());
  return 0;
}
int k_tke_restagger(Ctx& c) {
  const Geo& g = c.g;
  LaunchScope ls(c, KID_TKE);
  moloch_htozstag<<<grid3(g.jce2 - g.jce1 + 1, g.ice2 - g.ice1 + 1, g.kz - 1), dim3(BX, BY), 0, c.stream>>>(
      g, c.f[MB_TKEX].p, c.f[MB_TKE].p);
  MB_CUDA(cudaGetLastError());
  return 0;
}
int k_tke_update(Ctx& c, double dtinc) {
  const Geo& g = c.g;
  LaunchScope ls(c, KID_TKE);
  moloch_tke_update<<<grid3(g.jci2 - g.jci1 + 1, g.ici2 - g.ici1 + 1, g.kz + 1), dim3(BX, BY), 0, c.stream>>>(
      g, c.f[MB_TKE].p, c.f[MB_TKETEN].p, dtinc, c.cfg.tkemin);
  MB_CUDA(cudaGetLastError());
  return 0;
}

// ibnd upload: int32 host box -> padded device plane (cells outside the box stay -1)
__global__ void moloch_ibnd_fill(Geo g, int* __restrict__ dst, const int* __restrict__ src, int jlo, int jhi,
                                 int ilo, int ihi) {
  const long long n = g.plane;
  for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < n;
       id += (long long)gridDim.x * blockDim.x) {
    const int j = g.j0 + (int)(id % g.NJ), i = g.i0 + (int)(id / g.NJ);
    int v = -1;
    if (j >= jlo && j <= jhi && i >= ilo && i <= ihi) v = src[(long long)(i - ilo) * (jhi - jlo + 1) + (j - jlo)];
    dst[id] = v;
  }
}
int k_ibnd_fill(Ctx& c, int* dst, const int* src, int jlo, int jhi, int ilo, int ihi) {
  LaunchScope ls(c, KID_INIT);
  moloch_ibnd_fill<<<148, 256, 0, c.stream>>>(c.g, dst, src, jlo, jhi, ilo, ihi);
  MB_CUDA(cudaGetLastError());
  return 0;
}

// ---- mospectral_nudge ---------------------------------------------------------------------
// Six small launches per variable (t, u, v); every level is handled in parallel,
// every sum keeps the reference's order (bdy_cells.h).  Runs every dtrad only.
// On more than one rank row_reduce / column_reduce (MPI_Allreduce over the ranks
// with the same loci / locj, Main/mpplib/mod_mppparam.F90:1459-1469, 20618-20664)
// become halo_group_sum: the partial sums of all levels in one all-gather over
// NCCL, added in rank order (MPI leaves the order open; rank order is what the
// oracle's emulation uses, so the decomposed run equals it bit for bit).
__global__ void moloch_spec_zn(SpecArgs a) {
  const int j = a.j1 + blockIdx.x * BX + threadIdx.x, i = a.i1 + blockIdx.y * BY + threadIdx.y;
  if (j > a.j2 || i > a.i2) return;
  spec_zn_cell(a, j, i, 1 + (int)blockIdx.z);
}
__global__ void moloch_spec_sx(SpecArgs a) {   // x: i, y: kk, z: level
  const int i = a.i1 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i > a.i2) return;
  spec_sx_cell(a, i, 1 + (int)blockIdx.y, 1 + (int)blockIdx.z);
}
__global__ void moloch_spec_g1(SpecArgs a) {
  const int j = a.j1 + blockIdx.x * BX + threadIdx.x, i = a.i1 + blockIdx.y * BY + threadIdx.y;
  if (j > a.j2 || i > a.i2) return;
  spec_g1_cell(a, j, i, 1 + (int)blockIdx.z);
}
__global__ void moloch_spec_sy(SpecArgs a) {   // x: j, y: l, z: level
  const int j = a.j1 + blockIdx.x * blockDim.x + threadIdx.x;
  if (j > a.j2) return;
  spec_sy_cell(a, j, 1 + (int)blockIdx.y, 1 + (int)blockIdx.z);
}
__global__ void moloch_spec_update(SpecArgs a) {
  const int j = a.jj1 + blockIdx.x * BX + threadIdx.x, i = a.ii1 + blockIdx.y * BY + threadIdx.y;
  if (j > a.jj2 || i > a.ii2) return;
  spec_update_cell(a, j, i, 1 + (int)blockIdx.z);
}
int k_spectral_nudge(Ctx& c, double xbctime) {
  const Geo& g = c.g;
  const int kz = g.kz, km2 = 2 * c.cfg.km, lm2 = 2 * c.cfg.lm;
  const int ni = g.ide2 - g.ide1 + 1, nj = g.jde2 - g.jde1 + 1;
  const size_t nsx = (size_t)kz * km2 * ni, nsy = (size_t)kz * lm2 * nj;
  const size_t need = nsx + nsy + (size_t)km2 * ni + (size_t)lm2 * nj;
  if (c.spec_work_doubles < need) {
    if (c.spec_work) cudaFree(c.spec_work);
    MB_CUDA(cudaMalloc(&c.spec_work, need * sizeof(double)));
    // getmem zero-initialises sx, sxg, sy, syg (Share/mod_space.F90:304-311)
    MB_CUDA(cudaMemsetAsync(c.spec_work, 0, need * sizeof(double), c.stream));
    c.spec_work_doubles = need;
  }
  // ranks of this rank's row (same loci) and column (same locj); rank = locj*niycpus + loci
  std::vector<int> rowm, colm;
  if (c.cfg.nranks > 1) {
    const int py = c.cfg.niycpus, px = c.cfg.nranks / py;
    const int locj = c.cfg.rank / py, loci = c.cfg.rank % py;
    for (int q = 0; q < px; ++q) rowm.push_back(q * py + loci);
    for (int q = 0; q < py; ++q) colm.push_back(locj * py + q);
  }
  SpecArgs a;
  memset(&a, 0, sizeof(a));
  a.g = g;
  a.zn = c.wzall; a.g1 = c.wzall + (size_t)kz * g.plane;   // free outside `advection`
  a.sx = c.spec_work; a.sy = a.sx + nsx;
  double* sx_stale = a.sy + nsy; double* sy_stale = sx_stale + (size_t)km2 * ni;
  a.sx_stale = sx_stale; a.sy_stale = sy_stale;
  a.bvx = c.tab[MB_TAB_BVX]; a.bvy = c.tab[MB_TAB_BVY]; a.cnudge = c.tab[MB_TAB_CNUDGE];
  const double rtb = 1.0 / c.cfg.dtbdys;
  a.x1 = (xbctime + c.cfg.dtsec) * rtb; a.x0 = 1.0 - a.x1;
  a.km2 = km2; a.lm2 = lm2; a.ni = ni; a.nj = nj;
  const int fid[3] = {MB_T, MB_U, MB_V}, b0[3] = {MB_XTB0, MB_DUB0, MB_DVB0}, b1[3] = {MB_XTB1, MB_DUB1, MB_DVB1};
  for (int var = 0; var < 3; ++var) {
    a.f = c.f[fid[var]].p; a.b0 = c.f[b0[var]].p; a.b1 = c.f[b1[var]].p;
    spec_ranges(g, var, a);
    const int nbj = a.j2 - a.j1 + 1, nbi = a.i2 - a.i1 + 1;
    {
      LaunchScope ls(c, KID_SPECTRAL);
      moloch_spec_zn<<<grid3(nbj, nbi, kz), dim3(BX, BY), 0, c.stream>>>(a);
      MB_CUDA(cudaGetLastError());
    }
    {
      LaunchScope ls(c, KID_SPECTRAL);
      moloch_spec_sx<<<dim3((unsigned)((nbi + 63) / 64), (unsigned)km2, (unsigned)kz), 64, 0, c.stream>>>(a);
      MB_CUDA(cudaGetLastError());
    }
    if (rowm.size() > 1 && halo_group_sum(c, a.sx, nsx, rowm.data(), (int)rowm.size())) return 1;   // row_reduce
    if (a.count_x == km2 * ni)   // a full-length reduction: this is what later short calls find in the tail of sxg
      MB_CUDA(cudaMemcpyAsync(sx_stale, a.sx + (size_t)(kz - 1) * km2 * ni, (size_t)km2 * ni * sizeof(double),
                              cudaMemcpyDeviceToDevice, c.stream));
    {
      LaunchScope ls(c, KID_SPECTRAL);
      moloch_spec_g1<<<grid3(nbj, nbi, kz), dim3(BX, BY), 0, c.stream>>>(a);
      MB_CUDA(cudaGetLastError());
    }
    {
      LaunchScope ls(c, KID_SPECTRAL);
      moloch_spec_sy<<<dim3((unsigned)((nbj + 63) / 64), (unsigned)lm2, (unsigned)kz), 64, 0, c.stream>>>(a);
      MB_CUDA(cudaGetLastError());
    }
    if (colm.size() > 1 && halo_group_sum(c, a.sy, nsy, colm.data(), (int)colm.size())) return 1;   // column_reduce
    if (a.count_y == lm2 * nj)
      MB_CUDA(cudaMemcpyAsync(sy_stale, a.sy + (size_t)(kz - 1) * lm2 * nj, (size_t)lm2 * nj * sizeof(double),
                              cudaMemcpyDeviceToDevice, c.stream));
    {
      LaunchScope ls(c, KID_SPECTRAL);
      moloch_spec_update<<<grid3(a.jj2 - a.jj1 + 1, a.ii2 - a.ii1 + 1, kz), dim3(BX, BY), 0, c.stream>>>(a);
      MB_CUDA(cudaGetLastError());
    }
  }
  return 0;
}

}  // namespace mb
