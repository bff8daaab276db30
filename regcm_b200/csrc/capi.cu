// capi.cu -- the C ABI of include/moloch_b200.h: context, device memory,
// host<->device hand-off and the orchestration of one MOLOCH step
// (reference call tree: Main/mod_moloch.F90:312-446, 545-736, 767-836,
// 1085-1141, 1403-1443).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <utility>
#include "common.cuh"
#ifndef MB_HOST_EMU
#include <nvtx3/nvToolsExt.h>
#endif

namespace mb {

#ifndef MB_HOST_EMU
NvtxRange::NvtxRange(const char* name) { nvtxRangePushA(name); }
NvtxRange::~NvtxRange() { nvtxRangePop(); }
#endif

thread_local std::string g_err;
int fail(const std::string& msg) { g_err = msg; return 1; }

static const char* k_names[KID_COUNT] = {
    "reset_tendencies", "tetavf_init", "sound_pre", "divdamp_filter", "wsolve", "uvupdate", "sfinish",
    "destagger", "waf_vertical", "waf_meridional", "waf_zonal", "curvature", "restagger", "tvirt_temp",
    "diag_prq", "diag_ps", "status_update", "halo_local", "halo_pack", "halo_unpack", "init_static",
    "waf_horizontal", "box_copy", "bdyval", "bdy_relax", "bdy_finish", "mkslice", "tke", "spectral_nudge", "massck", "diag_tendencies"};
const char* kernel_name(int kid) { return (kid >= 0 && kid < KID_COUNT) ? k_names[kid] : "?"; }

LaunchScope::LaunchScope(Ctx& c_, int kid_) : c(c_), kid(kid_) {
  c.launches++;
  if (c.profiling) {
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, c.stream);
  }
}
LaunchScope::~LaunchScope() {
  if (a) {
    cudaEventRecord(b, c.stream);
    c.events.push_back(ProfEvent{a, b, kid});
  }
}

// number of wafone-advected fields, in the reference's order :786-807
static int count_adv(const moloch_b200_config& f) {
  int n = 6;
  if (f.ipptls > 0) n += (f.nqx - f.iqfrst + 1 > 0) ? f.nqx - f.iqfrst + 1 : 0;
  if (f.ibltyp == 2) n += 1;   // tkex :799-801
  return n + f.ntr;
}

// The arena layout is a pure function of a rank's config, so that a rank can
// address the arrays of its neighbours inside their (peer-mapped) arenas.
Layout make_layout(const moloch_b200_config& f) {
  const Geo g = geo_from_cfg(f);
  const int kz = g.kz, nadv = count_adv(f);
  const size_t pl = (size_t)g.plane * sizeof(double);
  auto al = [](size_t b) { return (b + 255) / 256 * 256; };
  Layout L;
  L.off.assign(SL_COUNT, 0); L.size.assign(SL_COUNT, 0);
  size_t total = 0;
  auto put = [&](int slot, size_t bytes) { L.off[slot] = total; L.size[slot] = bytes; total += al(bytes); };
  for (int id = 0; id < MB_NFIELDS; ++id) {
    if (id == MB_WZ || id == MB_P0) continue;
    int nk, nspec; field_shape(f, id, nk, nspec);
    put(id, pl * nk * (size_t)nspec);
  }
  put(SL_UD, 0); put(SL_VD, 0); put(SL_ZB, pl * kz);   // ud, vd: gone with round 2's uvupdate2
  L.stride2d = al(pl); put(SL_2D, L.stride2d * 4);
  L.stridezr = al(pl * kz); put(SL_ZR, L.stridezr * 2);
  put(SL_WZ, pl * kz * (size_t)nadv); put(SL_P0, pl * kz * (size_t)nadv);
  const size_t prof_len = (size_t)((kz + 2 > (g.ide2 - g.ide1 + 3)) ? kz + 2 : (g.ide2 - g.ide1 + 3));
  L.strideprof = al(prof_len * sizeof(double)); put(SL_PROF, L.strideprof * MB_NPROFILES);
  put(SL_TAB, sizeof(double*) * (size_t)nadv);
  put(SL_FLAGS, 256);
  L.total = total;
  return L;
}

static int check_cfg(const moloch_b200_config& f) {
  if (f.jx < 4 || f.iy < 4 || f.kz < 4) return fail("moloch_b200_create: jx, iy, kz must be >= 4");
  if (f.nqx < 1 || f.nqx > 10) return fail("moloch_b200_create: nqx must be in 1..10");
  if (f.ntr < 0) return fail("moloch_b200_create: ntr < 0");
  if (f.ipptls > 1 && f.nqx < 5) return fail("moloch_b200_create: ipptls=2 needs nqx >= 5");
  if (f.ipptls == 1 && f.nqx < 2) return fail("moloch_b200_create: ipptls=1 needs nqx >= 2");
  if (f.jde2 - f.jde1 + 1 < 3 || f.ide2 - f.ide1 + 1 < 3)
    return fail("Cannot have one processor with less than 3x3 points");  // mod_mppparam.F90:1605
  if (f.jce1 != f.jde1 || f.ice1 != f.ide1 || f.jce2 > f.jde2 || f.ice2 > f.ide2 || f.jce2 < f.jde2 - 1 ||
      f.ice2 < f.ide2 - 1)
    return fail("moloch_b200_create: inconsistent cross/dot ranges");
  if (f.mo_nadv < 1 || f.mo_nsound < 1) return fail("moloch_b200_create: mo_nadv, mo_nsound must be >= 1");
  if (!(f.dtsec > 0.0) || !(f.dx > 0.0) || !(f.mo_dzita > 0.0))
    return fail("moloch_b200_create: dtsec, dx, mo_dzita must be > 0");
  if (f.nranks < 1 || f.rank < 0 || f.rank >= f.nranks) return fail("moloch_b200_create: bad rank/nranks");
  if (f.ibltyp == 2 && !(f.tkemin >= 0.0)) return fail("moloch_b200_create: ibltyp=2 needs tkemin >= 0");
  if (f.do_bdy) {
    if (!(f.dtbdys > 0.0)) return fail("moloch_b200_create: do_bdy needs dtbdys > 0");
    if (f.nspgx < 0 || f.nspgx == 1 || f.nspgx == 2) return fail("moloch_b200_create: nspgx must be 0 or >= 3");
    if (f.mo_top_nudge && (f.nztop < 0 || f.nztop > f.kz)) return fail("moloch_b200_create: nztop out of range");
    if (f.mo_spectral_nudge) {
      if (!(f.dtrad > 0.0) || f.km < 1 || f.lm < 1)
        return fail("moloch_b200_create: mo_spectral_nudge needs dtrad > 0 and km, lm >= 1");
      if (f.nranks > 1 && (f.niycpus < 1 || f.nranks % f.niycpus != 0))
        return fail("moloch_b200_create: mo_spectral_nudge on more than one rank needs niycpus (cpus_per_dim(2)) "
                    "for row_reduce/column_reduce (Main/mpplib/mod_mppparam.F90:20618-20664)");
    }
  }
  if (f.do_slice && !(f.rhmax >= f.rhmin)) return fail("moloch_b200_create: do_slice needs rhmin <= rhmax");
  return 0;
}

int halo_timeout_check(Ctx& c) {
  if (c.p2p && c.flags) {   // a wait of the peer-store transport gave up: the ghost cells of that round are stale
    unsigned long long t = 0;
    MB_CUDA(cudaMemcpy(&t, c.flags + 5, sizeof(t), cudaMemcpyDeviceToHost));
    if (t != 0ULL)
      return fail("halo exchange timed out in round " + std::to_string(t) + " of this rank: a neighbour never "
                  "arrived (ranks out of step, or a rank died); results after that round are invalid");
  }
  return 0;
}
int sync_stream(Ctx& c) {
  MB_CUDA(cudaStreamSynchronize(c.stream));
  return halo_timeout_check(c);
}

// ---- orchestration ----------------------------------------------------------
// Ghost cells of u and v as the sound loop needs them.  The reference exchanges u left/right and v
// bottom/top, one point (:570-571); moloch_sound_div also evaluates zdiv2 on the one-cell ring around
// the rank (instead of receiving it, :745/:535), which reads u two columns / one row and v one column /
// two rows beyond the owned box, corner ghosts included.  Corners only exist on a 2-D decomposition:
// there the rows travel in a second round that carries the ghost columns of the first along.
// Advection's own exchange (:1532-1533: u left/right 2, v bottom/top 2) is a subset.
static int exchange_uv_wide(Ctx& c, const HaloItem* extra, int nextra) {
  const int kz = c.g.kz;
  const moloch_b200_config& f = c.cfg;
  const bool has_lr = f.nbr_left >= 0 || f.nbr_right >= 0, has_bt = f.nbr_bottom >= 0 || f.nbr_top >= 0;
  const HaloItem iu = {c.f[MB_U].p, kz}, iv = {c.f[MB_V].p, kz};
  HaloSpec sp[5];
  int n = 0;
  if (nextra > 0) sp[n++] = {extra, nextra, HS_CROSS, 1, true, true, 0};
  if (has_lr && has_bt) {
    sp[n++] = {&iu, 1, HS_U, 2, true, false, 0};
    sp[n++] = {&iv, 1, HS_V, 1, true, false, 0};
    if (halo_exchange_multi(c, sp, n)) return 1;
    const HaloSpec s2[2] = {{&iu, 1, HS_U, 1, false, true, 2}, {&iv, 1, HS_V, 2, false, true, 1}};
    return halo_exchange_multi(c, s2, 2);
  }
  sp[n++] = {&iu, 1, HS_U, 2, true, false, 0};
  sp[n++] = {&iv, 1, HS_V, 1, true, false, 0};
  sp[n++] = {&iu, 1, HS_U, 1, false, true, 0};
  sp[n++] = {&iv, 1, HS_V, 2, false, true, 0};
  return halo_exchange_multi(c, sp, n);
}

// for_adv: called from dynamical_core, `advection` follows at once -- with the fused transport the last
// sub-step's uvupdate then delivers advection's 2-wide u, v ghosts (:1532-1533) and destagger waits for them
static int do_sound(Ctx& c, bool for_adv = false) {
  NvtxRange nvtx("sound");
  const double dts = c.dtsound;
  const int kz = c.g.kz;
  const int nsound = c.cfg.mo_nsound;
  const moloch_b200_config& cf = c.cfg;
  // Peer-store transport: the two exchanges of a sub-step (pai :673; u, v :570-571) are fused into the
  // kernels around them (see common.cuh): wsolve / uvupdate2 store the edge cells they compute into the
  // neighbours' ghost cells, uvupdate2 / sound_div wait for the neighbours' word.  The edge cells the
  // sound kernels never update (physical-boundary rows and columns, whose values the advection and the
  // boundary update change between two sound calls) travel once, in the full round at the start of the
  // call.  u, v pushes are fused on 1-D decompositions only (no corner ghosts there).
  const bool fused = halo_fused_available(c);
  const bool fused_all = fused && c.fuse_level >= 2;   // also the first sub-step (and pai in the full round)
  const bool one_d = !((cf.nbr_left >= 0 || cf.nbr_right >= 0) && (cf.nbr_bottom >= 0 || cf.nbr_top >= 0));
  const bool fused_uv = fused && one_d;
  if (!(fused_all && fused_uv)) for_adv = false;
  HaloItem it;
  {
    // tetav (:562) together with the first sub-step's u, v: tetavf_init in between is a column operation
    // on owned cells.  Fused transport: pai too (its :673 exchange of every sub-step becomes wsolve's push).
    const HaloItem itv[2] = {{c.f[MB_TETAV].p, kz}, {c.f[MB_PAI].p, kz}};
    if (exchange_uv_wide(c, itv, fused_all ? 2 : 1)) return 1;
  }
  c.adv_wait_valid = false;
  if (k_tetavf_init(c)) return 1;
  WaitCtl w_uv = {}, w_pai = {};
  PushCtl p_uv = {}, p_pai = {};
  EdgePush e_u = {}, e_v = {}, e_pai = {};
  bool uv_pushed = false;   // u, v ghosts were delivered by the previous sub-step's uvupdate2
  for (int ns = 0; ns < nsound; ++ns) {
    const bool f = fused && (fused_all || ns > 0);
    if (!uv_pushed && ns > 0 && exchange_uv_wide(c, nullptr, 0)) return 1;
    if (k_sound_div(c, dts, uv_pushed ? &w_uv : nullptr)) return 1;
    if (f) {
      if (halo_fused_begin(c, &p_pai, &w_pai)) return 1;
      if (halo_fused_edge(c, c.f[MB_PAI].p, HS_CROSS, true, true, &e_pai)) return 1;
    }
    if (k_wsolve(c, dts, ns == nsound - 1, f ? &p_pai : nullptr, f ? &e_pai : nullptr)) return 1;
    if (!f) {
      it = {c.f[MB_PAI].p, kz};
      if (halo_exchange(c, &it, 1, HS_CROSS, 1, true, true)) return 1;  // :673
    }
    // this uvupdate2 delivers the next sub-step's u, v ghosts
    const bool push_uv = fused_uv && (ns + 1 < nsound || for_adv);
    if (push_uv) {
      if (halo_fused_begin(c, &p_uv, &w_uv)) return 1;
      if (halo_fused_edge(c, c.f[MB_U].p, HS_U, true, true, &e_u, 2, 1)) return 1;
      if (halo_fused_edge(c, c.f[MB_V].p, HS_V, true, true, &e_v, 1, 2)) return 1;
    }
    if (k_uvupdate2(c, dts, f ? &w_pai : nullptr, push_uv ? &p_uv : nullptr, push_uv ? &e_u : nullptr,
                    push_uv ? &e_v : nullptr)) return 1;
    uv_pushed = push_uv;
  }
  if (uv_pushed) { c.adv_wait = w_uv; c.adv_wait_valid = true; }   // consumed by advection's destagger
  return 0;  // :728-734 (finish of s) is fused into the last sub-step's wsolve
}

static int do_wafone_range(Ctx& c, int first, int count) {
  const double dta = c.dtstepa;
  const int kz = c.g.kz;
  const long long fsz = (long long)kz * c.g.plane;
  std::vector<HaloItem> items((size_t)count);
  if (c.waf_impl == 2) {
    // Field-batched path.  The fused horizontal kernel recomputes p0 on the two
    // ghost columns instead of receiving it (:955/:1012), so it needs wz with
    // corner ghosts and pp on two ghost columns: 3 batched messages per side
    // instead of the reference's 2 per field.
    for (int q = 0; q < count; ++q) items[q] = {c.wzall + q * fsz, kz};
    // Decomposition along i only, peer-store transport, fusion level 2: exchange_bt(wz, 2) (:924) is fused into
    // the two kernels -- the vertical kernel stores the two edge rows of every field into the neighbours' ghost
    // rows while it computes them, the horizontal kernel's edge CTAs wait for the neighbours' word.
    const moloch_b200_config& cf = c.cfg;
    const bool lr_nbr = cf.nbr_left >= 0 || cf.nbr_right >= 0;
    const bool bt_nbr = cf.nbr_bottom >= 0 || cf.nbr_top >= 0;
    if (c.fuse_wz && c.fuse_level >= 2 && halo_fused_available(c) && bt_nbr && !lr_nbr && first == 0) {
      PushCtl p_wz = {};
      WaitCtl w_wz = {};
      EdgePush e_wz = {};
      if (halo_fused_begin(c, &p_wz, &w_wz)) return 1;
      if (halo_fused_edge(c, c.wzall, HS_CROSS, false, true, &e_wz, 2, 2)) return 1;
      if (k_waf_z2(c, first, count, dta, &p_wz, &e_wz)) return 1;
      return k_waf_yx(c, first, count, dta, &w_wz);
    }
    if (k_waf_z2(c, first, count, dta)) return 1;
    if (halo_exchange(c, items.data(), count, HS_CROSS, 2, false, true)) return 1;
    // second round, left/right incl. the ghost rows (corners): wz, the
    // pre-advection pp snapshot written by the vertical kernel, and v (whose
    // ghost rows came with :1533)
    std::vector<HaloItem> snap((size_t)count);
    for (int q = 0; q < count; ++q) snap[q] = {c.p0all + q * fsz, kz};
    const HaloItem iv = {c.f[MB_V].p, kz};
    const HaloSpec sp[3] = {{items.data(), count, HS_CROSS, 2, true, false, 2},
                            {snap.data(), count, HS_CROSS, 2, true, false, 2},
                            {&iv, 1, HS_CROSS, 2, true, false, 2}};
    if (halo_exchange_multi(c, sp, 3)) return 1;
    return k_waf_yx(c, first, count, dta);
  }
  if (k_waf_z(c, first, count, dta)) return 1;
  for (int q = 0; q < count; ++q) items[q] = {c.wzall + q * fsz, kz};
  if (halo_exchange(c, items.data(), count, HS_CROSS, 2, false, true)) return 1;  // :924
  if (k_waf_y(c, first, count, dta)) return 1;
  for (int q = 0; q < count; ++q) items[q] = {c.p0all + q * fsz, kz};
  if (halo_exchange(c, items.data(), count, HS_P0, 2, true, false)) return 1;  // :955/:1012
  return k_waf_x(c, first, count, dta);
}

// uv_delivered: called from dynamical_core right after do_sound(c, true)
static int do_advection(Ctx& c, bool uv_delivered = false) {
  NvtxRange nvtx("advection");
  const int kz = c.g.kz;
  // Peer-store transport inside dynamical_core: the two exchanges around the advection proper are fused into
  // the kernels.  u, v (:1532-1533) came with the last uvupdate and destagger waits for them; ux, vx
  // (:1485-1486) are stored into the neighbours' ghost cells by destagger (all edge cells: the
  // physical-boundary rows and columns are final there) and again by curvature (the cells the advection
  // changed); restagger waits for the neighbours' word.
  const bool fused = uv_delivered && c.adv_wait_valid && halo_fused_available(c);
  PushCtl p_x0 = {}, p_x = {};
  WaitCtl w_x = {};
  EdgePush e_ux = {}, e_vx = {};
  if (fused) {
    if (halo_push_ctl(c, &p_x0)) return 1;
    if (halo_fused_edge(c, c.f[MB_UX].p, HS_CROSS, true, false, &e_ux, 2)) return 1;
    if (halo_fused_edge(c, c.f[MB_VX].p, HS_CROSS, false, true, &e_vx, 2)) return 1;
    if (k_destagger(c, &c.adv_wait, &p_x0, &e_ux, &e_vx)) return 1;
    c.adv_wait_valid = false;
  } else {   // :1532-1533, one round
    const HaloItem iu = {c.f[MB_U].p, kz}, iv = {c.f[MB_V].p, kz};
    const HaloSpec sp[2] = {{&iu, 1, HS_U, 2, true, false, 0}, {&iv, 1, HS_V, 2, false, true, 0}};
    if (halo_exchange_multi(c, sp, 2)) return 1;
    if (k_destagger(c)) return 1;
  }
  if (c.cfg.ibltyp == 2 && k_tke_destagger(c)) return 1;          // :782-784
  if (do_wafone_range(c, 0, c.nadv_fields)) return 1;             // :786-807
  if (fused) {
    // the round is numbered here, after the advection's own exchanges: restagger's signal is the latest word
    if (halo_fused_begin(c, &p_x, &w_x)) return 1;
    if (k_curvature(c, c.dtstepa, &p_x, &e_ux, &e_vx)) return 1;
    if (k_restagger(c, true, &w_x)) return 1;
  } else {
    if (k_curvature(c, c.dtstepa)) return 1;
    {   // :1485-1486, one round
      const HaloItem iu = {c.f[MB_UX].p, kz}, iv = {c.f[MB_VX].p, kz};
      const HaloSpec sp[2] = {{&iu, 1, HS_CROSS, 2, true, false, 0}, {&iv, 1, HS_CROSS, 2, false, true, 0}};
      if (halo_exchange_multi(c, sp, 2)) return 1;
    }
    if (k_restagger(c, true)) return 1;
  }
  if (c.cfg.ibltyp == 2 && k_tke_restagger(c)) return 1;          // :832-834
  return 0;
}

static int do_dynamical_core(Ctx& c) {
  NvtxRange nvtx("dynamical_core");
  if (k_diag(c, 0, false)) return 1;              // ten0, qen0, chiten0 :1092-1103
  for (int n = 0; n < c.cfg.mo_nadv; ++n) {
    if (do_sound(c, true)) return 1;
    if (do_advection(c, true)) return 1;
  }
  if (k_tvirt_temp(c)) return 1;
  return k_diag(c, 0, true);                      // tdiag%adh, qdiag%adh, cadvhdiag :1127-1139
}

static int do_reset_tendencies(Ctx& c) {
  NvtxRange nvtx("reset_tendencies");
  if (k_reset_tendencies(c)) return 1;
  if (c.cfg.ibltyp == 2) {   // tketen :1071-1075
    LaunchScope ls(c, KID_RESET);
    MB_CUDA(cudaMemsetAsync(c.f[MB_TKETEN].p, 0, c.layout.size[MB_TKETEN], c.stream));
  }
  return 0;
}

// `boundary` (Main/mod_moloch.F90:448-529)
static int do_boundary(Ctx& c) {
  NvtxRange nvtx("boundary");
  const int kz = c.g.kz;
  if (k_diag(c, 1, false)) return 1;              // :455-466
  if (k_bdyval(c, c.xbctime)) return 1;
  c.xbctime = c.xbctime + c.cfg.dtsec;            // Main/mod_bdycod.F90:2653
  if (k_bdy_relax(c, c.xbctime)) return 1;        // motopnudge + morelax: weights of the advanced xbctime
  if (c.cfg.mo_spectral_nudge) {                  // :499-506
    c.tspectral = c.tspectral + c.cfg.dtsec;
    if ((int)fmod(c.tspectral, c.cfg.dtrad) == 0 && k_spectral_nudge(c, c.xbctime)) return 1;
  }
  if (k_diag(c, 1, true)) return 1;               // tdiag%bdy, qdiag%bdy, cbdydiag :508-519
  {   // uvstagtouvx :1532-1533, one round
    const HaloItem iu = {c.f[MB_U].p, kz}, iv = {c.f[MB_V].p, kz};
    const HaloSpec sp[2] = {{&iu, 1, HS_U, 2, true, false, 0}, {&iv, 1, HS_V, 2, false, true, 0}};
    if (halo_exchange_multi(c, sp, 2)) return 1;
  }
  return k_bdy_finish(c);
}

static int require_bdy(Ctx& c) {
  if (!c.cfg.do_bdy) return fail("lateral boundary not configured (moloch_b200_config.do_bdy)");
  if (c.cfg.nspgx > 0) {
    for (int q = 0; q < 3; ++q)
      if (!c.ibnd_set[q]) return fail("boundary: ba_cr/ba_ud/ba_vd ibnd not set (moloch_b200_set_ibnd)");
    if (c.tab_n[MB_TAB_HEFC] == 0) return fail("boundary: hefc not set (moloch_b200_set_table)");
  }
  if (c.cfg.mo_top_nudge && c.tab_n[MB_TAB_TNUDGE] == 0) return fail("boundary: tnudge not set");
  if (c.cfg.ichem && c.cfg.ichebdy != 0 && c.cfg.ntr > 0 && c.cfg.nspgx > 0 && c.tab_n[MB_TAB_FCX] == 0)
    return fail("boundary: fcx not set");
  if (c.cfg.mo_spectral_nudge &&
      (c.tab_n[MB_TAB_CNUDGE] == 0 || c.tab_n[MB_TAB_BVX] == 0 || c.tab_n[MB_TAB_BVY] == 0))
    return fail("boundary: cnudge/bvx/bvy not set");
  return 0;
}

static int do_status_update(Ctx& c) {
  NvtxRange nvtx("status_update");
  const int kz = c.g.kz;
  if (c.fuse_status && c.fuse_level >= 2 && halo_fused_available(c)) {
    // the synchronisation-only round and the ux, vx round as fused rounds (see moloch_status_update); both are
    // signalled by their consumers' first CTAs (the first has no producer at all)
    PushCtl p_f = {}, p_x = {};
    WaitCtl w_f = {}, w_x = {};
    EdgePush e_ux = {}, e_vx = {};
    if (halo_fused_begin(c, &p_f, &w_f)) return 1;
    if (halo_fused_begin(c, &p_x, &w_x)) return 1;
    w_f.nosig = 0; p_x.sig = 0; w_x.nosig = 0;
    if (halo_fused_edge(c, c.f[MB_UX].p, HS_CROSS, true, false, &e_ux, 2)) return 1;
    if (halo_fused_edge(c, c.f[MB_VX].p, HS_CROSS, false, true, &e_vx, 2)) return 1;
    if (k_status_update(c, c.cfg.dtsec, &w_f, &p_x, &e_ux, &e_vx)) return 1;
    if (c.cfg.ibltyp == 2 && k_tke_update(c, c.cfg.dtsec)) return 1;   // :1419-1424
    return k_restagger(c, false, &w_x);
  }
  if (k_status_update(c, c.cfg.dtsec)) return 1;
  if (c.cfg.ibltyp == 2 && k_tke_update(c, c.cfg.dtsec)) return 1;   // :1419-1424
  if (halo_fence(c)) return 1;   // ux/vx ghosts of the previous round may still be read by a neighbour
  {
    const HaloItem iu = {c.f[MB_UX].p, kz}, iv = {c.f[MB_VX].p, kz};
    const HaloSpec sp[2] = {{&iu, 1, HS_CROSS, 2, true, false, 0}, {&iv, 1, HS_CROSS, 2, false, true, 0}};
    if (halo_exchange_multi(c, sp, 2)) return 1;
  }
  return k_restagger(c, false);
}

// ---- CUDA graphs -----------------------------------------------------------------
// A MOLOCH step is 60-100 short launches; on small per-GPU grids (config 2, or cordex25 on 8 GPUs) the host's
// launch path and the gaps between kernels are a visible part of it.  The call sequences that do not depend on
// host state are captured once (on their second call: the first one has warmed up every lazy allocation) and
// replayed.  What changes from step to step -- the halo round numbers -- is relative to a device-side base that
// the graph's last node advances (Ctx::seq_base).  Not captured: runs with the lateral boundary (host-side time
// bookkeeping and event-driven nudging), the NCCL transport (it may grow its buffers), profiled runs.
static void graphs_drop(Ctx& c) {
  for (auto& gs : c.graph) {
#ifndef MB_HOST_EMU
    if (gs.exec) cudaGraphExecDestroy((cudaGraphExec_t)gs.exec);
#endif
    gs = Ctx::GraphSlot{};
  }
}
static bool graph_ok(const Ctx& c) {
  if (!c.use_graph || c.profiling || c.cfg.do_bdy) return false;
  if (c.cfg.nranks > 1 && !c.p2p) return false;
  return true;
}
template <class F>
static int run_graphed(Ctx& c, int slot, F&& fn) {
  if (!graph_ok(c)) return fn();
  Ctx::GraphSlot& gs = c.graph[slot];
  if (!gs.exec && gs.calls++ < 1) return fn();
#ifdef MB_HOST_EMU
  // tests/emu: launches execute at once, so there is nothing to replay; the call runs eagerly and then advances
  // the base of the round numbers the way a replayed graph's last node does
  const unsigned long long seq0 = c.halo_seq;
  if (fn()) return 1;
  const unsigned long long rounds = c.halo_seq - seq0;
  if (rounds) { if (k_seq_bump(c, rounds)) return 1; c.seq_base += rounds; }
  return 0;
#else
  if (!gs.exec) {
    const unsigned long long seq0 = c.halo_seq;
    const long long l0 = c.launches;
    cudaGraph_t graph = nullptr;
    MB_CUDA(cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeThreadLocal));
    int rc = fn();
    const unsigned long long rounds = c.halo_seq - seq0;
    if (rc == 0 && rounds) rc = k_seq_bump(c, rounds);
    cudaError_t e = cudaStreamEndCapture(c.stream, &graph);
    gs.rounds = rounds; gs.launches = c.launches - l0;
    c.halo_seq = seq0; c.launches = l0;      // nothing has run yet
    if (rc) { if (graph) cudaGraphDestroy(graph); return 1; }
    if (e != cudaSuccess) { cudaGetLastError(); return fail(std::string("graph capture: ") + cudaGetErrorString(e)); }
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
    gs.exec = exec;
  }
  MB_CUDA(cudaGraphLaunch((cudaGraphExec_t)gs.exec, c.stream));
  c.halo_seq += gs.rounds; c.seq_base += gs.rounds; c.launches += gs.launches;
  return 0;
#endif
}

static int require_init(Ctx* c) {
  if (!c) return fail("null context");
  if (!c->initialised) return fail("moloch_b200_init has not been called");
  return 0;
}

}  // namespace mb

using namespace mb;
struct moloch_b200_ctx : public mb::Ctx {};

extern "C" {

const char* moloch_b200_last_error(void) { return g_err.c_str(); }
int moloch_b200_abi_version(void) { return MOLOCH_B200_ABI_VERSION; }
uint64_t moloch_b200_config_size(void) { return (uint64_t)sizeof(moloch_b200_config); }

int moloch_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int moloch_b200_destroy(moloch_b200_ctx* c);
int moloch_b200_create(const moloch_b200_config* cfg, moloch_b200_ctx** out) {
  if (!cfg || !out) return fail("moloch_b200_create: null argument");
  *out = nullptr;
  if (check_cfg(*cfg)) return 1;
  int ndev = moloch_b200_device_count();
  if (ndev <= 0) return fail("moloch_b200_create: no CUDA device available (there is no CPU fallback)");
  moloch_b200_ctx* c = new moloch_b200_ctx();
  c->cfg = *cfg;
  c->device = (cfg->device >= 0) ? cfg->device : (cfg->rank % ndev);
  if (cudaSetDevice(c->device) != cudaSuccess) {
    delete c;
    return fail("moloch_b200_create: cudaSetDevice failed");
  }
  c->g = geo_from_cfg(*cfg);
  const Geo& g = c->g;
  const int kz = g.kz;
  // advected fields, in the reference's order :786-807
  struct Adv { int fid; int spec; };
  std::vector<Adv> adv = {{MB_TETAV, 0}, {MB_PAI, 0}, {MB_UX, 0}, {MB_VX, 0}, {MB_WX, 0}, {MB_QX, 0}};
  if (cfg->ipptls > 0)
    for (int n = cfg->iqfrst; n <= cfg->nqx; ++n) adv.push_back({MB_QX, n - 1});
  if (cfg->ibltyp == 2) adv.push_back({MB_TKEX, 0});
  for (int n = 1; n <= cfg->ntr; ++n) adv.push_back({MB_TRAC, n - 1});
  c->nadv_fields = (int)adv.size();
  for (int id = 0; id < MB_NFIELDS; ++id) {
    field_shape(*cfg, id, c->f[id].nk, c->f[id].nspec);
    c->f[id].klo = 1; c->f[id].is2d = (c->f[id].nk == 1);
  }
  c->layout = make_layout(*cfg);
  const Layout& L = c->layout;
  const size_t total = L.total;
  cudaError_t e = cudaMalloc(&c->arena, total);
  if (e != cudaSuccess) {
    std::string m = std::string("moloch_b200_create: cudaMalloc of ") + std::to_string(total) + " bytes: " +
                    cudaGetErrorString(e);
    delete c;
    return fail(m);
  }
  c->arena_bytes = total;
  // every error exit below goes through moloch_b200_destroy (frees whatever exists by then)
  auto bail = [&](const std::string& m) { moloch_b200_destroy(c); return fail(m); };
  if ((e = cudaMemset(c->arena, 0, total)) != cudaSuccess)
    return bail(std::string("moloch_b200_create: cudaMemset of the arena: ") + cudaGetErrorString(e));
  for (int id = 0; id < MB_NFIELDS; ++id) {
    if (id == MB_WZ || id == MB_P0) continue;
    c->f[id].p = (c->f[id].nspec > 0 && L.size[id] > 0) ? (double*)(c->arena + L.off[id]) : nullptr;
  }
  c->zdiv2b = (double*)(c->arena + L.off[SL_ZB]);
  c->mx2 = (double*)(c->arena + L.off[SL_2D]); c->rmx = (double*)(c->arena + L.off[SL_2D] + L.stride2d);
  c->rmu = (double*)(c->arena + L.off[SL_2D] + 2 * L.stride2d);
  c->rmv = (double*)(c->arena + L.off[SL_2D] + 3 * L.stride2d);
  c->zru = (double*)(c->arena + L.off[SL_ZR]); c->zrd = (double*)(c->arena + L.off[SL_ZR] + L.stridezr);
  c->wzall = (double*)(c->arena + L.off[SL_WZ]); c->p0all = (double*)(c->arena + L.off[SL_P0]);
  c->f[MB_WZ].p = c->wzall; c->f[MB_P0].p = c->p0all;
  for (int q = 0; q < MB_NPROFILES; ++q) {
    c->prof[q] = (double*)(c->arena + L.off[SL_PROF] + (size_t)q * L.strideprof);
    c->prof_n[q] = 0;
  }
  c->flags = (unsigned long long*)(c->arena + L.off[SL_FLAGS]);
  const size_t o_tab = L.off[SL_TAB];
  c->d_ptrtab = (double**)(c->arena + o_tab);
  std::vector<double*> tab;
  for (auto& a : adv) tab.push_back(c->f[a.fid].p + (size_t)a.spec * kz * g.plane);
  for (int q = 0; q < 3; ++q) {
    if (!(cfg->do_bdy && cfg->nspgx > 0)) break;
    if (cudaMalloc(&c->ibnd[q], (size_t)g.plane * sizeof(int)) != cudaSuccess)
      return bail("moloch_b200_create: cudaMalloc of the ibnd planes failed");
  }
  if ((e = cudaMemcpy(c->d_ptrtab, tab.data(), tab.size() * sizeof(double*), cudaMemcpyHostToDevice)) != cudaSuccess)
    return bail(std::string("moloch_b200_create: upload of the field table: ") + cudaGetErrorString(e));
  c->h_ptrtab = tab;
  if (const char* e = getenv("MOLOCH_B200_WAF")) c->waf_impl = atoi(e) == 1 ? 1 : 2;
  if (const char* e = getenv("MOLOCH_B200_FUSE_HALO")) { c->fuse_halo = atoi(e) != 0; c->fuse_level = atoi(e) >= 2 ? 2 : 1; }
  if (const char* e = getenv("MOLOCH_B200_PSIGNAL")) c->psignal = atoi(e) != 0;
  if (const char* e = getenv("MOLOCH_B200_FUSE_WZ")) c->fuse_wz = atoi(e) != 0;
  if (const char* e = getenv("MOLOCH_B200_FUSE_STATUS")) c->fuse_status = atoi(e) != 0;
  if (const char* e = getenv("MOLOCH_B200_GRAPH")) c->use_graph = atoi(e) != 0;
  if (const char* e = getenv("MOLOCH_B200_WAF_ZEROSKIP")) c->waf_zero_skip = atoi(e) != 0;
  if (const char* e = getenv("MOLOCH_B200_HALO_TIMEOUT_MS")) { if (atoll(e) >= 1) c->halo_timeout_cycles = atoll(e) * 2000000LL; }
  if (const char* e = getenv("MOLOCH_B200_WSOLVE")) { const int v = atoi(e); c->wsolve_impl = (v == 2 || (v >= 5 && v <= 13)) ? v : 12; }
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    c->stream = nullptr;
    return bail("moloch_b200_create: cudaStreamCreate failed");
  }
  c->own_stream = true;
  c->rdx = 1.0 / cfg->dx;              // Main/mod_params.F90:2193-2200
  c->rdzita = 1.0 / cfg->mo_dzita;     // :275
  c->dtstepa = cfg->dtsec / (double)cfg->mo_nadv;      // :306
  c->dtsound = c->dtstepa / (double)cfg->mo_nsound;    // :307
  *out = c;
  return 0;
}

int moloch_b200_destroy(moloch_b200_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (auto& ev : c->events) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
  graphs_drop(*c);
  halo_free(*c);
  for (int q = 0; q < 3; ++q) if (c->ibnd[q]) cudaFree(c->ibnd[q]);
  for (int q = 0; q < MB_NTABLES; ++q) if (c->tab[q]) cudaFree(c->tab[q]);
  if (c->spec_work) cudaFree(c->spec_work);
  if (c->mass_work) cudaFree(c->mass_work);
  if (c->stage) cudaFree(c->stage);
  if (c->stage_down) cudaFree(c->stage_down);
  if (c->stage_up) cudaFree(c->stage_up);
  if (c->xs_down) cudaStreamDestroy(c->xs_down);
  if (c->xs_up) cudaStreamDestroy(c->xs_up);
  if (c->xs_pack) cudaStreamDestroy(c->xs_pack);
  if (c->xs_unpack) cudaStreamDestroy(c->xs_unpack);
  if (c->ev_ready) cudaEventDestroy(c->ev_ready);
  for (cudaEvent_t e : c->ev_slab) cudaEventDestroy(e);
  if (c->arena) cudaFree(c->arena);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

int moloch_b200_comm_id(void* id128) {
  if (!id128) return fail("moloch_b200_comm_id: null argument");
  return halo_comm_id(id128);
}
int moloch_b200_comm_init(moloch_b200_ctx* c, const void* id128) {
  if (!c || !id128) return fail("moloch_b200_comm_init: null argument");
  return halo_comm_init(*c, id128);
}

uint64_t moloch_b200_p2p_blob_size(void) { return (uint64_t)halo_p2p_blob_size(); }
int moloch_b200_p2p_export(moloch_b200_ctx* c, void* blob) {
  if (!c || !blob) return fail("p2p_export: null argument");
  return halo_p2p_export(*c, blob);
}
int moloch_b200_p2p_connect(moloch_b200_ctx* c, const void* blobs, int nranks) {
  if (!c || !blobs) return fail("p2p_connect: null argument");
  return halo_p2p_connect(*c, blobs, nranks);
}

int moloch_b200_set_option(moloch_b200_ctx* c, const char* name, int value) {
  if (!c || !name) return fail("set_option: null argument");
  const std::string n(name);
  if (c->stream) cudaStreamSynchronize(c->stream);
  graphs_drop(*c);      // a captured graph froze the variants it was captured with
  if (n == "wsolve") {
    if (value != 2 && (value < 5 || value > 13)) return fail("set_option: wsolve must be 2 or 5..13");
    c->wsolve_impl = value;
  } else if (n == "waf") {
    if (value != 1 && value != 2) return fail("set_option: waf must be 1 or 2");
    c->waf_impl = value;
  } else if (n == "fuse_halo") {
    if (value < 0 || value > 2) return fail("set_option: fuse_halo must be 0, 1 or 2");
    c->fuse_halo = value != 0;
    c->fuse_level = value >= 2 ? 2 : 1;
    c->adv_wait_valid = false;
  } else if (n == "halo_psignal") {     // all ranks must agree (like fuse_halo)
    c->psignal = value != 0;
    c->adv_wait_valid = false;
  } else if (n == "fuse_wz") {
    c->fuse_wz = value != 0;
  } else if (n == "fuse_status") {
    c->fuse_status = value != 0;
  } else if (n == "waf_zero_skip") {
    c->waf_zero_skip = value != 0;
  } else if (n == "graph") {
    c->use_graph = value != 0;
  } else if (n == "halo_timeout_ms") {
    if (value < 1) return fail("set_option: halo_timeout_ms must be >= 1");
    c->halo_timeout_cycles = (long long)value * 2000000LL;   // clock64 ticks at ~2 GHz
  } else {
    return fail("set_option: unknown option '" + n + "'");
  }
  return 0;
}

int moloch_b200_set_stream(moloch_b200_ctx* c, void* s) {
  if (!c) return fail("null context");
  cudaStreamSynchronize(c->stream);
  graphs_drop(*c);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  c->stream = (cudaStream_t)s;
  c->own_stream = false;
  return 0;
}
int moloch_b200_set_async(moloch_b200_ctx* c, int on) {
  if (!c) return fail("null context");
  c->async_xfer = on != 0;
  if (!on) return sync_stream(*c);
  return 0;
}
int moloch_b200_sync(moloch_b200_ctx* c) {
  if (!c) return fail("null context");
  return sync_stream(*c);
}

static int xfer_field(moloch_b200_ctx* c, int field, int n, double* host, int jlo, int jhi, int ilo, int ihi,
                      int klo, int khi, bool to_device) {
  if (!c) return fail("null context");
  if (field < 0 || field >= MB_NFIELDS) return fail("set/get_field: unknown field id");
  if (!host) return fail("set/get_field: null host pointer");
  Ctx::Field& f = c->f[field];
  if (!f.p) return fail("set/get_field: field not allocated (ntr == 0, or ibltyp/do_bdy/do_slice not configured)");
  int spec = 0;
  if (f.nspec > 1 || field == MB_QX || field == MB_TRAC || field == MB_QXTEN || field == MB_CHITEN ||
      field == MB_CHIB0 || field == MB_CHIB1 || field == MB_CHITEN0 || field == MB_CADVHDIAG || field == MB_CBDYDIAG) {
    if (n < 1 || n > f.nspec) return fail("set/get_field: species index out of range");
    spec = n - 1;
  }
  if (jhi < jlo || ihi < ilo || khi < klo) return fail("set/get_field: empty bounds");
  const Geo& g = c->g;
  const int ja = jlo > g.j0 ? jlo : g.j0, jb = jhi < g.j0 + g.NJ - 1 ? jhi : g.j0 + g.NJ - 1;
  const int ia = ilo > g.i0 ? ilo : g.i0, ib = ihi < g.i0 + g.NI - 1 ? ihi : g.i0 + g.NI - 1;
  const int ka = klo > 1 ? klo : 1, kb = khi < f.nk ? khi : f.nk;
  if (jb < ja || ib < ia || kb < ka) return 0;  // no overlap: nothing to move
  MB_CUDA(cudaSetDevice(c->device));
  double* dev = f.p + (size_t)spec * f.nk * g.plane;
  // The host box is contiguous, the device box is padded: gather/scatter on the
  // device through a contiguous staging buffer so that the PCIe transfer is one
  // linear copy (a strided cudaMemcpy3D reaches ~60 % of the link rate).
  const int nj = jb - ja + 1, ni = ib - ia + 1, nk = kb - ka + 1;
  const bool whole = (ja == jlo && jb == jhi && ia == ilo && ib == ihi && ka == klo && kb == khi);
  if (whole) {
    const size_t n_el = (size_t)nj * ni * nk;
    if (n_el > c->stage_doubles) {
      if (sync_stream(*c)) return 1;
      if (c->stage) cudaFree(c->stage);
      c->stage_doubles = n_el + n_el / 8;
      MB_CUDA(cudaMalloc(&c->stage, c->stage_doubles * sizeof(double)));
    }
    if (to_device) {
      MB_CUDA(cudaMemcpyAsync(c->stage, host, n_el * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      if (k_box_copy(*c, dev, c->stage, ja, ia, ka, nj, ni, nk, false)) return 1;
    } else {
      if (k_box_copy(*c, dev, c->stage, ja, ia, ka, nj, ni, nk, true)) return 1;
      MB_CUDA(cudaMemcpyAsync(host, c->stage, n_el * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
  } else {
    cudaMemcpy3DParms p;
    memset(&p, 0, sizeof(p));
    cudaPitchedPtr hp = make_cudaPitchedPtr((void*)host, (size_t)(jhi - jlo + 1) * sizeof(double),
                                            (size_t)(jhi - jlo + 1) * sizeof(double), (size_t)(ihi - ilo + 1));
    cudaPitchedPtr dp = make_cudaPitchedPtr((void*)dev, (size_t)g.NJ * sizeof(double),
                                            (size_t)g.NJ * sizeof(double), (size_t)g.NI);
    cudaPos hpos = make_cudaPos((size_t)(ja - jlo) * sizeof(double), (size_t)(ia - ilo), (size_t)(ka - klo));
    cudaPos dpos = make_cudaPos((size_t)(ja - g.j0) * sizeof(double), (size_t)(ia - g.i0), (size_t)(ka - 1));
    p.extent = make_cudaExtent((size_t)nj * sizeof(double), (size_t)ni, (size_t)nk);
    if (to_device) { p.srcPtr = hp; p.srcPos = hpos; p.dstPtr = dp; p.dstPos = dpos; p.kind = cudaMemcpyHostToDevice; }
    else { p.srcPtr = dp; p.srcPos = dpos; p.dstPtr = hp; p.dstPos = hpos; p.kind = cudaMemcpyDeviceToHost; }
    MB_CUDA(cudaMemcpy3DAsync(&p, c->stream));
  }
  if (!c->async_xfer) return sync_stream(*c);
  return 0;
}

int moloch_b200_set_field(moloch_b200_ctx* c, int field, int n, const double* host, int jlo, int jhi, int ilo,
                          int ihi, int klo, int khi) {
  return xfer_field(c, field, n, const_cast<double*>(host), jlo, jhi, ilo, ihi, klo, khi, true);
}
int moloch_b200_get_field(moloch_b200_ctx* c, int field, int n, double* host, int jlo, int jhi, int ilo,
                          int ihi, int klo, int khi) {
  return xfer_field(c, field, n, host, jlo, jhi, ilo, ihi, klo, khi, false);
}

// ---- pipelined physics hand-off ------------------------------------------------
namespace {
struct XferPlan {
  double* dev; double* host;
  int jlo, jhi, ilo, ihi, klo;     // host bounds
  int ja, jb, ia, ib, ka, kb;      // part that exists on the device
  int nkf, khi;                    // levels of the field on the device; last host level
};
int plan_xfer(moloch_b200_ctx* c, const moloch_b200_xfer& x, XferPlan& p, bool& empty) {
  if (x.field < 0 || x.field >= MB_NFIELDS) return fail("handoff: unknown field id");
  if (!x.host) return fail("handoff: null host pointer");
  Ctx::Field& f = c->f[x.field];
  if (!f.p) return fail("handoff: field not allocated in this configuration");
  int spec = 0;
  if (f.nspec > 1 || x.field == MB_QX || x.field == MB_TRAC || x.field == MB_QXTEN || x.field == MB_CHITEN) {
    if (x.n < 1 || x.n > f.nspec) return fail("handoff: species index out of range");
    spec = x.n - 1;
  }
  if (x.jhi < x.jlo || x.ihi < x.ilo || x.khi < x.klo) return fail("handoff: empty bounds");
  const Geo& g = c->g;
  p.dev = f.p + (size_t)spec * f.nk * g.plane;
  p.host = x.host;
  p.jlo = x.jlo; p.jhi = x.jhi; p.ilo = x.ilo; p.ihi = x.ihi; p.klo = x.klo;
  p.ja = x.jlo > g.j0 ? x.jlo : g.j0; p.jb = x.jhi < g.j0 + g.NJ - 1 ? x.jhi : g.j0 + g.NJ - 1;
  p.ia = x.ilo > g.i0 ? x.ilo : g.i0; p.ib = x.ihi < g.i0 + g.NI - 1 ? x.ihi : g.i0 + g.NI - 1;
  p.ka = x.klo > 1 ? x.klo : 1; p.kb = x.khi < f.nk ? x.khi : f.nk;
  p.nkf = f.nk; p.khi = x.khi;
  empty = (p.jb < p.ja || p.ib < p.ia || p.kb < p.ka);
  return 0;
}
// rows [i1, i2] of one array between its padded device box and its host array:
// gather/scatter kernel <-> contiguous staging buffer (k, i, j) <-> one 2-D copy
// whose "rows" are the nj*ni doubles a k-plane of the slab occupies on the host
// (host rows wider than the device box: a 3-D copy of the row pieces instead).
int slab_copy(moloch_b200_ctx* c, const XferPlan& p, int i1, int i2, bool to_device, cudaStream_t st, double* stage) {
  const int ia = p.ia > i1 ? p.ia : i1, ib = p.ib < i2 ? p.ib : i2;
  if (ib < ia) return 0;
  const int nj = p.jb - p.ja + 1, ni = ib - ia + 1, nk = p.kb - p.ka + 1;
  const size_t hrow = (size_t)(p.jhi - p.jlo + 1);                       // doubles per host row
  const size_t hplane = hrow * (size_t)(p.ihi - p.ilo + 1);
  double* h0 = p.host + (size_t)(p.ka - p.klo) * hplane + (size_t)(ia - p.ilo) * hrow + (size_t)(p.ja - p.jlo);
  const bool full_rows = (p.ja == p.jlo && p.jb == p.jhi);
  if (full_rows) {   // a k-plane of the slab is one contiguous run of nj*ni doubles on the host
    const size_t run = (size_t)nj * ni * sizeof(double);
    if (to_device) {
      MB_CUDA(cudaMemcpy2DAsync(stage, run, h0, hplane * sizeof(double), run, (size_t)nk, cudaMemcpyHostToDevice, st));
      if (k_box_copy(*c, p.dev, stage, p.ja, ia, p.ka, nj, ni, nk, false, st)) return 1;
    } else {
      if (k_box_copy(*c, p.dev, stage, p.ja, ia, p.ka, nj, ni, nk, true, st)) return 1;
      MB_CUDA(cudaMemcpy2DAsync(h0, hplane * sizeof(double), stage, run, run, (size_t)nk, cudaMemcpyDeviceToHost, st));
    }
    return 0;
  }
  cudaMemcpy3DParms q;
  memset(&q, 0, sizeof(q));
  cudaPitchedPtr hp = make_cudaPitchedPtr((void*)p.host, hrow * sizeof(double), hrow * sizeof(double),
                                          (size_t)(p.ihi - p.ilo + 1));
  cudaPitchedPtr sp = make_cudaPitchedPtr((void*)stage, (size_t)nj * sizeof(double), (size_t)nj * sizeof(double),
                                          (size_t)ni);
  cudaPos hpos = make_cudaPos((size_t)(p.ja - p.jlo) * sizeof(double), (size_t)(ia - p.ilo), (size_t)(p.ka - p.klo));
  q.extent = make_cudaExtent((size_t)nj * sizeof(double), (size_t)ni, (size_t)nk);
  if (to_device) {
    q.srcPtr = hp; q.srcPos = hpos; q.dstPtr = sp; q.kind = cudaMemcpyHostToDevice;
    MB_CUDA(cudaMemcpy3DAsync(&q, st));
    if (k_box_copy(*c, p.dev, stage, p.ja, ia, p.ka, nj, ni, nk, false, st)) return 1;
  } else {
    if (k_box_copy(*c, p.dev, stage, p.ja, ia, p.ka, nj, ni, nk, true, st)) return 1;
    q.srcPtr = sp; q.dstPtr = hp; q.dstPos = hpos; q.kind = cudaMemcpyDeviceToHost;
    MB_CUDA(cudaMemcpy3DAsync(&q, st));
  }
  return 0;
}
}  // namespace

int moloch_b200_handoff(moloch_b200_ctx* c, const moloch_b200_xfer* down, int ndown, const moloch_b200_xfer* up,
                        int nup, int nslabs, moloch_b200_physics_fn physics, void* user) {
  if (!c) return fail("null context");
  if (ndown < 0 || nup < 0 || (ndown > 0 && !down) || (nup > 0 && !up)) return fail("handoff: bad array lists");
  if (nslabs < 1) nslabs = 1;
  MB_CUDA(cudaSetDevice(c->device));
  std::vector<XferPlan> pd, pu;
  int I1 = 1 << 30, I2 = -(1 << 30);
  auto add = [&](const moloch_b200_xfer* xs, int n, std::vector<XferPlan>& out) -> int {
    for (int q = 0; q < n; ++q) {
      XferPlan p; bool empty = false;
      if (plan_xfer(c, xs[q], p, empty)) return 1;
      if (empty) continue;
      out.push_back(p);
      if (p.ia < I1) I1 = p.ia;
      if (p.ib > I2) I2 = p.ib;
    }
    return 0;
  };
  if (add(down, ndown, pd) || add(up, nup, pu)) return 1;
  if (pd.empty() && pu.empty()) return 0;
  // Consecutive species of a 4-D array (qx, trac, qxten, chiten: species-major on the host and on the device)
  // with identical bounds become ONE array with nspec*nk planes: its slab is one copy of nspec times the size
  // (the copy engines lose a fixed ~10 us per copy, scratch/pcie5.cu).
  auto merge = [&](std::vector<XferPlan>& v) {
    std::vector<XferPlan> out;
    for (const XferPlan& p : v) {
      if (!out.empty()) {
        XferPlan& a = out.back();
        const size_t hplane = (size_t)(a.jhi - a.jlo + 1) * (size_t)(a.ihi - a.ilo + 1);
        const int nka = a.kb - a.ka + 1;
        const bool whole = a.ka == a.klo && a.klo == 1 && nka % a.nkf == 0 && a.khi == a.nkf && p.khi == p.nkf &&
                           p.ka == 1 && p.klo == 1 && p.kb == p.nkf && p.nkf == a.nkf;
        if (whole && p.jlo == a.jlo && p.jhi == a.jhi && p.ilo == a.ilo && p.ihi == a.ihi && p.ja == a.ja && p.jb == a.jb &&
            p.ia == a.ia && p.ib == a.ib && p.host == a.host + (size_t)nka * hplane &&
            p.dev == a.dev + (size_t)nka * c->g.plane) {
          a.kb += p.nkf;
          continue;
        }
      }
      out.push_back(p);
    }
    v.swap(out);
  };
  merge(pd); merge(pu);
  const int rows = I2 - I1 + 1;
  if (nslabs > rows) nslabs = rows;
  const int per = (rows + nslabs - 1) / nslabs;
  nslabs = (rows + per - 1) / per;
  // An array whose host rows are exactly the device rows ("fast": the normal case) moves as ONE 2-D copy per
  // slab -- a k-plane of the slab is one contiguous run of nj*ni doubles on the host -- between the host array
  // and its packed run inside the slab's staging block; ONE kernel per slab and direction packs / unpacks all
  // arrays.  The copy engines therefore see back-to-back copies and nothing else, and the staging blocks are
  // double-buffered so that the pack kernel of slab s+1 runs while slab s travels.  Arrays whose host box is
  // wider than the device box go array by array through the 3-D copy path (slab_copy).
  auto fast = [](const XferPlan& p) { return p.ja == p.jlo && p.jb == p.jhi; };
  auto slab_tables = [&](const std::vector<XferPlan>& v, std::vector<SlabTable>& tabs, size_t& block, bool upward) {
    // tabs[s * nt + t]: t-th table (at most SLAB_MAX_ARRAYS arrays) of slab s; block: doubles of the largest slab
    std::vector<int> f;
    for (size_t q = 0; q < v.size(); ++q) if (fast(v[q])) f.push_back((int)q);
    const int nt = ((int)f.size() + SLAB_MAX_ARRAYS - 1) / SLAB_MAX_ARRAYS;
    tabs.assign((size_t)nslabs * (size_t)(nt > 0 ? nt : 0), SlabTable{});
    block = 0;
    for (int s = 0; s < nslabs; ++s) {
      const int i1 = I1 + s * per, i2 = (i1 + per - 1 < I2) ? i1 + per - 1 : I2;
      long long off = 0;
      for (size_t q = 0; q < f.size(); ++q) {
        const XferPlan& p = v[(size_t)f[q]];
        const int ia = p.ia > i1 ? p.ia : i1, ib = p.ib < i2 ? p.ib : i2;
        if (ib < ia) continue;
        SlabTable& T = tabs[(size_t)s * nt + q / SLAB_MAX_ARRAYS];
        SlabArray& A = T.a[T.n++];
        A.dev = p.dev; A.off = off; A.ja = p.ja; A.nj = p.jb - p.ja + 1; A.ia = ia; A.ni = ib - ia + 1;
        A.ka = p.ka; A.nk = p.kb - p.ka + 1; A.plan = f[q];
        A.w = A.nj * A.ni; A.a0 = 0; A.p8 = 0;
        if (upward) {
          // Host -> device reads run at full rate in both directions of the link only when they start on a
          // 128-byte boundary of host memory (measured: scratch/pcie4.cu, 47+47 GB/s against 37+39).  The planes
          // of a Fortran array start at multiples of 8 bytes: the copy of a plane therefore begins at the
          // 128-byte line that holds its first element (`lead` doubles early), and the planes whose starts are
          // congruent modulo 128 share one 2-D copy (table_copies).
          const size_t hrow = (size_t)(p.jhi - p.jlo + 1), hplane = hrow * (size_t)(p.ihi - p.ilo + 1);
          const uintptr_t h0 = (uintptr_t)(p.host + (size_t)(A.ka - p.klo) * hplane + (size_t)(ia - p.ilo) * hrow);
          int period = 1;
          while (((size_t)period * hplane * sizeof(double)) & 127) period *= 2;
          if (period == 1 || (size_t)A.nj * A.ni * sizeof(double) * (size_t)A.nk / period >= ((size_t)3 << 20)) {
            A.a0 = (int)(h0 & 127); A.p8 = (int)((hplane * sizeof(double)) & 127);
          }
          A.w = (A.nj * A.ni + 15 + 15) / 16 * 16;      // room for the largest lead, a multiple of 128 bytes
        }
        off += (long long)A.w * A.nk;
        off = (off + 15) / 16 * 16;
      }
      if ((size_t)off > block) block = (size_t)off;
    }
    return nt;
  };
  std::vector<SlabTable> td, tu;
  size_t need_d = 0, need_u = 0;
  const int ntd = slab_tables(pd, td, need_d, false), ntu = slab_tables(pu, tu, need_u, true);
  size_t need_slow = 0;     // the per-array path reuses the first staging block of its direction
  for (const auto* v : {&pd, &pu})
    for (const XferPlan& p : *v)
      if (!fast(p)) {
        const size_t n_el = (size_t)(p.jb - p.ja + 1) * (size_t)per * (size_t)(p.kb - p.ka + 1);
        if (n_el > need_slow) need_slow = n_el;
      }
  if (need_slow > need_d) need_d = need_slow;
  if (need_slow > need_u) need_u = need_slow;
  // streams, events, staging (created on first use; regrown only when idle)
  if (!c->xs_down) {
    MB_CUDA(cudaStreamCreateWithFlags(&c->xs_down, cudaStreamNonBlocking));
    MB_CUDA(cudaStreamCreateWithFlags(&c->xs_up, cudaStreamNonBlocking));
    MB_CUDA(cudaStreamCreateWithFlags(&c->xs_pack, cudaStreamNonBlocking));
    MB_CUDA(cudaStreamCreateWithFlags(&c->xs_unpack, cudaStreamNonBlocking));
    MB_CUDA(cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming));
  }
  const bool trace = getenv("MOLOCH_B200_HANDOFF_TRACE") != nullptr;   // timeline of the slabs on stderr
  if (trace && !c->ev_traced) {
    for (cudaEvent_t e : c->ev_slab) cudaEventDestroy(e);
    c->ev_slab.clear();
    cudaEventDestroy(c->ev_ready);
    MB_CUDA(cudaEventCreate(&c->ev_ready));
    c->ev_traced = true;
  }
  while ((int)c->ev_slab.size() < 4 * nslabs) {
    cudaEvent_t e;
    MB_CUDA(cudaEventCreateWithFlags(&e, trace ? cudaEventDefault : cudaEventDisableTiming));
    c->ev_slab.push_back(e);
  }
  auto grow = [&](double*& buf, size_t& have, size_t need) -> int {
    if (need <= have) return 0;
    for (cudaStream_t st : {c->xs_down, c->xs_up, c->xs_pack, c->xs_unpack}) MB_CUDA(cudaStreamSynchronize(st));
    if (buf) cudaFree(buf);
    buf = nullptr;
    have = need + need / 16;
    MB_CUDA(cudaMalloc(&buf, 2 * have * sizeof(double)));
    return 0;
  };
  if (grow(c->stage_down, c->stage_down_doubles, need_d) || grow(c->stage_up, c->stage_up_doubles, need_u)) return 1;
  cudaEvent_t* EV = c->ev_slab.data();
  auto ev_gathered = [&](int s) { return EV[4 * s]; };
  auto ev_arrived = [&](int s) { return EV[4 * s + 1]; };
  auto ev_uploaded = [&](int s) { return EV[4 * s + 2]; };
  auto ev_scattered = [&](int s) { return EV[4 * s + 3]; };
  // one 2-D copy per array of a slab table between the host array and its run in the staging block
  auto table_copies = [&](const std::vector<XferPlan>& v, const SlabTable& T, double* stage, bool to_device,
                          cudaStream_t st) -> int {
    for (int q = 0; q < T.n; ++q) {
      const SlabArray& A = T.a[q];
      const XferPlan& p = v[(size_t)A.plan];
      const size_t hrow = (size_t)(p.jhi - p.jlo + 1), hplane = hrow * (size_t)(p.ihi - p.ilo + 1);
      double* h0 = p.host + (size_t)(A.ka - p.klo) * hplane + (size_t)(A.ia - p.ilo) * hrow;
      const size_t run = (size_t)A.nj * A.ni * sizeof(double);
      if (to_device) {
        const size_t pitch = hplane * sizeof(double);
        int period = 1;
        while (((size_t)period * pitch) & 127) period *= 2;          // 1, 2, 4, 8 or 16
        if (period > 1 && run * (size_t)A.nk / period < ((size_t)3 << 20)) {
          // too few planes to stay with large copies: one unaligned copy (the kernel's lead formula needs a0 = p8 = 0,
          // which slab_tables chose by the same rule)
          MB_CUDA(cudaMemcpy2DAsync(stage + A.off, (size_t)A.w * sizeof(double), h0, pitch, run, (size_t)A.nk,
                                    cudaMemcpyHostToDevice, st));
          continue;
        }
        for (int r = 0; r < period && r < A.nk; ++r) {
          int k = r, nrow = (A.nk - r + period - 1) / period;
          uintptr_t a = (uintptr_t)h0 + (size_t)r * pitch, base = a & ~(uintptr_t)127;
          if (base < (uintptr_t)p.host) {
            // the very first plane of the array: its 128-byte line begins before the array (outside what the
            // caller page-locked): this one plane travels unaligned
            MB_CUDA(cudaMemcpyAsync(stage + A.off + (size_t)k * A.w + (a - base) / sizeof(double), (const void*)a, run,
                                    cudaMemcpyHostToDevice, st));
            k += period; --nrow; a += (size_t)period * pitch; base = a & ~(uintptr_t)127;
          }
          if (nrow > 0)
            MB_CUDA(cudaMemcpy2DAsync(stage + A.off + (size_t)k * A.w, (size_t)period * A.w * sizeof(double), (const void*)base,
                                      (size_t)period * pitch, (a - base) + run, (size_t)nrow, cudaMemcpyHostToDevice, st));
        }
      } else
        MB_CUDA(cudaMemcpy2DAsync(h0, hplane * sizeof(double), stage + A.off, run, run, (size_t)A.nk, cudaMemcpyDeviceToHost, st));
    }
    return 0;
  };
  // everything starts behind what is already enqueued on the context's stream
  MB_CUDA(cudaEventRecord(c->ev_ready, c->stream));
  for (cudaStream_t st : {c->xs_down, c->xs_up, c->xs_pack, c->xs_unpack}) MB_CUDA(cudaStreamWaitEvent(st, c->ev_ready, 0));
  // ---- the whole downward direction is enqueued at once ----
  for (int s = 0; s < nslabs; ++s) {
    const int i1 = I1 + s * per, i2 = (i1 + per - 1 < I2) ? i1 + per - 1 : I2;
    double* blk = c->stage_down + (size_t)(s & 1) * c->stage_down_doubles;
    if (s >= 2) MB_CUDA(cudaStreamWaitEvent(c->xs_pack, ev_arrived(s - 2), 0));   // the block is free again
    for (int t = 0; t < ntd; ++t)
      if (k_slab_copy(*c, td[(size_t)s * ntd + t], blk, true, c->xs_pack)) return 1;
    MB_CUDA(cudaEventRecord(ev_gathered(s), c->xs_pack));
    MB_CUDA(cudaStreamWaitEvent(c->xs_down, ev_gathered(s), 0));
    for (int t = 0; t < ntd; ++t)
      if (table_copies(pd, td[(size_t)s * ntd + t], blk, false, c->xs_down)) return 1;
    for (const XferPlan& p : pd)
      if (!fast(p) && slab_copy(c, p, i1, i2, false, c->xs_down, blk)) return 1;
    MB_CUDA(cudaEventRecord(ev_arrived(s), c->xs_down));
  }
  // ---- slab by slab: wait for the state, run the host physics, send its tendencies up ----
  int rc = 0;
  const bool nodep = getenv("MOLOCH_B200_HANDOFF_NODEP") != nullptr;   // probe only: both directions unordered
  for (int s = 0; s < nslabs && rc == 0; ++s) {
    const int i1 = I1 + s * per, i2 = (i1 + per - 1 < I2) ? i1 + per - 1 : I2;
    if (physics) {
      MB_CUDA(cudaEventSynchronize(ev_arrived(s)));   // slab s of the state is on the host
      if (physics(user, i1, i2) != 0) { rc = fail("handoff: the physics callback reported an error"); break; }
    } else if (!nodep) {
      // no host physics in between: the same ordering (tendencies of slab s after the state of slab s) is kept
      // on the device, without a host round trip per slab
      MB_CUDA(cudaStreamWaitEvent(c->xs_up, ev_arrived(s), 0));
    }
    double* blk = c->stage_up + (size_t)(s & 1) * c->stage_up_doubles;
    if (s >= 2) MB_CUDA(cudaStreamWaitEvent(c->xs_up, ev_scattered(s - 2), 0));
    for (int t = 0; t < ntu && rc == 0; ++t)
      if (table_copies(pu, tu[(size_t)s * ntu + t], blk, true, c->xs_up)) rc = 1;
    for (const XferPlan& p : pu)
      if (rc == 0 && !fast(p) && slab_copy(c, p, i1, i2, true, c->xs_up, blk)) rc = 1;
    if (rc) break;
    MB_CUDA(cudaEventRecord(ev_uploaded(s), c->xs_up));
    MB_CUDA(cudaStreamWaitEvent(c->xs_unpack, ev_uploaded(s), 0));
    for (int t = 0; t < ntu; ++t)
      if (k_slab_copy(*c, tu[(size_t)s * ntu + t], blk, false, c->xs_unpack)) { rc = 1; break; }
    MB_CUDA(cudaEventRecord(ev_scattered(s), c->xs_unpack));
  }
  // completion on return: host arrays may be reused, later work on the context's
  // stream is enqueued after this point and therefore sees the uploaded slabs
  for (cudaStream_t st : {c->xs_down, c->xs_pack, c->xs_up, c->xs_unpack}) MB_CUDA(cudaStreamSynchronize(st));
  if (trace && rc == 0) {
    fprintf(stderr, "handoff trace (ms after the start): slab  gathered  arrived  uploaded  scattered\n");
    for (int s = 0; s < nslabs; ++s) {
      float t[4] = {-1.f, -1.f, -1.f, -1.f};
      for (int q = 0; q < 4; ++q) cudaEventElapsedTime(&t[q], c->ev_ready, EV[4 * s + q]);
      fprintf(stderr, "  %2d  %8.3f %8.3f %8.3f %8.3f\n", s, t[0], t[1], t[2], t[3]);
    }
    cudaGetLastError();
  }
  if (rc == 0) rc = halo_timeout_check(*c);   // the downloaded state is host-visible now
  return rc;
}

int moloch_b200_set_profile(moloch_b200_ctx* c, int which, const double* v, int n) {
  if (!c || !v) return fail("set_profile: null argument");
  if (which < 0 || which >= MB_NPROFILES) return fail("set_profile: unknown profile id");
  const Geo& g = c->g;
  const int expect = (which == MB_GZITAK) ? g.kz + 1 : (which == MB_RLAT) ? (g.ide2 - g.ide1 + 2) : g.kz;
  if (n != expect) return fail("set_profile: wrong length");
  MB_CUDA(cudaSetDevice(c->device));
  std::vector<double> tmp(v, v + n);
  if (which == MB_RLAT) {
    // init-time constant of the ROTLLR curvature term: sin(dlat(i)) with
    // dlat = degrad*0.5*(rlat(i)+rlat(i+1)), Main/mod_moloch.F90:813-814
    for (int i = 0; i + 1 < n; ++i) tmp[i] = sin(degrad * 0.5 * (v[i] + v[i + 1]));
  }
  // stored 1-based: element k of the Fortran array at [k]
  MB_CUDA(cudaMemcpyAsync(c->prof[which] + 1, tmp.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice,
                          c->stream));
  if (sync_stream(*c)) return 1;
  c->prof_n[which] = n;
  return 0;
}

int moloch_b200_host_alloc(void** p, uint64_t bytes) {
  if (!p) return fail("host_alloc: null argument");
  MB_CUDA(cudaHostAlloc(p, (size_t)bytes, cudaHostAllocDefault));
  return 0;
}
int moloch_b200_host_free(void* p) {
  if (p) MB_CUDA(cudaFreeHost(p));
  return 0;
}

int moloch_b200_host_register(void* p, uint64_t bytes) {
  if (!p || bytes == 0) return fail("host_register: null argument");
  MB_CUDA(cudaHostRegister(p, (size_t)bytes, cudaHostRegisterDefault));
  return 0;
}
int moloch_b200_host_unregister(void* p) {
  if (p) MB_CUDA(cudaHostUnregister(p));
  return 0;
}

int moloch_b200_init(moloch_b200_ctx* c) {
  if (!c) return fail("null context");
  for (int q : {MB_GZITAK, MB_GZITAKH, MB_FFILT, MB_XKDAMP, MB_XKNU})
    if (c->prof_n[q] == 0) return fail("moloch_b200_init: vertical profiles not set (gzitak, gzitakh, ffilt, xkdamp, xknu)");
  if (c->cfg.lrotllr && c->prof_n[MB_RLAT] == 0) return fail("moloch_b200_init: rlat not set (ROTLLR)");
  MB_CUDA(cudaSetDevice(c->device));
  {
    // Widen the ghosts of the static metric fields from the reference's 1 to 2
    // points (with corners): the fused horizontal WAF kernel evaluates p0 on
    // two ghost columns.  The exchanged values equal the host-provided ones
    // wherever both exist.
    const int kz = c->g.kz;
    struct W { int fid; int stag; int nk; };
    const W ws[] = {{MB_FMZ, HS_CROSS, kz}, {MB_RFMZU, HS_U, kz}, {MB_RFMZV, HS_V, kz},
                    {MB_MSFX, HS_DOT, 1}, {MB_MSFU, HS_DOT, 1}, {MB_MSFV, HS_DOT, 1}};
    for (const W& w : ws) {
      HaloItem it = {c->f[w.fid].p, w.nk};
      if (halo_exchange(*c, &it, 1, w.stag, 2, true, false)) return 1;
      if (halo_exchange(*c, &it, 1, w.stag, 2, false, true, 2)) return 1;
    }
  }
  if (k_init_static(*c)) return 1;
  if (k_waf_ratios(*c)) return 1;
  if (sync_stream(*c)) return 1;
  c->initialised = true;
  return 0;
}

#define ENTRY(c)                                  \
  if (require_init(c)) return 1;                  \
  MB_CUDA(cudaSetDevice((c)->device));

int moloch_b200_reset_tendencies(moloch_b200_ctx* c) { ENTRY(c) return do_reset_tendencies(*c); }
int moloch_b200_sound(moloch_b200_ctx* c) { ENTRY(c) return do_sound(*c); }
int moloch_b200_advection(moloch_b200_ctx* c) { ENTRY(c) return do_advection(*c); }
int moloch_b200_wafone(moloch_b200_ctx* c, int field, int n) {
  ENTRY(c)
  if (field < 0 || field >= MB_NFIELDS || !c->f[field].p) return fail("wafone: unknown field");
  int spec = 0;
  if (field == MB_QX || field == MB_TRAC) {
    if (n < 1 || n > c->f[field].nspec) return fail("wafone: species index out of range");
    spec = n - 1;
  }
  double* want = c->f[field].p + (size_t)spec * c->g.kz * c->g.plane;
  for (int q = 0; q < c->nadv_fields; ++q)
    if (c->h_ptrtab[q] == want) return do_wafone_range(*c, q, 1);
  return fail("wafone: field is not one of the advected fields");
}
int moloch_b200_dynamical_core(moloch_b200_ctx* c) {
  ENTRY(c)
  return run_graphed(*c, Ctx::G_DYNCORE, [&] { return do_dynamical_core(*c); });
}
int moloch_b200_diagnostics(moloch_b200_ctx* c) { ENTRY(c) return k_diagnostics(*c); }
int moloch_b200_status_update(moloch_b200_ctx* c) {
  ENTRY(c)
  return run_graphed(*c, Ctx::G_STATUS, [&] { return do_status_update(*c); });
}
int moloch_b200_boundary(moloch_b200_ctx* c) {
  ENTRY(c)
  if (require_bdy(*c)) return 1;
  return do_boundary(*c);
}
int moloch_b200_bdyval(moloch_b200_ctx* c) {
  ENTRY(c)
  if (require_bdy(*c)) return 1;
  if (k_bdyval(*c, c->xbctime)) return 1;
  c->xbctime = c->xbctime + c->cfg.dtsec;
  return 0;
}
int moloch_b200_set_xbctime(moloch_b200_ctx* c, double t) {
  if (!c) return fail("null context");
  c->xbctime = t;
  return 0;
}
double moloch_b200_get_xbctime(moloch_b200_ctx* c) { return c ? c->xbctime : 0.0; }
int moloch_b200_bdy_shift(moloch_b200_ctx* c) {
  if (!c) return fail("null context");
  if (!c->cfg.do_bdy) return fail("lateral boundary not configured (moloch_b200_config.do_bdy)");
  // b0 <- b1 (Main/mod_bdycod.F90:1109-1135): the buffers swap roles, the host refills b1
  for (int id = MB_DUB0; id <= MB_CHIB0; id += 2) std::swap(c->f[id].p, c->f[id + 1].p);
  c->xbctime = 0.0;                                // :1158
  return 0;
}
int moloch_b200_set_calday(moloch_b200_ctx* c, double calday, double dayspy) {
  if (!c) return fail("null context");
  if (!(dayspy > 0.0)) return fail("set_calday: dayspy must be > 0");
  c->calday = calday; c->dayspy = dayspy;
  return 0;
}
int moloch_b200_mkslice(moloch_b200_ctx* c) {
  ENTRY(c)
  if (!c->cfg.do_slice) return fail("mkslice not configured (moloch_b200_config.do_slice)");
  return k_mkslice(*c);
}
int moloch_b200_massck(moloch_b200_ctx* c, double out[4]) {
  ENTRY(c)
  if (!out) return fail("massck: null argument");
  double o7[7];
  if (k_massck(*c, 1, o7)) return 1;
  for (int q = 0; q < 4; ++q) out[q] = o7[q];
  return 0;
}
int moloch_b200_ps_check(moloch_b200_ctx* c, double maxmin[2], int32_t* nonfinite) {
  ENTRY(c)
  if (!maxmin || !nonfinite) return fail("ps_check: null argument");
  double o7[7];
  if (k_massck(*c, 2, o7)) return 1;
  maxmin[0] = o7[4]; maxmin[1] = o7[5]; *nonfinite = (int32_t)o7[6];
  return 0;
}
int moloch_b200_step(moloch_b200_ctx* c, int nsteps) {
  ENTRY(c)
  if (c->cfg.do_bdy && require_bdy(*c)) return 1;
  auto one_step = [&]() -> int {
    NvtxRange nvtx("moloch");
    if (do_reset_tendencies(*c)) return 1;
    if (do_dynamical_core(*c)) return 1;
    if (c->cfg.do_bdy && do_boundary(*c)) return 1;          // :341-343
    if (k_diagnostics(*c)) return 1;
    if (c->cfg.do_slice) { NvtxRange nv2("mkslice"); if (k_mkslice(*c)) return 1; }   // :356-358
    return do_status_update(*c);
  };
  for (int n = 0; n < nsteps; ++n)
    if (run_graphed(*c, Ctx::G_STEP, one_step)) return 1;
  return 0;
}

int moloch_b200_set_table(moloch_b200_ctx* c, int which, const double* v, int n) {
  if (!c || !v) return fail("set_table: null argument");
  if (which < 0 || which >= MB_NTABLES) return fail("set_table: unknown table id");
  const Geo& g = c->g;
  const moloch_b200_config& f = c->cfg;
  int expect = 0;
  switch (which) {
    case MB_TAB_HEFC: expect = f.nspgx * g.kz; break;
    case MB_TAB_TNUDGE: case MB_TAB_CNUDGE: expect = g.kz; break;
    case MB_TAB_FCX: expect = f.nspgx; break;
    case MB_TAB_BVX: expect = (g.jde2 - g.jde1 + 1) * 2 * f.km; break;
    case MB_TAB_BVY: expect = (g.ide2 - g.ide1 + 1) * 2 * f.lm; break;
  }
  if (expect <= 0) return fail("set_table: this table is not used by the configuration");
  if (n != expect) return fail("set_table: wrong length");
  MB_CUDA(cudaSetDevice(c->device));
  if (!c->tab[which]) MB_CUDA(cudaMalloc(&c->tab[which], (size_t)n * sizeof(double)));
  MB_CUDA(cudaMemcpyAsync(c->tab[which], v, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (sync_stream(*c)) return 1;
  c->tab_n[which] = n;
  return 0;
}

int moloch_b200_set_ibnd(moloch_b200_ctx* c, int which, const int32_t* ibnd, int jlo, int jhi, int ilo, int ihi) {
  if (!c || !ibnd) return fail("set_ibnd: null argument");
  if (which < 0 || which > MB_IBND_VD) return fail("set_ibnd: unknown bound_area");
  if (!c->ibnd[which]) return fail("set_ibnd: lateral boundary not configured (do_bdy, nspgx > 0)");
  if (jhi < jlo || ihi < ilo) return fail("set_ibnd: empty bounds");
  MB_CUDA(cudaSetDevice(c->device));
  const size_t n = (size_t)(jhi - jlo + 1) * (size_t)(ihi - ilo + 1);
  int* tmp = nullptr;
  MB_CUDA(cudaMalloc(&tmp, n * sizeof(int)));
  cudaError_t e = cudaMemcpyAsync(tmp, ibnd, n * sizeof(int), cudaMemcpyHostToDevice, c->stream);
  int rc = 0;
  if (e != cudaSuccess) rc = fail(std::string("set_ibnd: ") + cudaGetErrorString(e));
  if (!rc) rc = k_ibnd_fill(*c, c->ibnd[which], tmp, jlo, jhi, ilo, ihi);
  if (sync_stream(*c)) rc = 1;
  cudaFree(tmp);
  if (!rc) c->ibnd_set[which] = true;
  return rc;
}

int moloch_b200_profile_enable(moloch_b200_ctx* c, int on) {
  if (!c) return fail("null context");
  if (sync_stream(*c)) return 1;
  for (auto& ev : c->events) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
  c->events.clear();
  for (int q = 0; q < KID_COUNT; ++q) { c->prof_ms[q] = 0.0; c->prof_n_launch[q] = 0; }
  c->profiling = on != 0;
  graphs_drop(*c);
  return 0;
}

int moloch_b200_profile_read(moloch_b200_ctx* c, int cap, char (*names)[48], double* total_ms,
                             int64_t* launches) {
  if (!c) { fail("null context"); return -1; }
  if (sync_stream(*c)) return -1;
  for (auto& ev : c->events) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ev.a, ev.b) == cudaSuccess) {
      c->prof_ms[ev.kid] += ms;
      c->prof_n_launch[ev.kid] += 1;
    }
    cudaEventDestroy(ev.a); cudaEventDestroy(ev.b);
  }
  c->events.clear();
  int n = 0;
  for (int q = 0; q < KID_COUNT; ++q) {
    if (c->prof_n_launch[q] == 0) continue;
    if (n < cap) {
      strncpy(names[n], kernel_name(q), 47); names[n][47] = 0;
      total_ms[n] = c->prof_ms[q];
      launches[n] = c->prof_n_launch[q];
    }
    ++n;
  }
  return n;
}

int64_t moloch_b200_launch_count(moloch_b200_ctx* c, int reset) {
  if (!c) return 0;
  const int64_t v = c->launches;
  if (reset) c->launches = 0;
  return v;
}
uint64_t moloch_b200_device_bytes(moloch_b200_ctx* c) { return c ? (uint64_t)c->arena_bytes : 0; }

int moloch_b200_halo_plan(const moloch_b200_config* cfg, int stag, int nex, int lr, int bt,
                          int32_t send_box[4][4], int32_t recv_box[4][4]) {
  if (!cfg || !send_box || !recv_box) return fail("halo_plan: null argument");
  if (stag < 0 || stag > HS_P0 || nex < 1 || nex > 2) return fail("halo_plan: bad stag/nex");
  halo_boxes(*cfg, stag, nex, lr != 0, bt != 0, send_box, recv_box);
  return 0;
}

}  // extern "C"
