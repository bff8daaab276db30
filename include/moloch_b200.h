/*
 * moloch_b200.h -- C ABI of the B200-native MOLOCH dynamical-core step.
 *
 * This is the drop-in boundary for RegCM's `mod_moloch` (reference:
 * Main/mod_moloch.F90).  RegCM has no plugin registry: the boundary is the
 * three argument-less module procedures
 *     allocate_moloch   Main/mod_moloch.F90:159   (called Main/mod_params.F90:1980)
 *     init_moloch       Main/mod_moloch.F90:201   (called Main/mod_init.F90:923)
 *     moloch            Main/mod_moloch.F90:312   (called Main/mod_regcm_interface.F90:275)
 * plus the module state they alias (mo_atm, mddom, sfs; Main/mod_moloch.F90:204-255).
 * A Fortran ISO_C_BINDING shim keeps those three names and
 * forwards to the entry points below; everything crosses as plain pointers,
 * Fortran array bounds and default integers -- no torch, no C++ types.
 *
 * Conventions
 *  - All arrays are RegCM's (j,i,k): j fastest, then i, then k; real(rk8).
 *    A host array is described by the address of its first element and its
 *    Fortran bounds (jlo:jhi, ilo:ihi, klo:khi) in GLOBAL indices, exactly what
 *    `c_loc(a)` + `lbound/ubound` give in the shim.
 *  - Every entry returns 0 on success; on failure it returns non-zero and
 *    moloch_b200_last_error() describes it (the shim turns this into RegCM's
 *    `fatal(__FILE__,__LINE__,msg)`, Share/mod_message.F90:86-99).
 *  - One context per MPI rank / GPU, single caller thread
 *    (MPI_THREAD_FUNNELED, Main/regcm.F90:60).
 *  - There is no CPU fallback: every compute entry fails if no CUDA device is
 *    usable.
 */
#ifndef MOLOCH_B200_H
#define MOLOCH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* v2: lateral boundary, mkslice, TKE, diagnostics.  v3: moloch_b200_handoff, _host_register/_unregister,
 * _set_option; `niycpus` takes the place of the reserved last configuration word (same struct size). */
#define MOLOCH_B200_ABI_VERSION 3

/* Replaces the module globals read by allocate_moloch/init_moloch:
 * mod_dynparam index ranges (Share/mod_dynparam.F90:252-310), `ma`
 * (model_area, Main/mpplib/mod_regcm_types.F90:66-111) and the &molochparam
 * knobs (Share/mod_dynparam.F90:203-211).                                   */
typedef struct {
  int32_t jx, iy, kz;                 /* &dimparam global dot-grid extents      */
  int32_t nqx, ntr, iqfrst;           /* water species, tracers, first non-qv   */
  int32_t jde1, jde2, ide1, ide2;     /* dot   external range of this rank      */
  int32_t jce1, jce2, ice1, ice2;     /* cross external range of this rank      */
  int32_t has_bdy_left, has_bdy_right, has_bdy_bottom, has_bdy_top; /* ma%has_bdy* */
  int32_t bandflag, crmflag;          /* ma%bandflag, ma%crmflag                */
  int32_t nbr_left, nbr_right, nbr_bottom, nbr_top; /* ma%left.. (-1 = mpi_proc_null) */
  int32_t rank, nranks;
  int32_t mo_nadv, mo_nsound;
  int32_t mo_divdamp, mo_divfilter;
  int32_t lrotllr;                    /* iproj == 'ROTLLR'                      */
  int32_t ipptls;                     /* 0,1,2: condensates entering tvirt      */
  int32_t device;                     /* CUDA ordinal; <0 = rank mod ndev
                                         (Main/mod_regcm_interface.F90:397-400) */
  int32_t ibltyp;                     /* 2 (UW PBL): TKE is advected by the dycore
                                         (Main/mod_moloch.F90:782-784,799-801,832-834) */
  double dtsec, dx, mo_dzita;         /* dt [s], dx [m], zita(kz) spacing [m]   */
  /* ---- ABI v2: lateral boundary, physics hand-off ---------- */
  int32_t do_bdy;                     /* do_apply_bdy (Main/mod_moloch.F90:305,341):
                                         allocate the b0/b1 buffers, enable
                                         moloch_b200_boundary                    */
  int32_t nspgx;                      /* sponge width (rows of hefc)            */
  int32_t present_qc, present_qi;     /* ICBC carries qc/qi (mod_bdycod.F90:695,699) */
  int32_t mo_top_nudge, mo_spectral_nudge; /* Share/mod_dynparam.F90:207,209     */
  int32_t nztop;                      /* levels above zztop (mod_bdycod.F90:505-519) */
  int32_t ichem, ichebdy;             /* tracer boundary: 0 flux dependent, 1 chib0/chib1 */
  int32_t do_slice;                   /* allocate the mkslice outputs            */
  int32_t icldmstrat;                 /* 1: mkslice finds theta at 700 hPa       */
  int32_t km, lm;                     /* spectral-nudging wave numbers (lowpass_init,
                                         mod_bdycod.F90:3852-3853); 0 when unused  */
  int32_t do_massck;                  /* keep zq on the device for moloch_b200_massck      */
  double dtbdys, dtrad;               /* boundary / radiation period [s]         */
  double rhmin, rhmax, tkemin;        /* Main/mod_params.F90:381-382, mod_pbl_interface.F90:50 */
  int32_t irceideal;                  /* 1: mkslice keeps ptrop as it is (Main/mod_slice.F90:345) */
  int32_t idiag, ichdiag;             /* > 0: tendency diagnostics of dynamical_core and boundary
                                         (Main/mod_moloch.F90:1092-1103,1127-1139,455-466,508-519) */
  int32_t niycpus;                    /* cpus_per_dim(2), ranks along i (rank = locj*niycpus + loci,
                                         Main/mpplib/mod_mppparam.F90:1381-1462); needed only for
                                         mo_spectral_nudge on more than one rank (row/column reductions) */
} moloch_b200_config;

typedef struct moloch_b200_ctx moloch_b200_ctx;

/* Field identifiers for set/get (the arrays of type(atmosphere) `mo_atm`,
 * Main/mpplib/mod_regcm_types.F90:131-161, `mddom` and the work arrays of
 * allocate_moloch, Main/mod_moloch.F90:159-199).                            */
enum moloch_b200_field {
  MB_U = 0, MB_V, MB_W, MB_PAI, MB_TETAV, MB_T, MB_QX, MB_TRAC, MB_UX, MB_VX,
  MB_TVIRT, MB_P, MB_RHO, MB_QSAT, MB_PS, MB_ZETA,
  MB_FMZ, MB_FMZF, MB_RFMZU, MB_RFMZV, MB_HX, MB_HY, MB_MSFX, MB_MSFU, MB_MSFV,
  MB_CORU, MB_CORV, MB_BDYWTU, MB_BDYWTV, MB_BDYWTW,
  MB_TTEN, MB_UTEN, MB_VTEN, MB_QXTEN, MB_CHITEN,
  MB_S, MB_ZDIV2, MB_WX, MB_WZ, MB_P0, MB_TETAVF,
  /* ---- ABI v2 ---- */
  MB_TKE, MB_TKETEN, MB_TKEX,          /* ibltyp == 2: kz+1, kz+1, kz levels        */
  /* v3dbound/v2dbound b0, b1 (Main/mod_atm_interface.F90:547-577); do_bdy     */
  MB_DUB0, MB_DUB1, MB_DVB0, MB_DVB1, MB_XTB0, MB_XTB1, MB_XPAIB0, MB_XPAIB1, MB_XQB0, MB_XQB1,
  MB_XLB0, MB_XLB1, MB_XIB0, MB_XIB1, MB_XPSB0, MB_XPSB1, MB_CHIB0, MB_CHIB1,
  /* mkslice outputs (Main/mod_slice.F90:115-173) and its static input zq; do_slice */
  MB_PF3D, MB_TH3D, MB_RHB3D, MB_WPX3D, MB_RHOX2D, MB_TP2D, MB_TH700, MB_ZETAF,
  /* common tail of mkslice (:342-384): input mddom%xlat; outputs ptrop and the level indices ktrop,
   * kmxpbl (integer-valued, carried as real(rk8) across the ABI); do_slice                       */
  MB_XLAT, MB_PTROP, MB_KTROP, MB_KMXPBL,
  /* idiag > 0: ten0, qen0 and tdiag%adh, qdiag%adh, tdiag%bdy, qdiag%bdy; ichdiag > 0 (with tracers):
   * chiten0 and cadvhdiag, cbdydiag (ntr species)                                              */
  MB_TEN0, MB_QEN0, MB_TDIAG_ADH, MB_QDIAG_ADH, MB_TDIAG_BDY, MB_QDIAG_BDY, MB_CHITEN0, MB_CADVHDIAG, MB_CBDYDIAG,
  MB_NFIELDS
};

/* tables of the lateral boundary (set once after create) */
enum moloch_b200_table {
  MB_TAB_HEFC = 0,  /* hefc(nspgx,kz), n fastest   Main/mod_bdycod.F90:520-545      */
  MB_TAB_TNUDGE,    /* tnudge(kz)                  :553-560                         */
  MB_TAB_CNUDGE,    /* cnudge(kz)                  :3892-3895                       */
  MB_TAB_FCX,       /* fcx(nspgx), tracers         Main/chemlib/mod_che_bdyco.F90:101 */
  MB_TAB_BVX,       /* bvx(jde1:jde2, 2*km)        Main/mod_bdycod.F90:3882-3886    */
  MB_TAB_BVY,       /* bvy(ide1:ide2, 2*lm)        :3887-3891                       */
  MB_NTABLES
};
/* ba%ibnd of the three staggerings (setup_boundaries, Main/mod_atm_interface.F90:384-532) */
enum moloch_b200_ibnd { MB_IBND_CR = 0, MB_IBND_UD, MB_IBND_VD };

/* 1-D profiles (k = 1..n) */
enum moloch_b200_profile {
  MB_GZITAK = 0,   /* gzita(zita)   kz+1   Main/mod_moloch.F90:273 */
  MB_GZITAKH,      /* gzita(zitah)  kz     :274                    */
  MB_FFILT,        /* kz   Main/mod_init.F90:1008-1026             */
  MB_XKDAMP,       /* kz   Main/mod_moloch.F90:295                 */
  MB_XKNU,         /* kz   :297                                    */
  MB_RLAT,         /* ide1..ide2+1 (ROTLLR curvature only) :813    */
  MB_NPROFILES
};

const char* moloch_b200_last_error(void);
int moloch_b200_abi_version(void);
/* sizeof(moloch_b200_config) as the library was compiled: bindings compare it with their own layout */
uint64_t moloch_b200_config_size(void);
/* number of usable CUDA devices (0 when none; never fails) */
int moloch_b200_device_count(void);

/* = allocate_moloch (Main/mod_moloch.F90:159-199) + device copies of the
 * mo_atm arrays (Main/mod_atm_interface.F90:579-624).                       */
int moloch_b200_create(const moloch_b200_config* cfg, moloch_b200_ctx** out);
int moloch_b200_destroy(moloch_b200_ctx* ctx);

/* Halo-exchange transport (replaces exchange_lr/_bt/_lrbt,
 * Main/mpplib/mod_mppparam.F90:3809-3878,4257-4309,4661-4712).  Rank 0 calls
 * comm_id, the host broadcasts the 128 bytes (MPI_Bcast in the shim), every
 * rank calls comm_init.  Not needed when nranks == 1.                        */
int moloch_b200_comm_id(void* id128);
int moloch_b200_comm_init(moloch_b200_ctx* ctx, const void* id128);

/* Optional direct peer transport (NVLink/NVSwitch peer stores instead of NCCL
 * send/recv): every rank exports a blob, the host all-gathers the blobs in
 * rank order (MPI_Allgather in the shim) and every rank connects.  Call
 * before moloch_b200_init; all ranks of a run must make the same choice.     */
uint64_t moloch_b200_p2p_blob_size(void);
int moloch_b200_p2p_export(moloch_b200_ctx* ctx, void* blob);
int moloch_b200_p2p_connect(moloch_b200_ctx* ctx, const void* blobs, int nranks);

/* Kernel-variant switches of an existing context (what the MOLOCH_B200_* environment variables set at create):
 *   "wsolve"    12 (default) | 11 | 13   implicit-w column solver, thread per column, sweep arrays in tensor memory
 *               (tcgen05.st/ld), eight warps per SM, cp.async ring of 6 / 8 / 4 levels; kz <= 41, taller grids run 8
 *               8 | 9 | 10   row tiles, 16-byte cp.async.cg ring, sweep arrays in shared memory (4-5 warps per SM)
 *               5 | 6 | 7 | 2   round 1: 8-byte cp.async.ca ring (three / two sweep arrays), CTA-parallel variant
 *   "waf"       2 | 1       field-batched fused WAF kernels, or one kernel per reference loop nest
 *   "fuse_halo" 0 | 1 | 2   peer-store transport: exchanges fused into the kernels around them (none / the sound
 *                           loop's sub-steps 2.. / all); every rank must use the same value
 *   "fuse_wz"   1 | 0       fusion level 2 on a rows-only decomposition: exchange_bt(wz, 2) between the two WAF kernels
 *                           is stored by the vertical kernel and awaited by the horizontal one (every rank alike)
 *   "fuse_status" 1 | 0     fusion level 2: status_update's synchronisation-only round and its ux, vx round folded into
 *                           status_update / uvxtouvstag (every rank alike)
 *   "halo_psignal" 0 | 1    fused rounds signalled by the consumer's first CTA (default) or by the producer's last
 *                           edge CTA (measured slower on 8 GPUs; every rank alike)
 *   "waf_zero_skip" 1 | 0   fused WAF kernels: a field that is exactly +0 in a CTA's / warp's window is not advected
 *   "graph"     1 | 0       the step / dynamical_core / status_update sequences replayed as CUDA graphs
 *   "halo_timeout_ms"       how long a kernel of the peer-store transport waits for a neighbour's arrival
 *                           counter before it gives up (default 30 s; MOLOCH_B200_HALO_TIMEOUT_MS at create)
 * All variants give bit-identical results; they exist for measurement (bench.py times them and keeps the faster). */
int moloch_b200_set_option(moloch_b200_ctx* ctx, const char* name, int value);

/* run on a caller-owned CUDA stream (cudaStream_t) instead of the context's */
int moloch_b200_set_stream(moloch_b200_ctx* ctx, void* cuda_stream);
/* Waits for everything enqueued on the context's stream.
 * FAILURE CONTRACT of the peer-store halo transport: a kernel whose neighbour does not arrive within
 * halo_timeout_ms gives up its wait, marks the round, and every later wait gives up at once (NCCL would
 * block forever; a rank that died would hang the job).  The compute entries (_step, _sound, ...) only
 * ENQUEUE work and cannot report this.  It is reported -- non-zero return, "halo exchange timed out in
 * round N" -- by EVERY entry through which results become visible to the host: moloch_b200_sync,
 * _get_field/_set_field (unless set_async is on: then the closing _sync), _handoff, _massck, _ps_check,
 * _profile_read, _set_profile/_set_table/_set_ibnd.  Fields read after such an error are invalid; the shim
 * turns it into fatal().  The marker is sticky for the life of the context.                              */
int moloch_b200_sync(moloch_b200_ctx* ctx);

/* host -> device / device -> host of one array (species n = 1.. for the 4-D
 * qx/trac/qxten/chiten; n ignored otherwise).  The intersection of the given
 * bounds with the device box (owned range + ghosts) is transferred.  The
 * host pointer may be pageable or pinned.                                   */
int moloch_b200_set_field(moloch_b200_ctx* ctx, int field, int n, const double* host,
                          int jlo, int jhi, int ilo, int ihi, int klo, int khi);
int moloch_b200_get_field(moloch_b200_ctx* ctx, int field, int n, double* host,
                          int jlo, int jhi, int ilo, int ihi, int klo, int khi);
int moloch_b200_set_profile(moloch_b200_ctx* ctx, int profile, const double* v, int n);
int moloch_b200_set_table(moloch_b200_ctx* ctx, int table, const double* v, int n);
/* integer(ik4) ibnd(jlo:jhi, ilo:ihi) of one bound_area; values <= 0 mean
 * "not in the sponge" (the reference initialises ibnd to -1)                 */
int moloch_b200_set_ibnd(moloch_b200_ctx* ctx, int which, const int32_t* ibnd, int jlo, int jhi, int ilo,
                         int ihi);
/* on: set_field/get_field only enqueue their transfer; the caller ends a batch
 * with moloch_b200_sync (host buffers must stay valid until then).           */
int moloch_b200_set_async(moloch_b200_ctx* ctx, int on);

/* Pipelined physics hand-off (the state the column physics of
 * physical_parametrizations reads, Main/mod_moloch.F90:1143-1401, and the
 * tendencies status_update consumes, :1410-1430).  The rank's rows are cut into
 * `nslabs` slabs along i.  Slab by slab the `down` arrays travel device -> host
 * on one copy stream; as soon as a slab has arrived `physics(user, i1, i2)` is
 * called on the host (may be NULL) and that slab of the `up` arrays travels
 * host -> device on a second copy stream, so that both directions of the link
 * and the host physics overlap.  Every array is gathered/scattered on the device
 * through a contiguous staging buffer; the PCIe transfers are long linear runs.
 * On return all transfers have completed and later calls on the context see
 * the uploaded arrays.  `host` arrays should be pinned (moloch_b200_host_alloc)
 * for the two directions to overlap.  A non-zero return of `physics` aborts the
 * hand-off with an error.                                                     */
typedef struct {
  int32_t field;                        /* MB_* id                                         */
  int32_t n;                            /* species 1.. for the 4-D arrays, ignored otherwise */
  double* host;                         /* host array (jlo:jhi, ilo:ihi, klo:khi), j fastest  */
  int32_t jlo, jhi, ilo, ihi, klo, khi;
} moloch_b200_xfer;
typedef int (*moloch_b200_physics_fn)(void* user, int32_t i1, int32_t i2);
int moloch_b200_handoff(moloch_b200_ctx* ctx, const moloch_b200_xfer* down, int ndown,
                        const moloch_b200_xfer* up, int nup, int nslabs,
                        moloch_b200_physics_fn physics, void* user);

/* pinned host memory for the per-step state/tendency hand-off */
int moloch_b200_host_alloc(void** p, uint64_t bytes);
int moloch_b200_host_free(void* p);
/* page-lock an array the host already owns (RegCM's getmem pools), so that the
 * hand-off copies run asynchronously at full link rate (cudaHostRegister)       */
int moloch_b200_host_register(void* p, uint64_t bytes);
int moloch_b200_host_unregister(void* p);

/* = the device part of init_moloch (Main/mod_moloch.F90:263-308): mx2, rmx,
 * rmu, rmv and their halos, w(:,:,1)=0, clamp limits, dtstepa/dtsound.  Call
 * after the static fields (fmz..bdywt*, profiles) have been set.            */
int moloch_b200_init(moloch_b200_ctx* ctx);

/* The hot path, one entry per reference subroutine */
int moloch_b200_reset_tendencies(moloch_b200_ctx* ctx);          /* :1044 */
int moloch_b200_sound(moloch_b200_ctx* ctx);                     /* :545  sound(dtsound)     */
int moloch_b200_advection(moloch_b200_ctx* ctx);                 /* :767  advection(dtstepa) */
int moloch_b200_wafone(moloch_b200_ctx* ctx, int field, int n);  /* :838  wafone(field,dtstepa) */
int moloch_b200_dynamical_core(moloch_b200_ctx* ctx);            /* :1085 */
int moloch_b200_diagnostics(moloch_b200_ctx* ctx);               /* :348-354 p,rho,qsat,ps   */
int moloch_b200_status_update(moloch_b200_ctx* ctx);             /* :1403 (tendencies on device) */
/* `boundary` (Main/mod_moloch.F90:448-529): bdyval MOLOCH branch
 * (Main/mod_bdycod.F90:1618-1875) + chem_bdyval
 * (Main/chemlib/mod_che_bdyco.F90:391-535), motopnudge (:4049-4081),
 * morelax_external/_fraction (:3962-4031), morelax_chiten
 * (mod_che_bdyco.F90:965-1026), mospectral_nudge (:3898-3960), uvstagtouvx,
 * temp_to_tvirt, tetav.  The context keeps RegCM's xbctime: bdyval advances it
 * by dtsec (:2653), bdyin resets it (moloch_b200_bdy_shift).                  */
int moloch_b200_boundary(moloch_b200_ctx* ctx);
int moloch_b200_bdyval(moloch_b200_ctx* ctx);                    /* bdyval only            */
int moloch_b200_set_xbctime(moloch_b200_ctx* ctx, double xbctime);
double moloch_b200_get_xbctime(moloch_b200_ctx* ctx);
/* what bdyin does every dtbdys (Main/mod_bdycod.F90:1079-1423): b0 <- b1 for
 * every boundary variable (a pointer swap on the device), xbctime = 0; the
 * host then uploads the new b1 with moloch_b200_set_field.                    */
int moloch_b200_bdy_shift(moloch_b200_ctx* ctx);
/* mkslice, idynamic == 3 branch (Main/mod_slice.F90:115-173): pf3d, th3d,
 * rhb3d, wpx3d, rhox2d, tp2d, th700, the clipping of qx / trac and the common
 * tail (:342-384): ptrop from xlat and the calendar day, ktrop, kmxpbl        */
int moloch_b200_mkslice(moloch_b200_ctx* ctx);
/* calday (Main/mod_sun.F90:316) and dayspy of the run's calendar, read by mkslice's ptrop */
int moloch_b200_set_calday(moloch_b200_ctx* ctx, double calday, double dayspy);
/* The device part of massck, idynamic == 3 branch (Main/mod_massck.F90:77-175):
 * this rank's partial sums out[0..3] = tdrym, tdadv, tqmass, tqadv [kg]; the
 * host adds the surface terms (rain, evaporation), does the sumall over ranks
 * and the bookkeeping (:303-372).  Row sums run along j in the reference's
 * order, rows are added level by level: deterministic, equal to the
 * reference's single running sum to rounding (1e-12 relative for the masses).  */
int moloch_b200_massck(moloch_b200_ctx* ctx, double out[4]);
/* max/min of ps over the interior and the number of non-finite values there:
 * the CFL guard of moloch (Main/mod_moloch.F90:407-422); exact.              */
int moloch_b200_ps_check(moloch_b200_ctx* ctx, double maxmin[2], int32_t* nonfinite);
/* nsteps x [reset_tendencies, dynamical_core, boundary (do_bdy), diagnostics,
 * mkslice (do_slice), status_update]: `moloch` (Main/mod_moloch.F90:312-446)
 * with the host physics producing zero tendencies.                           */
int moloch_b200_step(moloch_b200_ctx* ctx, int nsteps);

/* built-in per-kernel device timing (CUDA events on the launching stream)   */
int moloch_b200_profile_enable(moloch_b200_ctx* ctx, int on);
/* fills up to `cap` entries; returns the number of kernel classes seen      */
int moloch_b200_profile_read(moloch_b200_ctx* ctx, int cap, char (*names)[48],
                             double* total_ms, int64_t* launches);
/* launches of this library's kernels since the last call (gpu_launches)     */
int64_t moloch_b200_launch_count(moloch_b200_ctx* ctx, int reset);
uint64_t moloch_b200_device_bytes(moloch_b200_ctx* ctx);

/* Host-only description of one halo exchange, for tests of the N>1 logic
 * without a GPU: fills send/recv boxes {j1,j2,i1,i2} for the four sides
 * (left,right,bottom,top) of an exchange of width `nex` over the owned box
 * of staggering `stag` (0 cross, 1 U, 2 V, 3 dot, 4 the wafone p0 box).
 * A side with no neighbour gets j1 > j2.                                    */
int moloch_b200_halo_plan(const moloch_b200_config* cfg, int stag, int nex, int lr, int bt,
                          int32_t send_box[4][4], int32_t recv_box[4][4]);

#ifdef __cplusplus
}
#endif
#endif
