/*
 * moloch_oracle.h -- C API of the CPU oracle for the MOLOCH dycore step.
 *
 * TEST INFRASTRUCTURE ONLY.  This library is a plain C++ (FP64, no FMA
 * contraction) restatement of the reference algorithm in
 * /root/reference/Main/mod_moloch.F90 (RegCM 5.0.0).  It is the checker the
 * CUDA product path is compared against; it is never linked, imported or
 * called by the product (regcm_b200/).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * PARITY PINNED ON THE REFERENCE'S OWN SOURCE.  The reference ships no unit
 * tests, golden vectors or fixtures for this path and
 * cannot be compiled in this container (no Fortran compiler, MPI or NetCDF).
 * Instead its source is EXECUTED: oracle/refrun translates the routines of the
 * MOLOCH step (moloch, sound, advection, wafone, boundary, bdyval, mkslice, ...)
 * mechanically from the .F90 files under /root/reference/Main to Python and runs them on one
 * rank; this oracle reproduces those runs bit for bit
 * (tests/test_reference_pin.py, six cases), and the digests of the reference
 * runs are committed as golden fixtures (tests/golden/reference_moloch.json)
 * that both the oracle and the CUDA path are checked against.  Not covered by
 * that pin: the set-up code (compute_moloch_static, paicompute, setup_bdycon),
 * which is an input to both sides, and runs on more than one rank, which are
 * tied to the single-rank result by decomposition invariance.  Analytic
 * known-answer tests (tests/test_oracle*.py) cover those.
 *
 * Conventions
 *  - Global arrays cross the API in C order: 2-D (iy, jx), 3-D (nk, iy, jx),
 *    4-D (n, nk, iy, jx) -- i.e. RegCM's (j,i,k,n) with j fastest, on the full
 *    dot-grid extent jx x iy (cross-grid fields leave the last row/column
 *    unused when that direction is not periodic).
 *  - The oracle internally splits the domain into px x py subdomains exactly
 *    like set_nproc (Main/mpplib/mod_mppparam.F90:1250-1641), allocates every
 *    array with the reference's bounds and ghost widths
 *    (Main/mod_atm_interface.F90:579-624, Main/mod_moloch.F90:159-199) and
 *    emulates exchange_lr/_bt/_lrbt (mod_mppparam.F90:3809-3878,4257-4309,
 *    4661-4712) by direct copies between subdomains.
 */
#ifndef MOLOCH_ORACLE_H
#define MOLOCH_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int jx, iy, kz;          /* &dimparam                                        */
  int nqx, ntr;            /* water species (ipptls=2 -> 5), tracers           */
  int i_band, i_crm;       /* &geoparam periodicity                            */
  int px, py;              /* njxcpus, niycpus (1,1 = single domain)           */
  int mo_nadv, mo_nsound;  /* &molochparam                                     */
  int mo_divdamp, mo_divfilter;
  int lrotllr;             /* iproj == 'ROTLLR'                                */
  int ipptls;              /* 0,1,2: which condensates enter tvirt             */
  int nspgx;               /* sponge width for bdywt masks (0 = no sponge)     */
  double dtsec, dx;        /* dt [s], ds*1000 [m]                              */
  double mo_ztop, mo_h, mo_a0;
} oracle_config;

/* Lateral boundary (Main/mod_moloch.F90:448-529 `boundary`), mkslice
 * (Main/mod_slice.F90:115-173) and the UW-PBL TKE path (ibltyp == 2).  Set
 * once after oracle_create and before oracle_setup_static.                  */
typedef struct {
  int do_bdy;                  /* do_apply_bdy (Main/mod_moloch.F90:305,341)       */
  int present_qc, present_qi;  /* ICBC carries qc / qi (Main/mod_bdycod.F90:695,699) */
  int mo_top_nudge, mo_spectral_nudge; /* Share/mod_dynparam.F90:207,209           */
  int ichem, ichebdy;          /* tracer boundary: 0 flux dependent, 1 chib0/chib1  */
  int ibltyp;                  /* 2: TKE is a prognostic, advected variable         */
  int icldmstrat;              /* 1: mkslice finds theta at 700 hPa                 */
  int do_slice;                /* call mkslice inside oracle_step                   */
  int idiag;                   /* > 0: tdiag%adh/bdy, qdiag%adh/bdy [F90:1092,1127,455,508] */
  int irceideal;               /* 1: mkslice keeps ptrop (Main/mod_slice.F90:345)   */
  double dtbdys, dtrad;        /* boundary / radiation periods [s]                  */
  double rhmin, rhmax, tkemin; /* Main/mod_params.F90:381-382, mod_pbl_interface:50 */
  double calday, dayspy;       /* calendar day / days per year for mkslice's ptrop  */
  int ichdiag;                 /* > 0: cadvhdiag, cbdydiag                          */
  int reserved;
} oracle_ext_config;

void*  oracle_create(const oracle_config* cfg);
int    oracle_set_ext(void* h, const oracle_ext_config* x);
void   oracle_destroy(void* h);
const char* oracle_last_error(void);

/* number of values of a named global array (0 = unknown name) */
long   oracle_global_size(void* h, const char* name);
/* scatter a global array into every subdomain (ghosts included, periodic wrap) */
int    oracle_set_global(void* h, const char* name, const double* src);
/* gather the owned cells of every subdomain into a global array              */
int    oracle_get_global(void* h, const char* name, double* dst);

/* compute_moloch_static (Main/mod_params.F90:3316-3395) + init_moloch
 * (Main/mod_moloch.F90:201-308) + ffilt (Main/mod_init.F90:1008-1026) from
 * ht,htu,htv,msfx,msfu,msfv,ulat,vlat,rlat[,hefc]                            */
int    oracle_setup_static(void* h);
/* paicompute (Main/mod_bdycod.F90:3762-3796) + Main/mod_init.F90:941-953:
 * pai,tvirt,tetav,p,rho,qs from t,qx(:,:,:,iqv),ps; w = 0                     */
int    oracle_init_state(void* h);

/* the hot path */
int    oracle_step(void* h, int nsteps);             /* moloch (physics off)  */
int    oracle_reset_tendencies(void* h);
int    oracle_sound(void* h);                        /* one sound(dtsound)    */
int    oracle_advection(void* h);                    /* one advection(dtstepa)*/
int    oracle_wafone(void* h, const char* field, int n); /* wafone(field(:,:,:,n)) */
int    oracle_dynamical_core(void* h);
int    oracle_diagnostics(void* h);                  /* p,rho,qsat,ps :348-354*/
int    oracle_status_update(void* h);
/* boundary (Main/mod_moloch.F90:448-529): bdyval MOLOCH branch
 * (Main/mod_bdycod.F90:1618-1875, :2653), chem_bdyval
 * (Main/chemlib/mod_che_bdyco.F90:391-535), motopnudge (:4049-4081),
 * morelax_external/_fraction (:3962-4031), morelax_chiten
 * (mod_che_bdyco.F90:965-1026), mospectral_nudge (:3844-3960), uvstagtouvx,
 * temp_to_tvirt, tetav.  Uses and advances the oracle's xbctime.             */
int    oracle_boundary(void* h);
int    oracle_bdyval(void* h);
int    oracle_mkslice(void* h);                     /* Main/mod_slice.F90:115-173 */
/* massck, MOLOCH branch (Main/mod_massck.F90:77-185) without the surface terms:
 * out = drymass, dryadv, qmass, qadv after the sumall over subdomains          */
int    oracle_massck(void* h, double* out4);
int    oracle_ps_check(void* h, double* maxmin, int* nonfinite);   /* Main/mod_moloch.F90:407-422 */
int    oracle_set_xbctime(void* h, double xbctime);
double oracle_get_xbctime(void* h);
int    oracle_get_int(void* h, const char* name);   /* nztop, km, lm              */

void   oracle_set_threads(int n);
int    oracle_get_threads(void);

#ifdef __cplusplus
}
#endif
#endif
