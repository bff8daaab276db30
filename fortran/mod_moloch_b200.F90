!-----------------------------------------------------------------------------
! mod_moloch_b200.F90 -- ISO_C_BINDING shim between RegCM's mod_moloch and
! libmoloch_b200.so (include/moloch_b200.h, ABI version 3).
!
! What it replaces in the reference (Main/mod_moloch.F90):
!   allocate_moloch :159-199  -> b200_allocate      (device arena instead of the work arrays)
!   init_moloch     :201-308  -> b200_init          (static fields, tables, initial state to the device)
!   moloch          :312-446  -> b200_dycore / b200_boundary / b200_to_host / b200_from_host /
!                                b200_status_update around the UNCHANGED host physics
! INTEGRATION.md shows the dozen lines of Main/mod_moloch.F90 that call these.
!
! The file is self-contained Fortran 2008: every C entry has an interface block, the
! derived types mirror the C structs field by field, the enumerators mirror the C
! enums in order.  There is no Fortran compiler in the build image of the CUDA
! library, so tests/test_fortran_shim.py checks those three things mechanically
! against the header (argument count, kinds, VALUE attributes, field order).
!-----------------------------------------------------------------------------
module mod_moloch_b200

  use, intrinsic :: iso_c_binding
  use mod_intkinds
  use mod_realkinds
  use mod_dynparam
  use mod_runparams
  use mod_mppparam
  use mod_atm_interface
  use mod_message

  implicit none

  private

  public :: b200_allocate, b200_init, b200_finalize
  public :: b200_dycore, b200_boundary, b200_bdyin, b200_diagnostics, b200_status_update
  public :: b200_to_host, b200_from_host, b200_handoff, b200_savefile_state
  public :: b200_ps_report, b200_massck

  type, bind(C) :: moloch_b200_config
    integer(c_int32_t) :: jx
    integer(c_int32_t) :: iy
    integer(c_int32_t) :: kz
    integer(c_int32_t) :: nqx
    integer(c_int32_t) :: ntr
    integer(c_int32_t) :: iqfrst
    integer(c_int32_t) :: jde1
    integer(c_int32_t) :: jde2
    integer(c_int32_t) :: ide1
    integer(c_int32_t) :: ide2
    integer(c_int32_t) :: jce1
    integer(c_int32_t) :: jce2
    integer(c_int32_t) :: ice1
    integer(c_int32_t) :: ice2
    integer(c_int32_t) :: has_bdy_left
    integer(c_int32_t) :: has_bdy_right
    integer(c_int32_t) :: has_bdy_bottom
    integer(c_int32_t) :: has_bdy_top
    integer(c_int32_t) :: bandflag
    integer(c_int32_t) :: crmflag
    integer(c_int32_t) :: nbr_left
    integer(c_int32_t) :: nbr_right
    integer(c_int32_t) :: nbr_bottom
    integer(c_int32_t) :: nbr_top
    integer(c_int32_t) :: rank
    integer(c_int32_t) :: nranks
    integer(c_int32_t) :: mo_nadv
    integer(c_int32_t) :: mo_nsound
    integer(c_int32_t) :: mo_divdamp
    integer(c_int32_t) :: mo_divfilter
    integer(c_int32_t) :: lrotllr
    integer(c_int32_t) :: ipptls
    integer(c_int32_t) :: device
    integer(c_int32_t) :: ibltyp
    real(c_double) :: dtsec
    real(c_double) :: dx
    real(c_double) :: mo_dzita
    integer(c_int32_t) :: do_bdy
    integer(c_int32_t) :: nspgx
    integer(c_int32_t) :: present_qc
    integer(c_int32_t) :: present_qi
    integer(c_int32_t) :: mo_top_nudge
    integer(c_int32_t) :: mo_spectral_nudge
    integer(c_int32_t) :: nztop
    integer(c_int32_t) :: ichem
    integer(c_int32_t) :: ichebdy
    integer(c_int32_t) :: do_slice
    integer(c_int32_t) :: icldmstrat
    integer(c_int32_t) :: km
    integer(c_int32_t) :: lm
    integer(c_int32_t) :: do_massck
    real(c_double) :: dtbdys
    real(c_double) :: dtrad
    real(c_double) :: rhmin
    real(c_double) :: rhmax
    real(c_double) :: tkemin
    integer(c_int32_t) :: irceideal
    integer(c_int32_t) :: idiag
    integer(c_int32_t) :: ichdiag
    integer(c_int32_t) :: niycpus
  end type moloch_b200_config

  type, bind(C) :: moloch_b200_xfer
    integer(c_int32_t) :: field
    integer(c_int32_t) :: n
    type(c_ptr) :: host
    integer(c_int32_t) :: jlo
    integer(c_int32_t) :: jhi
    integer(c_int32_t) :: ilo
    integer(c_int32_t) :: ihi
    integer(c_int32_t) :: klo
    integer(c_int32_t) :: khi
  end type moloch_b200_xfer

  enum, bind(C)   ! moloch_b200_field
    enumerator :: MB_U = 0
    enumerator :: MB_V
    enumerator :: MB_W
    enumerator :: MB_PAI
    enumerator :: MB_TETAV
    enumerator :: MB_T
    enumerator :: MB_QX
    enumerator :: MB_TRAC
    enumerator :: MB_UX
    enumerator :: MB_VX
    enumerator :: MB_TVIRT
    enumerator :: MB_P
    enumerator :: MB_RHO
    enumerator :: MB_QSAT
    enumerator :: MB_PS
    enumerator :: MB_ZETA
    enumerator :: MB_FMZ
    enumerator :: MB_FMZF
    enumerator :: MB_RFMZU
    enumerator :: MB_RFMZV
    enumerator :: MB_HX
    enumerator :: MB_HY
    enumerator :: MB_MSFX
    enumerator :: MB_MSFU
    enumerator :: MB_MSFV
    enumerator :: MB_CORU
    enumerator :: MB_CORV
    enumerator :: MB_BDYWTU
    enumerator :: MB_BDYWTV
    enumerator :: MB_BDYWTW
    enumerator :: MB_TTEN
    enumerator :: MB_UTEN
    enumerator :: MB_VTEN
    enumerator :: MB_QXTEN
    enumerator :: MB_CHITEN
    enumerator :: MB_S
    enumerator :: MB_ZDIV2
    enumerator :: MB_WX
    enumerator :: MB_WZ
    enumerator :: MB_P0
    enumerator :: MB_TETAVF
    enumerator :: MB_TKE
    enumerator :: MB_TKETEN
    enumerator :: MB_TKEX
    enumerator :: MB_DUB0
    enumerator :: MB_DUB1
    enumerator :: MB_DVB0
    enumerator :: MB_DVB1
    enumerator :: MB_XTB0
    enumerator :: MB_XTB1
    enumerator :: MB_XPAIB0
    enumerator :: MB_XPAIB1
    enumerator :: MB_XQB0
    enumerator :: MB_XQB1
    enumerator :: MB_XLB0
    enumerator :: MB_XLB1
    enumerator :: MB_XIB0
    enumerator :: MB_XIB1
    enumerator :: MB_XPSB0
    enumerator :: MB_XPSB1
    enumerator :: MB_CHIB0
    enumerator :: MB_CHIB1
    enumerator :: MB_PF3D
    enumerator :: MB_TH3D
    enumerator :: MB_RHB3D
    enumerator :: MB_WPX3D
    enumerator :: MB_RHOX2D
    enumerator :: MB_TP2D
    enumerator :: MB_TH700
    enumerator :: MB_ZETAF
    enumerator :: MB_XLAT
    enumerator :: MB_PTROP
    enumerator :: MB_KTROP
    enumerator :: MB_KMXPBL
    enumerator :: MB_TEN0
    enumerator :: MB_QEN0
    enumerator :: MB_TDIAG_ADH
    enumerator :: MB_QDIAG_ADH
    enumerator :: MB_TDIAG_BDY
    enumerator :: MB_QDIAG_BDY
    enumerator :: MB_CHITEN0
    enumerator :: MB_CADVHDIAG
    enumerator :: MB_CBDYDIAG
    enumerator :: MB_NFIELDS
  end enum

  enum, bind(C)   ! moloch_b200_table
    enumerator :: MB_TAB_HEFC = 0
    enumerator :: MB_TAB_TNUDGE
    enumerator :: MB_TAB_CNUDGE
    enumerator :: MB_TAB_FCX
    enumerator :: MB_TAB_BVX
    enumerator :: MB_TAB_BVY
    enumerator :: MB_NTABLES
  end enum

  enum, bind(C)   ! moloch_b200_ibnd
    enumerator :: MB_IBND_CR = 0
    enumerator :: MB_IBND_UD
    enumerator :: MB_IBND_VD
  end enum

  enum, bind(C)   ! moloch_b200_profile
    enumerator :: MB_GZITAK = 0
    enumerator :: MB_GZITAKH
    enumerator :: MB_FFILT
    enumerator :: MB_XKDAMP
    enumerator :: MB_XKNU
    enumerator :: MB_RLAT
    enumerator :: MB_NPROFILES
  end enum

  interface
    function moloch_b200_last_error() bind(C, name='moloch_b200_last_error') result(rc)
      import
      type(c_ptr) :: rc
    end function moloch_b200_last_error
    function moloch_b200_abi_version() bind(C, name='moloch_b200_abi_version') result(rc)
      import
      integer(c_int) :: rc
    end function moloch_b200_abi_version
    function moloch_b200_config_size() bind(C, name='moloch_b200_config_size') result(rc)
      import
      integer(c_int64_t) :: rc
    end function moloch_b200_config_size
    function moloch_b200_device_count() bind(C, name='moloch_b200_device_count') result(rc)
      import
      integer(c_int) :: rc
    end function moloch_b200_device_count
    function moloch_b200_create(cfg, out) bind(C, name='moloch_b200_create') result(rc)
      import
      type(moloch_b200_config) :: cfg
      type(c_ptr), intent(out) :: out
      integer(c_int) :: rc
    end function moloch_b200_create
    function moloch_b200_destroy(ctx) bind(C, name='moloch_b200_destroy') result(rc)
      import
      type(c_ptr), value :: ctx
      integer(c_int) :: rc
    end function moloch_b200_destroy
    function moloch_b200_comm_id(id128) bind(C, name='moloch_b200_comm_id') result(rc)
      import
      type(c_ptr), value :: id128
      integer(c_int) :: rc
    end function moloch_b200_comm_id
    function moloch_b200_comm_init(ctx, id128) bind(C, name='moloch_b200_comm_init') result(rc)
      import
      type(c_ptr), value :: ctx
      type(c_ptr), value :: id128
      integer(c_int) :: rc
    end function moloch_b200_comm_init
    function moloch_b200_p2p_blob_size() bind(C, name='moloch_b200_p2p_blob_size') result(rc)
      import
      integer(c_int64_t) :: rc
    end function moloch_b200_p2p_blob_size
    function moloch_b200_p2p_export(ctx, blob) bind(C, name='moloch_b200_p2p_export') result(rc)
      import
      type(c_ptr), value :: ctx
      type(c_ptr), value :: blob
      integer(c_int) :: rc
    end function moloch_b200_p2p_export
    function moloch_b200_p2p_connect(ctx, blobs, nranks) bind(C, name='moloch_b200_p2p_connect') result(rc)
      import
      type(c_ptr), value :: ctx
      type(c_ptr), value :: blobs
      integer(c_int), value :: nranks
      integer(c_int) :: rc
    end function moloch_b200_p2p_connect
    function moloch_b200_set_option(ctx, name, value) bind(C, name='moloch_b200_set_option') result(rc)
      import
      type(c_ptr), value :: ctx
      character(kind=c_char), intent(in) :: name(*)
      integer(c_int), value :: value
      integer(c_int) :: rc
    end function moloch_b200_set_option
    function moloch_b200_set_stream(ctx, cuda_stream) bind(C, name='moloch_b200_set_stream') result(rc)
      import
      type(c_ptr), value :: ctx
      type(c_ptr), value :: cuda_stream
      integer(c_int) :: rc
    end function moloch_b200_set_stream
    function moloch_b200_sync(ctx) bind(C, name='moloch_b200_sync') result(rc)
      import
      type(c_ptr), value :: ctx
      integer(c_int) :: rc
    end function moloch_b200_sync
    function moloch_b200_set_field(ctx, field, n, host, jlo, jhi, ilo, ihi, klo, khi) &
        bind(C, name='moloch_b200_set_field') result(rc)
      import
      type(c_ptr), value :: ctx
      integer(c_int), value :: field
      integer(c_int), value :: n
      type(c_ptr), value :: host
      integer(c_int), value :: jlo
      integer(c_int), value :: jhi
      integer(c_int), value :: ilo
      integer(c_int), value :: ihi
      integer(c_int), value :: klo
      integer(c_int), value :: khi
      integer(c_int) :: rc
    end function moloch_b200_set_field
    function moloch_b200_get_field(ctx, field, n, host, jlo, jhi, ilo, ihi, klo, khi) &
        bind(C, name='moloch_b200_get_field') result(rc)
      import
      type(c_ptr), value :: ctx
      integer(c_int), value :: field
      integer(c_int), value :: n
      type(c_ptr), value :: host
      integer(c_int), value :: jlo
      integer(c_int), value :: jhi
      integer(c_int), value :: ilo
      integer(c_int), value :: ihi
      integer(c_int), value :: klo
      integer(c_int), value :: khi
      integer(c_int) :: rc
    end function moloch_b200_get_field
    function moloch_b200_set_profile(ctx, profile, v, n) bind(C, name='moloch_b200_set_profile') result(rc)
      import
      type(c_ptr), value :: ctx
      integer(c_int), value :: profile
      real(c_double) :: v(*)
      integer(c_int), value :: n
      integer(c_int) :: rc
    end function moloch_b200_set_profile
    function moloch_b200_set_table(ctx, table, v, n) bind(C, name='moloch_b200_set_table') result(rc)
      import
      type(c_ptr), value :: ctx
      integer(c_int), value :: table
      real(c_double) :: v(*)
      integer(c_int), value :: n
      integer(c_int) :: rc
    end function moloch_b200_set_table
    function moloch_b200_set_ibnd(ctx, which, ibnd, jlo, jhi, ilo, ihi) bind(C, name='moloch_b200_set_ibnd') result(rc)
      import
      type(c_ptr), value :: ctx
      integer(c_int), value :: which
      integer(c_int32_t) :: ibnd(*)
      integer(c_int), value :: jlo
      integer(c_int), value :: jhi
      integer(c_int), value :: ilo
      integer(c_int), value :: ihi
      integer(c_int) :: rc
    end function moloch_b200_set_ibnd
    function moloch_b200_set_async(ctx, on) bind(C, name='moloch_b200_set_async') result(rc)
      import
      type(c_ptr), value :: ctx
      integer(c_int), value :: on
      integer(c_int) :: rc
    end function moloch_b200_set_async
    function moloch_b200_handoff(ctx, down, ndown, up, nup, nslabs, physics, user) &
        bind(C, name='moloch_b200_handoff') result(rc)
      import
      type(c_ptr), value :: ctx
      type(moloch_b200_xfer) :: down(*)
      integer(c_int), value :: ndown
      type(moloch_b200_xfer) :: up(*)
      integer(c_int), value :: nup
      integer(c_int), value :: nslabs
      type(c_funptr), value :: physics
      type(c_ptr), value :: user
      integer(c_int) :: rc
    end function moloch_b200_handoff
    function moloch_b200_host_alloc(p, bytes) bind(C, name='moloch_b200_host_alloc') result(rc)
      import
      type(c_ptr), intent(out) :: p
      integer(c_int64_t), value :: bytes
      integer(c_int) :: rc
    end function moloch_b200_host_alloc
    function moloch_b200_host_free(p) bind(C, name='moloch_b200_host_free') result(rc)
      import
      type(c_ptr), value :: p
      integer(c_int) :: rc
    end function moloch_b200_host_free
    function moloch_b200_host_register(p, bytes) bind(C, name='moloch_b200_host_register') result(rc)
      import
      type(c_ptr), value :: p
      integer(c_int64_t), value :: bytes
      integer(c_int) :: rc
    end function moloch_b200_host_register
    function moloch_b200_host_unregister(p) bind(C, name='moloch_b200_host_unregister') result(rc)
      import
      type(c_ptr), value :: p
      integer(c_int) :: rc
    end function moloch_b200_host_unregister
    function moloch_b200_init(ctx) bind(C, name='moloch_b200_init') result(rc)
      import
      type(c_ptr), value :: ctx
      integer(c_int) :: rc
    end function moloch_b200_init
    function moloch_b200_reset_tendencies(ctx) bind(C, name='moloch_b200_reset_tendencies') result(rc)
      import
      type(c_ptr), value :: ctx
      integer(c_int) :: rc
    end function moloch_b200_reset_tendencies
    function moloch_b200_sound(ctx) bind(C, name='moloch_b200_sound') result(rc)
      import
      type(c_ptr), value :: ctx
      integer(c_int) :: rc
    end function moloch_b200_sound
    function moloch_b200_advection(ctx) bind(C, name='moloch_b200_advection') result(rc)
      import
      type(c_ptr), value :: ctx
      integer(c_int) :: rc
    end function moloch_b200_advection
    function moloch_b200_wafone(ctx, field, n) bind(C, name='moloch_b200_wafone') result(rc)
      import
      type(c_ptr), value :: ctx
      integer(c_int), value :: field
      integer(c_int), value :: n
      integer(c_int) :: rc
    end function moloch_b200_wafone
    function moloch_b200_dynamical_core(ctx) bind(C, name='moloch_b200_dynamical_core') result(rc)
      import
      type(c_ptr), value :: ctx
      integer(c_int) :: rc
    end function moloch_b200_dynamical_core
    function moloch_b200_diagnostics(ctx) bind(C, name='moloch_b200_diagnostics') result(rc)
      import
      type(c_ptr), value :: ctx
      integer(c_int) :: rc
    end function moloch_b200_diagnostics
    function moloch_b200_status_update(ctx) bind(C, name='moloch_b200_status_update') result(rc)
      import
      type(c_ptr), value :: ctx
      integer(c_int) :: rc
    end function moloch_b200_status_update
    function moloch_b200_boundary(ctx) bind(C, name='moloch_b200_boundary') result(rc)
      import
      type(c_ptr), value :: ctx
      integer(c_int) :: rc
    end function moloch_b200_boundary
    function moloch_b200_bdyval(ctx) bind(C, name='moloch_b200_bdyval') result(rc)
      import
      type(c_ptr), value :: ctx
      integer(c_int) :: rc
    end function moloch_b200_bdyval
    function moloch_b200_set_xbctime(ctx, xbctime) bind(C, name='moloch_b200_set_xbctime') result(rc)
      import
      type(c_ptr), value :: ctx
      real(c_double), value :: xbctime
      integer(c_int) :: rc
    end function moloch_b200_set_xbctime
    function moloch_b200_get_xbctime(ctx) bind(C, name='moloch_b200_get_xbctime') result(rc)
      import
      type(c_ptr), value :: ctx
      real(c_double) :: rc
    end function moloch_b200_get_xbctime
    function moloch_b200_bdy_shift(ctx) bind(C, name='moloch_b200_bdy_shift') result(rc)
      import
      type(c_ptr), value :: ctx
      integer(c_int) :: rc
    end function moloch_b200_bdy_shift
    function moloch_b200_mkslice(ctx) bind(C, name='moloch_b200_mkslice') result(rc)
      import
      type(c_ptr), value :: ctx
      integer(c_int) :: rc
    end function moloch_b200_mkslice
    function moloch_b200_set_calday(ctx, calday, dayspy) bind(C, name='moloch_b200_set_calday') result(rc)
      import
      type(c_ptr), value :: ctx
      real(c_double), value :: calday
      real(c_double), value :: dayspy
      integer(c_int) :: rc
    end function moloch_b200_set_calday
    function moloch_b200_massck(ctx, out) bind(C, name='moloch_b200_massck') result(rc)
      import
      type(c_ptr), value :: ctx
      real(c_double) :: out(*)
      integer(c_int) :: rc
    end function moloch_b200_massck
    function moloch_b200_ps_check(ctx, maxmin, nonfinite) bind(C, name='moloch_b200_ps_check') result(rc)
      import
      type(c_ptr), value :: ctx
      real(c_double) :: maxmin(*)
      integer(c_int32_t) :: nonfinite(*)
      integer(c_int) :: rc
    end function moloch_b200_ps_check
    function moloch_b200_step(ctx, nsteps) bind(C, name='moloch_b200_step') result(rc)
      import
      type(c_ptr), value :: ctx
      integer(c_int), value :: nsteps
      integer(c_int) :: rc
    end function moloch_b200_step
    function moloch_b200_profile_enable(ctx, on) bind(C, name='moloch_b200_profile_enable') result(rc)
      import
      type(c_ptr), value :: ctx
      integer(c_int), value :: on
      integer(c_int) :: rc
    end function moloch_b200_profile_enable
    function moloch_b200_profile_read(ctx, cap, names, total_ms, launches) &
        bind(C, name='moloch_b200_profile_read') result(rc)
      import
      type(c_ptr), value :: ctx
      integer(c_int), value :: cap
      character(kind=c_char), intent(in) :: names(*)
      real(c_double) :: total_ms(*)
      integer(c_int64_t) :: launches(*)
      integer(c_int) :: rc
    end function moloch_b200_profile_read
    function moloch_b200_launch_count(ctx, reset) bind(C, name='moloch_b200_launch_count') result(rc)
      import
      type(c_ptr), value :: ctx
      integer(c_int), value :: reset
      integer(c_int64_t) :: rc
    end function moloch_b200_launch_count
    function moloch_b200_device_bytes(ctx) bind(C, name='moloch_b200_device_bytes') result(rc)
      import
      type(c_ptr), value :: ctx
      integer(c_int64_t) :: rc
    end function moloch_b200_device_bytes
    function moloch_b200_halo_plan(cfg, stag, nex, lr, bt, send_box, recv_box) &
        bind(C, name='moloch_b200_halo_plan') result(rc)
      import
      type(moloch_b200_config) :: cfg
      integer(c_int), value :: stag
      integer(c_int), value :: nex
      integer(c_int), value :: lr
      integer(c_int), value :: bt
      integer(c_int32_t) :: send_box(*)
      integer(c_int32_t) :: recv_box(*)
      integer(c_int) :: rc
    end function moloch_b200_halo_plan
  end interface

  type(c_ptr) :: ctx = c_null_ptr
  logical :: use_device_bdy = .false.
  logical :: use_device_slice = .false.
  ! the hand-off lists (built once in b200_init)
  type(moloch_b200_xfer), allocatable, target :: xdown(:), xup(:)
  integer(c_int) :: ndown = 0, nup = 0

  contains

  !---------------------------------------------------------------------------
  ! error convention: non-zero -> fatal(__FILE__,__LINE__,msg)   (Share/mod_message.F90:86-99)
  !---------------------------------------------------------------------------
  subroutine chk(rc, line)
    implicit none
    integer(c_int), intent(in) :: rc
    integer, intent(in) :: line
    if ( rc /= 0 ) then
      call fatal(__FILE__, line, c_to_f(moloch_b200_last_error( )))
    end if
  end subroutine chk

  function c_to_f(p) result(s)
    implicit none
    type(c_ptr), intent(in) :: p
    character(len=:), allocatable :: s
    character(kind=c_char), pointer :: c(:)
    integer :: n
    if ( .not. c_associated(p) ) then
      s = 'moloch_b200: unknown error'
      return
    end if
    call c_f_pointer(p, c, [4096])
    n = 0
    do while ( n < 4096 )
      if ( c(n+1) == c_null_char ) exit
      n = n + 1
    end do
    allocate(character(len=n) :: s)
    s = transfer(c(1:n), s)
  end function c_to_f

  ! mpi_proc_null -> -1 (moloch_b200_config%nbr_*)
  integer(c_int32_t) function nbr(r)
    implicit none
    integer(ik4), intent(in) :: r
    nbr = int(r, c_int32_t)
    if ( r == mpi_proc_null ) nbr = -1_c_int32_t
  end function nbr

  integer(c_int32_t) function l2i(l)
    implicit none
    logical, intent(in) :: l
    l2i = 0_c_int32_t
    if ( l ) l2i = 1_c_int32_t
  end function l2i

  !---------------------------------------------------------------------------
  ! one array across the ABI: address of the first element + the Fortran bounds
  !---------------------------------------------------------------------------
  subroutine put2(f, a)
    implicit none
    integer(c_int), intent(in) :: f
    real(rkx), pointer, contiguous, intent(in) :: a(:,:)
    call chk(moloch_b200_set_field(ctx, f, 0_c_int, c_loc(a), &
             int(lbound(a,1),c_int), int(ubound(a,1),c_int), &
             int(lbound(a,2),c_int), int(ubound(a,2),c_int), 1_c_int, 1_c_int), __LINE__)
  end subroutine put2

  subroutine put3(f, a)
    implicit none
    integer(c_int), intent(in) :: f
    real(rkx), pointer, contiguous, intent(in) :: a(:,:,:)
    call chk(moloch_b200_set_field(ctx, f, 0_c_int, c_loc(a), &
             int(lbound(a,1),c_int), int(ubound(a,1),c_int), &
             int(lbound(a,2),c_int), int(ubound(a,2),c_int), &
             int(lbound(a,3),c_int), int(ubound(a,3),c_int)), __LINE__)
  end subroutine put3

  subroutine put4(f, a)            ! qx, trac, qxten, chiten: species by species
    implicit none
    integer(c_int), intent(in) :: f
    real(rkx), pointer, contiguous, intent(in) :: a(:,:,:,:)
    integer :: n
    do n = lbound(a,4), ubound(a,4)
      call chk(moloch_b200_set_field(ctx, f, int(n-lbound(a,4)+1,c_int), &
               c_loc(a(lbound(a,1),lbound(a,2),lbound(a,3),n)), &
               int(lbound(a,1),c_int), int(ubound(a,1),c_int), &
               int(lbound(a,2),c_int), int(ubound(a,2),c_int), &
               int(lbound(a,3),c_int), int(ubound(a,3),c_int)), __LINE__)
    end do
  end subroutine put4

  subroutine get2(f, a)
    implicit none
    integer(c_int), intent(in) :: f
    real(rkx), pointer, contiguous, intent(in) :: a(:,:)
    call chk(moloch_b200_get_field(ctx, f, 0_c_int, c_loc(a), &
             int(lbound(a,1),c_int), int(ubound(a,1),c_int), &
             int(lbound(a,2),c_int), int(ubound(a,2),c_int), 1_c_int, 1_c_int), __LINE__)
  end subroutine get2

  subroutine get3(f, a)
    implicit none
    integer(c_int), intent(in) :: f
    real(rkx), pointer, contiguous, intent(in) :: a(:,:,:)
    call chk(moloch_b200_get_field(ctx, f, 0_c_int, c_loc(a), &
             int(lbound(a,1),c_int), int(ubound(a,1),c_int), &
             int(lbound(a,2),c_int), int(ubound(a,2),c_int), &
             int(lbound(a,3),c_int), int(ubound(a,3),c_int)), __LINE__)
  end subroutine get3

  subroutine get4(f, a)
    implicit none
    integer(c_int), intent(in) :: f
    real(rkx), pointer, contiguous, intent(in) :: a(:,:,:,:)
    integer :: n
    do n = lbound(a,4), ubound(a,4)
      call chk(moloch_b200_get_field(ctx, f, int(n-lbound(a,4)+1,c_int), &
               c_loc(a(lbound(a,1),lbound(a,2),lbound(a,3),n)), &
               int(lbound(a,1),c_int), int(ubound(a,1),c_int), &
               int(lbound(a,2),c_int), int(ubound(a,2),c_int), &
               int(lbound(a,3),c_int), int(ubound(a,3),c_int)), __LINE__)
    end do
  end subroutine get4

  subroutine putprof(p, v)
    implicit none
    integer(c_int), intent(in) :: p
    real(rkx), intent(in) :: v(:)
    real(c_double), allocatable :: tmp(:)
    allocate(tmp(size(v)))
    tmp(:) = real(v(:), c_double)
    call chk(moloch_b200_set_profile(ctx, p, tmp, int(size(v),c_int)), __LINE__)
  end subroutine putprof

  ! one entry of a hand-off list (species n of a 4-D array, or n = 0)
  subroutine xentry(x, f, n, a)
    implicit none
    type(moloch_b200_xfer), intent(out) :: x
    integer(c_int), intent(in) :: f, n
    real(rkx), pointer, contiguous, intent(in) :: a(:,:,:)
    x%field = f
    x%n = n
    x%host = c_loc(a)
    x%jlo = int(lbound(a,1),c_int32_t) ; x%jhi = int(ubound(a,1),c_int32_t)
    x%ilo = int(lbound(a,2),c_int32_t) ; x%ihi = int(ubound(a,2),c_int32_t)
    x%klo = int(lbound(a,3),c_int32_t) ; x%khi = int(ubound(a,3),c_int32_t)
    ! RegCM owns the array: page-lock it in place so that the copies run asynchronously at link rate
    call chk(moloch_b200_host_register(c_loc(a), int(size(a),c_int64_t)*8_c_int64_t), __LINE__)
  end subroutine xentry

  !---------------------------------------------------------------------------
  ! allocate_moloch (Main/mod_moloch.F90:159-199): the configuration the reference
  ! reads from mod_dynparam / mod_runparams / ma, and the device arena
  !---------------------------------------------------------------------------
  subroutine b200_allocate(device_boundary, device_slice, nztop, km, lm, dtbdys_, dtrad_, rhmin_, rhmax_, tkemin_)
    implicit none
    logical, intent(in) :: device_boundary, device_slice
    integer(ik4), intent(in) :: nztop, km, lm
    real(rkx), intent(in) :: dtbdys_, dtrad_, rhmin_, rhmax_, tkemin_
    type(moloch_b200_config) :: cfg
    character(kind=c_char), target :: id(128)
    character(kind=c_char), allocatable, target :: blob(:), blobs(:)
    integer(c_int64_t) :: nb
    integer(ik4) :: ierr

    use_device_bdy = device_boundary
    use_device_slice = device_slice
    cfg%jx = jx ; cfg%iy = iy ; cfg%kz = kz
    cfg%nqx = nqx ; cfg%ntr = ntr ; cfg%iqfrst = iqfrst
    cfg%jde1 = jde1 ; cfg%jde2 = jde2 ; cfg%ide1 = ide1 ; cfg%ide2 = ide2
    cfg%jce1 = jce1 ; cfg%jce2 = jce2 ; cfg%ice1 = ice1 ; cfg%ice2 = ice2
    cfg%has_bdy_left = l2i(ma%has_bdyleft) ; cfg%has_bdy_right = l2i(ma%has_bdyright)
    cfg%has_bdy_bottom = l2i(ma%has_bdybottom) ; cfg%has_bdy_top = l2i(ma%has_bdytop)
    cfg%bandflag = l2i(ma%bandflag) ; cfg%crmflag = l2i(ma%crmflag)
    cfg%nbr_left = nbr(ma%left) ; cfg%nbr_right = nbr(ma%right)
    cfg%nbr_bottom = nbr(ma%bottom) ; cfg%nbr_top = nbr(ma%top)
    cfg%rank = myid ; cfg%nranks = nproc
    cfg%mo_nadv = mo_nadv ; cfg%mo_nsound = mo_nsound
    cfg%mo_divdamp = l2i(mo_divdamp) ; cfg%mo_divfilter = l2i(mo_divfilter)
    cfg%lrotllr = l2i(iproj == 'ROTLLR')
    cfg%ipptls = ipptls
    cfg%device = -1_c_int32_t                 ! rank mod ndev, Main/mod_regcm_interface.F90:397-400
    cfg%ibltyp = ibltyp
    cfg%dtsec = dtsec ; cfg%dx = dx ; cfg%mo_dzita = mo_dzita
    cfg%do_bdy = l2i(device_boundary)
    cfg%nspgx = nspgx
    cfg%present_qc = l2i(present_qc) ; cfg%present_qi = l2i(present_qi)
    cfg%mo_top_nudge = l2i(mo_top_nudge) ; cfg%mo_spectral_nudge = l2i(mo_spectral_nudge)
    cfg%nztop = nztop
    cfg%ichem = ichem ; cfg%ichebdy = ichebdy
    cfg%do_slice = l2i(device_slice)
    cfg%icldmstrat = icldmstrat
    cfg%km = km ; cfg%lm = lm
    cfg%do_massck = l2i(debug_level > 0)
    cfg%dtbdys = dtbdys_ ; cfg%dtrad = dtrad_
    cfg%rhmin = rhmin_ ; cfg%rhmax = rhmax_ ; cfg%tkemin = tkemin_
    cfg%irceideal = irceideal
    cfg%idiag = idiag ; cfg%ichdiag = ichdiag
    cfg%niycpus = niycpus                     ! cpus_per_dim(2), Main/mpplib/mod_mppparam.F90:1381-1462

    if ( moloch_b200_abi_version( ) /= 3 ) then
      call fatal(__FILE__,__LINE__,'libmoloch_b200.so: ABI version mismatch')
    end if
    if ( moloch_b200_config_size( ) /= int(c_sizeof(cfg),c_int64_t) ) then
      call fatal(__FILE__,__LINE__,'moloch_b200_config: layout mismatch with the library')
    end if
    call chk(moloch_b200_create(cfg, ctx), __LINE__)

    if ( nproc > 1 ) then
      ! NCCL communicator (row/column reductions of the spectral nudging; halo fallback)
      if ( myid == 0 ) call chk(moloch_b200_comm_id(c_loc(id)), __LINE__)
      call mpi_bcast(id, 128, mpi_character, 0, mycomm, ierr)
      call chk(moloch_b200_comm_init(ctx, c_loc(id)), __LINE__)
      ! direct NVLink peer stores for the halo exchange: all-gather the blobs in rank order
      nb = moloch_b200_p2p_blob_size( )
      allocate(blob(nb), blobs(nb*nproc))
      call chk(moloch_b200_p2p_export(ctx, c_loc(blob)), __LINE__)
      call mpi_allgather(blob, int(nb), mpi_character, blobs, int(nb), mpi_character, mycomm, ierr)
      call chk(moloch_b200_p2p_connect(ctx, c_loc(blobs), int(nproc,c_int)), __LINE__)
      deallocate(blob, blobs)
    end if
  end subroutine b200_allocate

  !---------------------------------------------------------------------------
  ! init_moloch (Main/mod_moloch.F90:201-308).  The caller (the reference's own
  ! init_moloch, whose host lines :256-302 stay as they are) passes the arrays it
  ! has just computed; the mo_atm / mddom arrays come from mod_atm_interface.
  !---------------------------------------------------------------------------
  subroutine b200_init(coru, corv, bdywtu, bdywtv, bdywtw, gzitak, gzitakh, xkdamp, xknu, hefc_, tnudge_, cnudge_, fcx_)
    implicit none
    real(rkx), pointer, contiguous, intent(in) :: coru(:,:), corv(:,:)
    real(rkx), pointer, contiguous, intent(in) :: bdywtu(:,:,:), bdywtv(:,:,:), bdywtw(:,:,:)
    real(rkx), intent(in) :: gzitak(:), gzitakh(:), xkdamp(:), xknu(:)
    real(rkx), intent(in), optional :: hefc_(:,:), tnudge_(:), cnudge_(:), fcx_(:)
    real(c_double), allocatable :: tmp(:)
    integer :: n, q

    ! static fields: compute_moloch_static (Main/mod_params.F90:3316-3395) filled them
    call put3(MB_FMZ, mo_atm%fmz) ; call put3(MB_FMZF, mo_atm%fmzf)
    call put3(MB_RFMZU, mo_atm%rfmzu) ; call put3(MB_RFMZV, mo_atm%rfmzv)
    call put3(MB_ZETA, mo_atm%zeta)
    call put2(MB_HX, mddom%hx) ; call put2(MB_HY, mddom%hy)
    call put2(MB_MSFX, mddom%msfx) ; call put2(MB_MSFU, mddom%msfu) ; call put2(MB_MSFV, mddom%msfv)
    call put2(MB_CORU, coru) ; call put2(MB_CORV, corv)
    call put3(MB_BDYWTU, bdywtu) ; call put3(MB_BDYWTV, bdywtv) ; call put3(MB_BDYWTW, bdywtw)
    call putprof(MB_GZITAK, gzitak) ; call putprof(MB_GZITAKH, gzitakh)
    call putprof(MB_FFILT, ffilt)                ! Main/mod_init.F90:1008-1026
    call putprof(MB_XKDAMP, xkdamp) ; call putprof(MB_XKNU, xknu)
    if ( iproj == 'ROTLLR' ) call putprof(MB_RLAT, mddom%rlat(ide1:ide2+1))
    if ( use_device_slice ) then
      call put3(MB_ZETAF, mo_atm%zetaf) ; call put2(MB_XLAT, mddom%xlat)
    else if ( debug_level > 0 ) then
      call put3(MB_ZETAF, mo_atm%zetaf)          ! massck
    end if

    ! lateral boundary tables (setup_bdycon, Main/mod_bdycod.F90:478-568) and the sponge index planes
    if ( use_device_bdy ) then
      if ( nspgx > 0 .and. present(hefc_) ) then
        allocate(tmp(size(hefc_)))
        tmp = reshape(real(hefc_, c_double), [size(hefc_)])
        call chk(moloch_b200_set_table(ctx, MB_TAB_HEFC, tmp, int(size(tmp),c_int)), __LINE__)
        deallocate(tmp)
        call chk(moloch_b200_set_ibnd(ctx, MB_IBND_CR, ba_cr%ibnd(jde1:jde2,ide1:ide2), jde1, jde2, ide1, ide2), __LINE__)
        call chk(moloch_b200_set_ibnd(ctx, MB_IBND_UD, ba_ud%ibnd(jde1:jde2,ide1:ide2), jde1, jde2, ide1, ide2), __LINE__)
        call chk(moloch_b200_set_ibnd(ctx, MB_IBND_VD, ba_vd%ibnd(jde1:jde2,ide1:ide2), jde1, jde2, ide1, ide2), __LINE__)
      end if
      if ( present(tnudge_) ) then
        call chk(moloch_b200_set_table(ctx, MB_TAB_TNUDGE, real(tnudge_,c_double), int(size(tnudge_),c_int)), __LINE__)
      end if
      if ( present(cnudge_) ) then
        call chk(moloch_b200_set_table(ctx, MB_TAB_CNUDGE, real(cnudge_,c_double), int(size(cnudge_),c_int)), __LINE__)
      end if
      if ( present(fcx_) ) then
        call chk(moloch_b200_set_table(ctx, MB_TAB_FCX, real(fcx_,c_double), int(size(fcx_),c_int)), __LINE__)
      end if
    end if

    call upload_state
    call chk(moloch_b200_init(ctx), __LINE__)

    ! the per-step hand-off: what physical_parametrizations and mkslice read goes down, the
    ! tendencies status_update consumes come up (Main/mod_moloch.F90:362-376, 1410-1430)
    ndown = 12 + nqx + ntr
    nup = 3 + nqx + ntr
    allocate(xdown(ndown), xup(nup))
    q = 0
    q = q + 1 ; call xentry(xdown(q), MB_U, 0, mo_atm%u)
    q = q + 1 ; call xentry(xdown(q), MB_V, 0, mo_atm%v)
    q = q + 1 ; call xentry(xdown(q), MB_W, 0, mo_atm%w)
    q = q + 1 ; call xentry(xdown(q), MB_UX, 0, mo_atm%ux)
    q = q + 1 ; call xentry(xdown(q), MB_VX, 0, mo_atm%vx)
    q = q + 1 ; call xentry(xdown(q), MB_PAI, 0, mo_atm%pai)
    q = q + 1 ; call xentry(xdown(q), MB_TETAV, 0, mo_atm%tetav)
    q = q + 1 ; call xentry(xdown(q), MB_T, 0, mo_atm%t)
    q = q + 1 ; call xentry(xdown(q), MB_TVIRT, 0, mo_atm%tvirt)
    q = q + 1 ; call xentry(xdown(q), MB_P, 0, mo_atm%p)
    q = q + 1 ; call xentry(xdown(q), MB_RHO, 0, mo_atm%rho)
    q = q + 1 ; call xentry(xdown(q), MB_QSAT, 0, mo_atm%qs)
    do n = 1, nqx
      q = q + 1 ; call xentry(xdown(q), MB_QX, n, sp3(mo_atm%qx, n))
    end do
    do n = 1, ntr
      q = q + 1 ; call xentry(xdown(q), MB_TRAC, n, sp3(mo_atm%trac, n))
    end do
    q = 0
    q = q + 1 ; call xentry(xup(q), MB_TTEN, 0, mo_atm%tten)
    q = q + 1 ; call xentry(xup(q), MB_UTEN, 0, mo_atm%uten)
    q = q + 1 ; call xentry(xup(q), MB_VTEN, 0, mo_atm%vten)
    do n = 1, nqx
      q = q + 1 ; call xentry(xup(q), MB_QXTEN, n, sp3(mo_atm%qxten, n))
    end do
    do n = 1, ntr
      q = q + 1 ; call xentry(xup(q), MB_CHITEN, n, sp3(mo_atm%chiten, n))
    end do
  end subroutine b200_init

  ! species n of a 4-D array as a 3-D pointer with the array's own bounds
  function sp3(a, n) result(p)
    implicit none
    real(rkx), pointer, contiguous, intent(in) :: a(:,:,:,:)
    integer, intent(in) :: n
    real(rkx), pointer, contiguous :: p(:,:,:)
    p(lbound(a,1):,lbound(a,2):,lbound(a,3):) => a(:,:,:,lbound(a,4)+n-1)
  end function sp3

  ! the prognostic and derived state of mo_atm, as `init` left it (Main/mod_init.F90:156-221, 941-953)
  subroutine upload_state
    implicit none
    call put3(MB_U, mo_atm%u) ; call put3(MB_V, mo_atm%v) ; call put3(MB_W, mo_atm%w)
    call put3(MB_UX, mo_atm%ux) ; call put3(MB_VX, mo_atm%vx)
    call put3(MB_PAI, mo_atm%pai) ; call put3(MB_TETAV, mo_atm%tetav)
    call put3(MB_T, mo_atm%t) ; call put3(MB_TVIRT, mo_atm%tvirt)
    call put3(MB_P, mo_atm%p) ; call put3(MB_RHO, mo_atm%rho) ; call put3(MB_QSAT, mo_atm%qs)
    call put2(MB_PS, sfs%psa)
    call put4(MB_QX, mo_atm%qx)
    if ( ichem == 1 ) call put4(MB_TRAC, mo_atm%trac)
    if ( ibltyp == 2 ) call put3(MB_TKE, mo_atm%tke)
  end subroutine upload_state

  ! everything the host reads after the dycore, one array after the other (start-up, output steps)
  subroutine download_state
    implicit none
    call chk(moloch_b200_set_async(ctx, 1_c_int), __LINE__)
    call get3(MB_U, mo_atm%u) ; call get3(MB_V, mo_atm%v) ; call get3(MB_W, mo_atm%w)
    call get3(MB_UX, mo_atm%ux) ; call get3(MB_VX, mo_atm%vx)
    call get3(MB_PAI, mo_atm%pai) ; call get3(MB_TETAV, mo_atm%tetav)
    call get3(MB_T, mo_atm%t) ; call get3(MB_TVIRT, mo_atm%tvirt)
    call get3(MB_P, mo_atm%p) ; call get3(MB_RHO, mo_atm%rho) ; call get3(MB_QSAT, mo_atm%qs)
    call get2(MB_PS, sfs%psa)
    call get4(MB_QX, mo_atm%qx)
    if ( ichem == 1 ) call get4(MB_TRAC, mo_atm%trac)
    if ( ibltyp == 2 ) call get3(MB_TKE, mo_atm%tke)
    call chk(moloch_b200_set_async(ctx, 0_c_int), __LINE__)      ! one synchronisation for the batch
  end subroutine download_state

  !---------------------------------------------------------------------------
  ! the pieces of `moloch` (Main/mod_moloch.F90:312-446)
  !---------------------------------------------------------------------------
  subroutine b200_dycore                     ! :327 reset_tendencies + :334 dynamical_core
    implicit none
    call chk(moloch_b200_reset_tendencies(ctx), __LINE__)
    call chk(moloch_b200_dynamical_core(ctx), __LINE__)
  end subroutine b200_dycore

  subroutine b200_boundary                   ! :341-343, when the boundary runs on the device
    implicit none
    if ( use_device_bdy ) call chk(moloch_b200_boundary(ctx), __LINE__)
  end subroutine b200_boundary

  ! bdyin (Main/mod_bdycod.F90:1079-1423) has read a new boundary record into the b1 buffers:
  ! b0 <- b1 on the device (a pointer swap), then the new b1 arrays
  subroutine b200_bdyin(dub1, dvb1, xtb1, xpaib1, xqb1, xpsb1)
    implicit none
    real(rkx), pointer, contiguous, intent(in) :: dub1(:,:,:), dvb1(:,:,:), xtb1(:,:,:), xpaib1(:,:,:), xqb1(:,:,:)
    real(rkx), pointer, contiguous, intent(in) :: xpsb1(:,:)
    if ( .not. use_device_bdy ) return
    call chk(moloch_b200_bdy_shift(ctx), __LINE__)
    call put3(MB_DUB1, dub1) ; call put3(MB_DVB1, dvb1) ; call put3(MB_XTB1, xtb1)
    call put3(MB_XPAIB1, xpaib1) ; call put3(MB_XQB1, xqb1) ; call put2(MB_XPSB1, xpsb1)
  end subroutine b200_bdyin

  subroutine b200_diagnostics(calday, dayspy_)   ! :348-357 p, rho, qsat, ps [, mkslice]
    implicit none
    real(rkx), intent(in) :: calday, dayspy_
    call chk(moloch_b200_diagnostics(ctx), __LINE__)
    if ( use_device_slice ) then
      call chk(moloch_b200_set_calday(ctx, real(calday,c_double), real(dayspy_,c_double)), __LINE__)
      call chk(moloch_b200_mkslice(ctx), __LINE__)
    end if
  end subroutine b200_diagnostics

  ! state to the host arrays (before physical_parametrizations), tendencies back (after it)
  subroutine b200_to_host
    implicit none
    type(moloch_b200_xfer), target :: none(1)
    call chk(moloch_b200_handoff(ctx, xdown, ndown, none, 0_c_int, 8_c_int, c_null_funptr, c_null_ptr), __LINE__)
    call get2(MB_PS, sfs%psa)
  end subroutine b200_to_host

  subroutine b200_from_host
    implicit none
    type(moloch_b200_xfer), target :: none(1)
    call chk(moloch_b200_handoff(ctx, none, 0_c_int, xup, nup, 8_c_int, c_null_funptr, c_null_ptr), __LINE__)
  end subroutine b200_from_host

  ! both directions in one pipelined call; physics_slab(user, i1, i2) is a bind(C) wrapper that runs the
  ! column physics on rows i1:i2 (c_null_funptr: plain exchange)
  subroutine b200_handoff(physics_slab)
    implicit none
    type(c_funptr), intent(in) :: physics_slab
    call chk(moloch_b200_handoff(ctx, xdown, ndown, xup, nup, 8_c_int, physics_slab, c_null_ptr), __LINE__)
    call get2(MB_PS, sfs%psa)
  end subroutine b200_handoff

  subroutine b200_status_update              ! :376
    implicit none
    call chk(moloch_b200_status_update(ctx), __LINE__)
  end subroutine b200_status_update

  ! restart / history (Main/mod_savefile.F90:232-242, 618-627): the save set, asynchronously
  subroutine b200_savefile_state
    implicit none
    call download_state
  end subroutine b200_savefile_state

  ! the CFL guard of moloch (:407-422): max / min of ps over the interior, non-finite count
  subroutine b200_ps_report(maxps, minps, nbad)
    implicit none
    real(rkx), intent(out) :: maxps, minps
    integer(ik4), intent(out) :: nbad
    real(c_double) :: mm(2)
    integer(c_int32_t), target :: bad(1)
    call chk(moloch_b200_ps_check(ctx, mm, bad), __LINE__)
    maxps = real(mm(1), rkx) ; minps = real(mm(2), rkx) ; nbad = bad(1)
  end subroutine b200_ps_report

  ! the atmosphere sums of massck (Main/mod_massck.F90:77-185): this rank's partial sums
  subroutine b200_massck(tdrym, tdadv, tqmass, tqadv)
    implicit none
    real(rk8), intent(out) :: tdrym, tdadv, tqmass, tqadv
    real(c_double) :: o(4)
    call chk(moloch_b200_massck(ctx, o), __LINE__)
    tdrym = o(1) ; tdadv = o(2) ; tqmass = o(3) ; tqadv = o(4)
  end subroutine b200_massck

  subroutine b200_finalize
    implicit none
    integer :: q
    do q = 1, ndown
      call chk(moloch_b200_host_unregister(xdown(q)%host), __LINE__)
    end do
    do q = 1, nup
      call chk(moloch_b200_host_unregister(xup(q)%host), __LINE__)
    end do
    call chk(moloch_b200_destroy(ctx), __LINE__)
    ctx = c_null_ptr
  end subroutine b200_finalize

end module mod_moloch_b200
