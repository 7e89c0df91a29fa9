!-----------------------------------------------------------------------------------------------------------
! wuming_b200_shim.f90 -- ISO_C_BINDING layer between WumingPIC's unmodified drivers (proj/*/app.f90) and
! libwuming_b200.so (include/wuming_b200.h).
!
! It provides modules with the SAME names, public procedures and argument lists as the reference's
!     3d/common/particle.f90           (particle__init, particle__solv, particle__solv_vay)
!     3d/common/field.f90              (field__init, field__fdtd_i)
!     3d/common/sort.f90               (sort__init, sort__bucket)
!     3d/common/boundary_periodic.f90  (boundary_periodic__init, __particle_x, __particle_yz, __dfield, __curre, __phi)
!     3d/proj/reconnection/boundary_reconnection.f90  (boundary_reconnection__init, __particle_x, __particle_yz, ...)
!     3d/proj/shock/boundary_shock.f90                (boundary_shock__init, __injection, __particle_yz, ...)
! so that `use wuming3d` / `use boundary_periodic, bc__init => boundary_periodic__init, ...` in
! 3d/proj/weibel/app.f90:1-58 keep compiling.  Link these objects INSTEAD of the four reference files
! (INTEGRATION.md shows the two-line Makefile change); everything else of libwuming3d_common.a (mpi_set, paraio,
! mom_calc) and the driver stay as they are.
!
! NOTE: this image has no Fortran compiler (gfortran / f951 / nvfortran absent), so this file is checked by
! review against the reference's interface blocks only; the C ABI underneath is exercised by tests/ through
! ctypes with the same call sequence.
!
! Host arrays are a cache of the device state (SURVEY.md 8b).  Two modes, set with wm_shim_set_mode():
!   WM_SHIM_SYNC_EVERY_CALL (default)  every procedure uploads its intent(in) arrays and downloads its
!                                      intent(out) arrays -- a bit-for-bit drop-in for an unmodified driver.
!   WM_SHIM_RESIDENT                   state stays on the GPU between calls; the driver calls
!                                      wm_shim_sync_to_host(up,uf,np2,cumcnt) before it reads the arrays
!                                      (io__ptcl / io__mom / energy_history / save_restart) and
!                                      wm_shim_host_modified() after it writes them (shock inject / relocate).
!-----------------------------------------------------------------------------------------------------------
module wuming_b200_c
  use iso_c_binding
  implicit none
  public

  integer(c_int), parameter :: WM_BC_PERIODIC = 0, WM_BC_RECONNECTION = 1, WM_BC_SHOCK = 2
  integer, parameter :: WM_SHIM_SYNC_EVERY_CALL = 0, WM_SHIM_RESIDENT = 1

  type, bind(c) :: wm_params            ! include/wuming_b200.h: struct wm_params
    integer(c_int) :: dim, ndim, np, nsp
    integer(c_int) :: nxgs, nxge, nygs, nyge, nzgs, nzge
    integer(c_int) :: nys, nye, nzs, nze
    integer(c_int) :: nproc_j, nproc_k, rank_j, rank_k
    integer(c_int) :: bc_kind, device
    real(c_double) :: delx, delt, c, gfac
    real(c_double) :: q(2), r(2)
  end type wm_params

  type, bind(c) :: wm_shock_params      ! include/wuming_b200.h: struct wm_shock_params (2d/proj/shock/app.f90 constants)
    integer(c_int)       :: n0
    real(c_double)       :: v0, v_thi, v_the, b0, theta_bn, phi_bn, l_damp_ini
    integer(c_long_long) :: seed
  end type wm_shock_params

  type(c_ptr), save :: ctx = c_null_ptr
  type(wm_params), save :: prm
  integer, save :: shim_mode = WM_SHIM_SYNC_EVERY_CALL
  logical, save :: host_dirty = .true.    ! host arrays are newer than the device copy
  logical, save :: have_q = .false., have_gfac = .false., have_geom = .false.

  interface
    function wm_create(p, out) bind(c, name='wm_create') result(ierr)
      import :: c_int, c_ptr, wm_params
      type(wm_params), intent(in) :: p
      type(c_ptr), intent(out)    :: out
      integer(c_int)              :: ierr
    end function
    function wm_destroy(c) bind(c, name='wm_destroy') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: c
      integer(c_int)     :: ierr
    end function
    function wm_comm_unique_id(id) bind(c, name='wm_comm_unique_id') result(ierr)
      import :: c_int, c_char
      character(kind=c_char) :: id(128)
      integer(c_int)         :: ierr
    end function
    function wm_comm_init(c, nranks, rank, id) bind(c, name='wm_comm_init') result(ierr)
      import :: c_int, c_ptr, c_char
      type(c_ptr), value     :: c
      integer(c_int), value  :: nranks, rank
      character(kind=c_char) :: id(128)
      integer(c_int)         :: ierr
    end function
    function wm_upload(c, up, np2, cumcnt, uf) bind(c, name='wm_upload') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: c, up, np2, cumcnt, uf       ! c_null_ptr skips an array
      integer(c_int)     :: ierr
    end function
    function wm_download(c, up, np2, cumcnt, uf, gp) bind(c, name='wm_download') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: c, up, np2, cumcnt, uf, gp
      integer(c_int)     :: ierr
    end function
    function wm_particle_solv(c, nxs, nxe) bind(c, name='wm_particle_solv') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value    :: c
      integer(c_int), value :: nxs, nxe
      integer(c_int)        :: ierr
    end function
    function wm_particle_solv_vay(c, nxs, nxe) bind(c, name='wm_particle_solv_vay') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value    :: c
      integer(c_int), value :: nxs, nxe
      integer(c_int)        :: ierr
    end function
    function wm_field_fdtd_i(c, nxs, nxe) bind(c, name='wm_field_fdtd_i') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value    :: c
      integer(c_int), value :: nxs, nxe
      integer(c_int)        :: ierr
    end function
    function wm_bc_particle_x(c, nxs, nxe) bind(c, name='wm_bc_particle_x') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value    :: c
      integer(c_int), value :: nxs, nxe
      integer(c_int)        :: ierr
    end function
    function wm_bc_injection(c, nxs, nxe, u0) bind(c, name='wm_bc_injection') result(ierr)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value    :: c
      integer(c_int), value :: nxs, nxe
      real(c_double), value :: u0
      integer(c_int)        :: ierr
    end function
    function wm_bc_particle_yz(c) bind(c, name='wm_bc_particle_yz') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value :: c
      integer(c_int)     :: ierr
    end function
    function wm_sort_bucket(c, nxs, nxe) bind(c, name='wm_sort_bucket') result(ierr)
      import :: c_int, c_ptr
      type(c_ptr), value    :: c
      integer(c_int), value :: nxs, nxe
      integer(c_int)        :: ierr
    end function
    function wm_last_error() bind(c, name='wm_last_error') result(msg)
      import :: c_ptr
      type(c_ptr) :: msg
    end function
  end interface

contains

  ! print the library's message and stop, like the reference's `write(6,*) ...; stop`
  ! (3d/common/particle.f90:69-72, field.f90:522-525, boundary_periodic.f90:435-438)
  subroutine wm_check(ierr, where)
    integer(c_int), intent(in)   :: ierr
    character(len=*), intent(in) :: where
    character(kind=c_char), pointer :: s(:)
    integer :: n
    if (ierr == 0) return
    call c_f_pointer(wm_last_error(), s, [512])
    n = 1
    do while (n < 512 .and. s(n) /= c_null_char)
      n = n + 1
    end do
    write(6,*) where, ': ', s(1:n-1)
    stop
  end subroutine wm_check

  ! called by each __init; creates the context once geometry, (q,r) and gfac are all known
  subroutine wm_shim_try_create()
    integer(c_int) :: ierr
    if (c_associated(ctx)) return
    if (.not.(have_geom .and. have_q .and. have_gfac)) return
    prm%device = -1                       ! current device; set CUDA_VISIBLE_DEVICES per rank, or rank mod ngpu
    ierr = wm_create(prm, ctx)
    call wm_check(ierr, 'wm_create')
  end subroutine wm_shim_try_create

  subroutine wm_shim_set_geom(ndim,np,nsp,nxgs,nxge,nygs,nyge,nzgs,nzge,nys,nye,nzs,nze,delx,delt,c)
    integer, intent(in) :: ndim,np,nsp,nxgs,nxge,nygs,nyge,nzgs,nzge,nys,nye,nzs,nze
    real(8), intent(in) :: delx,delt,c
    prm%dim = 3;  prm%ndim = ndim; prm%np = np; prm%nsp = nsp
    prm%nxgs = nxgs; prm%nxge = nxge; prm%nygs = nygs; prm%nyge = nyge; prm%nzgs = nzgs; prm%nzge = nzge
    prm%nys = nys; prm%nye = nye; prm%nzs = nzs; prm%nze = nze
    prm%delx = delx; prm%delt = delt; prm%c = c
    have_geom = .true.
  end subroutine wm_shim_set_geom

  subroutine wm_shim_set_mode(mode)
    integer, intent(in) :: mode
    shim_mode = mode
  end subroutine wm_shim_set_mode

  subroutine wm_shim_host_modified()
    host_dirty = .true.
  end subroutine wm_shim_host_modified

  ! the rank grid of mpi_set__init (3d/common/mpi_set.f90:45-60) and the NCCL communicator: call once after
  ! the __init calls, e.g. right after bc__init in app.f90:341.  `id` comes from rank 0 via MPI_Bcast.
  subroutine wm_shim_comm_init(nproc, nproc_j, nproc_k, nrank, ncomw)
    integer, intent(in) :: nproc, nproc_j, nproc_k, nrank, ncomw
    character(kind=c_char) :: id(128)
    integer :: nerr
    integer(c_int) :: ierr
    include 'mpif.h'
    prm%nproc_j = nproc_j; prm%nproc_k = nproc_k
    prm%rank_j = nrank / nproc_k; prm%rank_k = mod(nrank, nproc_k)
    call wm_shim_try_create()
    if (nproc == 1) return
    if (nrank == 0) then
      ierr = wm_comm_unique_id(id)
      call wm_check(ierr, 'wm_comm_unique_id')
    end if
    call MPI_BCAST(id, 128, MPI_CHARACTER, 0, ncomw, nerr)
    ierr = wm_comm_init(ctx, int(nproc, c_int), int(nrank, c_int), id)
    call wm_check(ierr, 'wm_comm_init')
  end subroutine wm_shim_comm_init

  subroutine wm_shim_sync_to_host(up, uf, np2, cumcnt)
    real(8), intent(inout), target    :: up(*), uf(*)
    integer, intent(inout), target    :: np2(*), cumcnt(*)
    integer(c_int) :: ierr
    ierr = wm_download(ctx, c_loc(up), c_loc(np2), c_loc(cumcnt), c_loc(uf), c_null_ptr)
    call wm_check(ierr, 'wm_download')
    host_dirty = .false.
  end subroutine wm_shim_sync_to_host

end module wuming_b200_c

!-----------------------------------------------------------------------------------------------------------
module particle                      ! replaces 3d/common/particle.f90
  use iso_c_binding
  use wuming_b200_c
  implicit none
  private
  public :: particle__init, particle__solv, particle__solv_vay
  integer, save :: ndim, np, nsp, nxgs, nxge, nygs, nyge, nzgs, nzge, nys, nye, nzs, nze
  logical, save :: is_init = .false.
contains

  subroutine particle__init(ndim_in,np_in,nsp_in,nxgs_in,nxge_in,nygs_in,nyge_in,nzgs_in,nzge_in,nys_in,nye_in,nzs_in,nze_in, &
                            delx_in,delt_in,c_in,q_in,r_in)                    ! particle.f90:18-49
    integer, intent(in) :: ndim_in, np_in, nsp_in
    integer, intent(in) :: nxgs_in, nxge_in, nygs_in, nyge_in, nzgs_in, nzge_in, nys_in, nye_in, nzs_in, nze_in
    real(8), intent(in) :: delx_in, delt_in, c_in, q_in(nsp_in), r_in(nsp_in)
    ndim = ndim_in; np = np_in; nsp = nsp_in
    nxgs = nxgs_in; nxge = nxge_in; nygs = nygs_in; nyge = nyge_in; nzgs = nzgs_in; nzge = nzge_in
    nys = nys_in; nye = nye_in; nzs = nzs_in; nze = nze_in
    call wm_shim_set_geom(ndim,np,nsp,nxgs,nxge,nygs,nyge,nzgs,nzge,nys,nye,nzs,nze,delx_in,delt_in,c_in)
    prm%q(1:2) = q_in(1:2); prm%r(1:2) = r_in(1:2); have_q = .true.
    call wm_shim_try_create()
    is_init = .true.
  end subroutine particle__init

  subroutine particle__solv(gp,up,uf,cumcnt,nxs,nxe)                            ! particle.f90:52-233
    integer, intent(in)          :: nxs, nxe
    integer, intent(in), target  :: cumcnt(nxgs:nxge+1,nys:nye,nzs:nze,nsp)
    real(8), intent(in), target  :: up(ndim,np,nys:nye,nzs:nze,nsp)
    real(8), intent(in), target  :: uf(6,nxgs-2:nxge+2,nys-2:nye+2,nzs-2:nze+2)
    real(8), intent(out), target :: gp(ndim,np,nys:nye,nzs:nze,nsp)
    integer, target :: np2(nys:nye,nzs:nze,nsp)
    integer(c_int)  :: ierr
    if(.not.is_init)then
      write(6,*)'Initialize first by calling particle__init()'
      stop
    endif
    if (shim_mode == WM_SHIM_SYNC_EVERY_CALL .or. host_dirty) then
      np2(:,:,:) = cumcnt(nxe+1,:,:,:)          ! the pencil population is the last prefix count (sort.f90:71-74)
      ierr = wm_upload(ctx, c_loc(up), c_loc(np2), c_loc(cumcnt), c_loc(uf))
      call wm_check(ierr, 'wm_upload')
      host_dirty = .false.
    end if
    ierr = wm_particle_solv(ctx, int(nxs,c_int), int(nxe,c_int))
    call wm_check(ierr, 'particle__solv')
    if (shim_mode == WM_SHIM_SYNC_EVERY_CALL) then
      ierr = wm_download(ctx, c_null_ptr, c_null_ptr, c_null_ptr, c_null_ptr, c_loc(gp))
      call wm_check(ierr, 'wm_download(gp)')
    end if
  end subroutine particle__solv

  subroutine particle__solv_vay(gp,up,uf,cumcnt,nxs,nxe)                            ! particle.f90:236-419
    integer, intent(in)          :: nxs, nxe
    integer, intent(in), target  :: cumcnt(nxgs:nxge+1,nys:nye,nzs:nze,nsp)
    real(8), intent(in), target  :: up(ndim,np,nys:nye,nzs:nze,nsp)
    real(8), intent(in), target  :: uf(6,nxgs-2:nxge+2,nys-2:nye+2,nzs-2:nze+2)
    real(8), intent(out), target :: gp(ndim,np,nys:nye,nzs:nze,nsp)
    integer, target :: np2(nys:nye,nzs:nze,nsp)
    integer(c_int)  :: ierr
    if(.not.is_init)then
      write(6,*)'Initialize first by calling particle__init()'
      stop
    endif
    if (shim_mode == WM_SHIM_SYNC_EVERY_CALL .or. host_dirty) then
      np2(:,:,:) = cumcnt(nxe+1,:,:,:)          ! the pencil population is the last prefix count (sort.f90:71-74)
      ierr = wm_upload(ctx, c_loc(up), c_loc(np2), c_loc(cumcnt), c_loc(uf))
      call wm_check(ierr, 'wm_upload')
      host_dirty = .false.
    end if
    ierr = wm_particle_solv_vay(ctx, int(nxs,c_int), int(nxe,c_int))
    call wm_check(ierr, 'particle__solv_vay')
    if (shim_mode == WM_SHIM_SYNC_EVERY_CALL) then
      ierr = wm_download(ctx, c_null_ptr, c_null_ptr, c_null_ptr, c_null_ptr, c_loc(gp))
      call wm_check(ierr, 'wm_download(gp)')
    end if
  end subroutine particle__solv_vay

end module particle

!-----------------------------------------------------------------------------------------------------------
module field                         ! replaces 3d/common/field.f90
  use iso_c_binding
  use wuming_b200_c
  implicit none
  private
  public :: field__init, field__fdtd_i
  integer, save :: ndim, np, nsp, nxgs, nxge, nygs, nyge, nzgs, nzge, nys, nye, nzs, nze
  logical, save :: is_init = .false.
contains

  subroutine field__init(ndim_in,np_in,nsp_in,nxgs_in,nxge_in,nygs_in,nyge_in,nzgs_in,nzge_in,nys_in,nye_in,nzs_in,nze_in, &
                         mnpr_in,ncomw_in,opsum_in,nerr_in,                                                                &
                         delx_in,delt_in,c_in,q_in,r_in,gfac_in)                ! field.f90:22-67
    integer, intent(in) :: ndim_in, np_in, nsp_in
    integer, intent(in) :: nxgs_in, nxge_in, nygs_in, nyge_in, nzgs_in, nzge_in, nys_in, nye_in, nzs_in, nze_in
    integer, intent(in) :: mnpr_in, ncomw_in, opsum_in, nerr_in                  ! MPI handles: unused, NCCL replaces them
    real(8), intent(in) :: delx_in, delt_in, c_in, q_in(nsp_in), r_in(nsp_in), gfac_in
    ndim = ndim_in; np = np_in; nsp = nsp_in
    nxgs = nxgs_in; nxge = nxge_in; nygs = nygs_in; nyge = nyge_in; nzgs = nzgs_in; nzge = nzge_in
    nys = nys_in; nye = nye_in; nzs = nzs_in; nze = nze_in
    prm%gfac = gfac_in; have_gfac = .true.
    call wm_shim_try_create()
    is_init = .true.
  end subroutine field__init

  ! The three boundary procedures arrive as dummy arguments exactly as in field.f90:70-93; the device applies
  ! the boundary kind registered by the boundary_* module's __init (prm%bc_kind), so they are accepted and not called.
  subroutine field__fdtd_i(uf,up,gp,cumcnt,nxs,nxe,set_boundary_dfield,set_boundary_curre,set_boundary_phi)
    external :: set_boundary_dfield, set_boundary_curre, set_boundary_phi
    integer, intent(in)            :: nxs, nxe
    integer, intent(in), target    :: cumcnt(nxgs:nxge+1,nys:nye,nzs:nze,nsp)
    real(8), intent(in), target    :: gp(ndim,np,nys:nye,nzs:nze,nsp)
    real(8), intent(in), target    :: up(ndim,np,nys:nye,nzs:nze,nsp)
    real(8), intent(inout), target :: uf(6,nxgs-2:nxge+2,nys-2:nye+2,nzs-2:nze+2)
    integer(c_int) :: ierr
    if(.not.is_init)then
      write(6,*)'Initialize first by calling field__init()'
      stop
    endif
    ! up, gp and cumcnt are the arrays particle__solv just consumed/produced: already on the device
    ierr = wm_field_fdtd_i(ctx, int(nxs,c_int), int(nxe,c_int))
    call wm_check(ierr, 'field__fdtd_i')          ! WM_ERR_CG_ITEMAX reproduces "stop at cgm after ite_max"
    if (shim_mode == WM_SHIM_SYNC_EVERY_CALL) then
      ierr = wm_download(ctx, c_null_ptr, c_null_ptr, c_null_ptr, c_loc(uf), c_null_ptr)
      call wm_check(ierr, 'wm_download(uf)')
    end if
  end subroutine field__fdtd_i

end module field

!-----------------------------------------------------------------------------------------------------------
module sort                          ! replaces 3d/common/sort.f90
  use iso_c_binding
  use wuming_b200_c
  implicit none
  private
  public :: sort__init, sort__bucket
  integer, save :: ndim, np, nsp, nxgs, nxge, nygs, nyge, nzgs, nzge, nys, nye, nzs, nze
  logical, save :: is_init = .false.
contains

  subroutine sort__init(ndim_in,np_in,nsp_in,nxgs_in,nxge_in,nygs_in,nyge_in,nzgs_in,nzge_in,nys_in,nye_in,nzs_in,nze_in)
    integer, intent(in) :: ndim_in, np_in, nsp_in
    integer, intent(in) :: nxgs_in, nxge_in, nygs_in, nyge_in, nzgs_in, nzge_in, nys_in, nye_in, nzs_in, nze_in
    ndim = ndim_in; np = np_in; nsp = nsp_in
    nxgs = nxgs_in; nxge = nxge_in; nygs = nygs_in; nyge = nyge_in; nzgs = nzgs_in; nzge = nzge_in
    nys = nys_in; nye = nye_in; nzs = nzs_in; nze = nze_in
    is_init = .true.
  end subroutine sort__init

  ! NB the reference's dummy names: the FIRST argument is the sorted output (sort.f90:40-47)
  subroutine sort__bucket(gp,up,cumcnt,np2,nxs,nxe)
    integer, intent(in)           :: nxs, nxe
    integer, intent(in), target   :: np2(nys:nye,nzs:nze,nsp)
    integer, intent(out), target  :: cumcnt(nxgs:nxge+1,nys:nye,nzs:nze,nsp)
    real(8), intent(in), target   :: up(ndim,np,nys:nye,nzs:nze,nsp)
    real(8), intent(out), target  :: gp(ndim,np,nys:nye,nzs:nze,nsp)
    integer, target :: np2_out(nys:nye,nzs:nze,nsp)
    integer(c_int)  :: ierr
    if(.not.is_init)then
      write(6,*)'Initialize first by calling sort__init()'
      stop
    endif
    ierr = wm_sort_bucket(ctx, int(nxs,c_int), int(nxe,c_int))
    call wm_check(ierr, 'sort__bucket')           ! WM_ERR_MEMORY_OVER reproduces "memory over (np2 > np)"
    if (shim_mode == WM_SHIM_SYNC_EVERY_CALL) then
      ierr = wm_download(ctx, c_loc(gp), c_loc(np2_out), c_loc(cumcnt), c_null_ptr, c_null_ptr)
      call wm_check(ierr, 'wm_download(up)')
    end if
  end subroutine sort__bucket

end module sort

!-----------------------------------------------------------------------------------------------------------
module boundary_periodic             ! replaces 3d/common/boundary_periodic.f90
  use iso_c_binding
  use wuming_b200_c
  implicit none
  private
  public :: boundary_periodic__init
  public :: boundary_periodic__dfield, boundary_periodic__particle_x, boundary_periodic__particle_yz
  public :: boundary_periodic__curre, boundary_periodic__phi
  integer, save :: ndim, np, nsp, nxgs, nxge, nygs, nyge, nzgs, nzge, nys, nye, nzs, nze
  logical, save :: is_init = .false.
contains

  subroutine boundary_periodic__init(ndim_in,np_in,nsp_in,nxgs_in,nxge_in,nygs_in,nyge_in,nzgs_in, &
       & nzge_in,nys_in,nye_in,nzs_in,nze_in,jup_in,jdown_in,kup_in,kdown_in,mnpi_in,mnpr_in, &
       & ncomw_in,nerr_in,nstat_in,delx_in,delt_in,c_in)                       ! boundary_periodic.f90:25-65
    integer, intent(in) :: ndim_in, np_in, nsp_in
    integer, intent(in) :: nxgs_in, nxge_in, nygs_in, nyge_in, nzgs_in, nzge_in, nys_in, nye_in, nzs_in, nze_in
    integer, intent(in) :: jup_in, jdown_in, kup_in, kdown_in, mnpi_in, mnpr_in, ncomw_in, nerr_in, nstat_in(:)
    real(8), intent(in) :: delx_in, delt_in, c_in
    ndim = ndim_in; np = np_in; nsp = nsp_in
    nxgs = nxgs_in; nxge = nxge_in; nygs = nygs_in; nyge = nyge_in; nzgs = nzgs_in; nzge = nzge_in
    nys = nys_in; nye = nye_in; nzs = nzs_in; nze = nze_in
    prm%bc_kind = WM_BC_PERIODIC                  ! boundary_shock / boundary_reconnection shims set their own kind
    call wm_shim_set_geom(ndim,np,nsp,nxgs,nxge,nygs,nyge,nzgs,nzge,nys,nye,nzs,nze,delx_in,delt_in,c_in)
    is_init = .true.
  end subroutine boundary_periodic__init

  subroutine boundary_periodic__particle_x(up,np2)                             ! boundary_periodic.f90:68-101
    real(8), intent(inout) :: up(ndim,np,nys:nye,nzs:nze,nsp)
    integer, intent(in)    :: np2(nys:nye,nzs:nze,nsp)
    integer(c_int) :: ierr
    ierr = wm_bc_particle_x(ctx, int(nxgs,c_int), int(nxge,c_int))
    call wm_check(ierr, 'bc__particle_x')
  end subroutine boundary_periodic__particle_x

  subroutine boundary_periodic__particle_yz(up,np2)                            ! boundary_periodic.f90:104-455
    real(8), intent(inout) :: up(ndim,np,nys:nye,nzs:nze,nsp)
    integer, intent(inout) :: np2(nys:nye,nzs:nze,nsp)
    integer(c_int) :: ierr
    ierr = wm_bc_particle_yz(ctx)
    call wm_check(ierr, 'bc__particle_yz')
    ! np2 is refreshed by sort__bucket's download (the movers travel inside the device sort)
  end subroutine boundary_periodic__particle_yz

  ! The field-side boundary procedures are only ever passed to field__fdtd_i (app.f90:104-105); the device applies
  ! them inside wm_field_fdtd_i.  They exist so that the driver's procedure arguments resolve.
  subroutine boundary_periodic__dfield(df,nxs,nxe,nys_in,nye_in,nzs_in,nze_in,nxgs_in,nxge_in)
    integer, intent(in)    :: nxs, nxe, nys_in, nye_in, nzs_in, nze_in, nxgs_in, nxge_in
    real(8), intent(inout) :: df(6,nxgs_in-2:nxge_in+2,nys_in-2:nye_in+2,nzs_in-2:nze_in+2)
  end subroutine boundary_periodic__dfield
  subroutine boundary_periodic__curre(uj,nxs,nxe,nys_in,nye_in,nzs_in,nze_in,nxgs_in,nxge_in)
    integer, intent(in)    :: nxs, nxe, nys_in, nye_in, nzs_in, nze_in, nxgs_in, nxge_in
    real(8), intent(inout) :: uj(3,nxgs_in-2:nxge_in+2,nys_in-2:nye_in+2,nzs_in-2:nze_in+2)
  end subroutine boundary_periodic__curre
  subroutine boundary_periodic__phi(phi,nxs,nxe,nys_in,nye_in,nzs_in,nze_in,l)
    integer, intent(in)    :: nxs, nxe, nys_in, nye_in, nzs_in, nze_in, l
    real(8), intent(inout) :: phi(nxs-1:nxe+1,nys_in-1:nye_in+1,nzs_in-1:nze_in+1)
  end subroutine boundary_periodic__phi

end module boundary_periodic

!-----------------------------------------------------------------------------------------------------------
! Wall set-ups.  Same pattern as boundary_periodic: __init registers the boundary kind (so that wm_field_fdtd_i applies
! the conducting-wall rules of __dfield / __phi and skips the x fold of __curre), the particle procedures forward to the
! C ABI, the field-side procedures only exist so that the driver's procedure arguments resolve.
!-----------------------------------------------------------------------------------------------------------
module boundary_reconnection         ! replaces 3d/proj/reconnection/boundary_reconnection.f90
  use iso_c_binding
  use wuming_b200_c
  implicit none
  private
  public :: boundary_reconnection__init
  public :: boundary_reconnection__dfield, boundary_reconnection__particle_x, boundary_reconnection__particle_yz
  public :: boundary_reconnection__curre, boundary_reconnection__phi
  integer, save :: ndim, np, nsp, nxgs, nxge, nygs, nyge, nzgs, nzge, nys, nye, nzs, nze
contains

  subroutine boundary_reconnection__init(ndim_in,np_in,nsp_in,nxgs_in,nxge_in,nygs_in,nyge_in,nzgs_in, &
       & nzge_in,nys_in,nye_in,nzs_in,nze_in,jup_in,jdown_in,kup_in,kdown_in,mnpi_in,mnpr_in, &
       & ncomw_in,nerr_in,nstat_in,delx_in,delt_in,c_in)                       ! boundary_reconnection.f90:26-66
    integer, intent(in) :: ndim_in, np_in, nsp_in
    integer, intent(in) :: nxgs_in, nxge_in, nygs_in, nyge_in, nzgs_in, nzge_in, nys_in, nye_in, nzs_in, nze_in
    integer, intent(in) :: jup_in, jdown_in, kup_in, kdown_in, mnpi_in, mnpr_in, ncomw_in, nerr_in, nstat_in(:)
    real(8), intent(in) :: delx_in, delt_in, c_in
    ndim = ndim_in; np = np_in; nsp = nsp_in
    nxgs = nxgs_in; nxge = nxge_in; nygs = nygs_in; nyge = nyge_in; nzgs = nzgs_in; nzge = nzge_in
    nys = nys_in; nye = nye_in; nzs = nzs_in; nze = nze_in
    prm%bc_kind = WM_BC_RECONNECTION
    call wm_shim_set_geom(ndim,np,nsp,nxgs,nxge,nygs,nyge,nzgs,nzge,nys,nye,nzs,nze,delx_in,delt_in,c_in)
  end subroutine boundary_reconnection__init

  subroutine boundary_reconnection__particle_x(up,np2,nxs,nxe)                 ! boundary_reconnection.f90:69-110
    integer, intent(in)    :: nxs, nxe
    real(8), intent(inout) :: up(ndim,np,nys:nye,nzs:nze,nsp)
    integer, intent(in)    :: np2(nys:nye,nzs:nze,nsp)
    integer(c_int) :: ierr
    ierr = wm_bc_particle_x(ctx, int(nxs,c_int), int(nxe,c_int))              ! reflecting walls at nxs+1, nxe-1
    call wm_check(ierr, 'bc__particle_x')
  end subroutine boundary_reconnection__particle_x

  subroutine boundary_reconnection__particle_yz(up,np2)                        ! boundary_reconnection.f90:112-464
    real(8), intent(inout) :: up(ndim,np,nys:nye,nzs:nze,nsp)
    integer, intent(inout) :: np2(nys:nye,nzs:nze,nsp)
    integer(c_int) :: ierr
    ierr = wm_bc_particle_yz(ctx)
    call wm_check(ierr, 'bc__particle_yz')
  end subroutine boundary_reconnection__particle_yz

  subroutine boundary_reconnection__dfield(df,nxs,nxe,nys_in,nye_in,nzs_in,nze_in,nxgs_in,nxge_in)
    integer, intent(in)    :: nxs, nxe, nys_in, nye_in, nzs_in, nze_in, nxgs_in, nxge_in
    real(8), intent(inout) :: df(6,nxgs_in-2:nxge_in+2,nys_in-2:nye_in+2,nzs_in-2:nze_in+2)
  end subroutine boundary_reconnection__dfield
  subroutine boundary_reconnection__curre(uj,nxs,nxe,nys_in,nye_in,nzs_in,nze_in,nxgs_in,nxge_in)
    integer, intent(in)    :: nxs, nxe, nys_in, nye_in, nzs_in, nze_in, nxgs_in, nxge_in
    real(8), intent(inout) :: uj(3,nxgs_in-2:nxge_in+2,nys_in-2:nye_in+2,nzs_in-2:nze_in+2)
  end subroutine boundary_reconnection__curre
  subroutine boundary_reconnection__phi(phi,nxs,nxe,nys_in,nye_in,nzs_in,nze_in,l)
    integer, intent(in)    :: nxs, nxe, nys_in, nye_in, nzs_in, nze_in, l
    real(8), intent(inout) :: phi(nxs-1:nxe+1,nys_in-1:nye_in+1,nzs_in-1:nze_in+1)
  end subroutine boundary_reconnection__phi

end module boundary_reconnection

!-----------------------------------------------------------------------------------------------------------
module boundary_shock                ! replaces 3d/proj/shock/boundary_shock.f90
  use iso_c_binding
  use wuming_b200_c
  implicit none
  private
  public :: boundary_shock__init
  public :: boundary_shock__dfield, boundary_shock__particle_yz, boundary_shock__injection
  public :: boundary_shock__curre, boundary_shock__phi
  integer, save :: ndim, np, nsp, nxgs, nxge, nygs, nyge, nzgs, nzge, nys, nye, nzs, nze
contains

  subroutine boundary_shock__init(ndim_in,np_in,nsp_in,nxgs_in,nxge_in,nygs_in,nyge_in,nzgs_in,nzge_in, &
       & nys_in,nye_in,nzs_in,nze_in,jup_in,jdown_in,kup_in,kdown_in,mnpi_in,mnpr_in, &
       & ncomw_in,nerr_in,nstat_in,delx_in,delt_in,c_in)                       ! boundary_shock.f90:25-67
    integer, intent(in) :: ndim_in, np_in, nsp_in
    integer, intent(in) :: nxgs_in, nxge_in, nygs_in, nyge_in, nzgs_in, nzge_in, nys_in, nye_in, nzs_in, nze_in
    integer, intent(in) :: jup_in, jdown_in, kup_in, kdown_in, mnpi_in, mnpr_in, ncomw_in, nerr_in, nstat_in(:)
    real(8), intent(in) :: delx_in, delt_in, c_in
    ndim = ndim_in; np = np_in; nsp = nsp_in
    nxgs = nxgs_in; nxge = nxge_in; nygs = nygs_in; nyge = nyge_in; nzgs = nzgs_in; nzge = nzge_in
    nys = nys_in; nye = nye_in; nzs = nzs_in; nze = nze_in
    prm%bc_kind = WM_BC_SHOCK
    call wm_shim_set_geom(ndim,np,nsp,nxgs,nxge,nygs,nyge,nzgs,nzge,nys,nye,nzs,nze,delx_in,delt_in,c_in)
  end subroutine boundary_shock__init

  ! The shock driver mutates up, np2, cumcnt(nxe), uf(:,nxe-1:nxe) and nxe on the host after every sort (inject /
  ! relocate, 3d/proj/shock/app.f90): in WM_SHIM_RESIDENT mode it brackets those calls with wm_shim_sync_to_host() and
  ! wm_shim_host_modified(); the default WM_SHIM_SYNC_EVERY_CALL mode needs no change.
  subroutine boundary_shock__injection(up,np2,nxs,nxe,u0)                      ! boundary_shock.f90:424-469
    integer, intent(in)    :: nxs, nxe
    real(8), intent(inout) :: up(ndim,np,nys:nye,nzs:nze,nsp)
    integer, intent(in)    :: np2(nys:nye,nzs:nze,nsp)
    real(8), intent(in)    :: u0
    integer(c_int) :: ierr
    ierr = wm_bc_injection(ctx, int(nxs,c_int), int(nxe,c_int), real(u0,c_double))
    call wm_check(ierr, 'bc__injection')
  end subroutine boundary_shock__injection

  subroutine boundary_shock__particle_yz(up,np2)                               ! boundary_shock.f90:70-421
    real(8), intent(inout) :: up(ndim,np,nys:nye,nzs:nze,nsp)
    integer, intent(inout) :: np2(nys:nye,nzs:nze,nsp)
    integer(c_int) :: ierr
    ierr = wm_bc_particle_yz(ctx)
    call wm_check(ierr, 'bc__particle_yz')
  end subroutine boundary_shock__particle_yz

  subroutine boundary_shock__dfield(df,nxs,nxe,nys_in,nye_in,nzs_in,nze_in,nxgs_in,nxge_in)
    integer, intent(in)    :: nxs, nxe, nys_in, nye_in, nzs_in, nze_in, nxgs_in, nxge_in
    real(8), intent(inout) :: df(6,nxgs_in-2:nxge_in+2,nys_in-2:nye_in+2,nzs_in-2:nze_in+2)
  end subroutine boundary_shock__dfield
  subroutine boundary_shock__curre(uj,nxs,nxe,nys_in,nye_in,nzs_in,nze_in,nxgs_in,nxge_in)
    integer, intent(in)    :: nxs, nxe, nys_in, nye_in, nzs_in, nze_in, nxgs_in, nxge_in
    real(8), intent(inout) :: uj(3,nxgs_in-2:nxge_in+2,nys_in-2:nye_in+2,nzs_in-2:nze_in+2)
  end subroutine boundary_shock__curre
  subroutine boundary_shock__phi(phi,nxs,nxe,nys_in,nye_in,nzs_in,nze_in,l)
    integer, intent(in)    :: nxs, nxe, nys_in, nye_in, nzs_in, nze_in, l
    real(8), intent(inout) :: phi(nxs-1:nxe+1,nys_in-1:nye_in+1,nzs_in-1:nze_in+1)
  end subroutine boundary_shock__phi

end module boundary_shock

!-----------------------------------------------------------------------------------------------------------
! Moments (SURVEY.md 8f #1).  The drivers' output block is
!     call mom_calc__accl(gp,up,uf,cumcnt,nxs,nxe); call mom_calc__nvt(mom,gp,np2); call bc__mom(mom); call io__mom(mom,uf,it)
! (3d/proj/weibel/app.f90:121-125).  wm_mom_calc does accl + nvt + the ghost fold of bc__mom on the device-resident
! particles, so __accl is a no-op, __nvt fills mom with the folded result, and the boundary_*__mom shims are no-ops.
!-----------------------------------------------------------------------------------------------------------
module mom_calc                      ! replaces 3d/common/mom_calc.f90
  use iso_c_binding
  use wuming_b200_c
  implicit none
  private
  public :: mom_calc__init, mom_calc__accl, mom_calc__nvt
  integer, save :: ndim, np, nsp, nxgs, nxge, nygs, nyge, nzgs, nzge, nys, nye, nzs, nze
  integer, save :: last_nxs, last_nxe
  interface
    function wm_mom_calc(c, nxs, nxe, mom) bind(c, name='wm_mom_calc') result(ierr)
      import; type(c_ptr), value :: c; integer(c_int), value :: nxs, nxe; type(c_ptr), value :: mom; integer(c_int) :: ierr
    end function
  end interface
contains

  subroutine mom_calc__init(ndim_in,np_in,nsp_in,nxgs_in,nxge_in,nygs_in,nyge_in,nzgs_in,nzge_in,nys_in,nye_in,nzs_in,nze_in, &
                            delx_in,delt_in,c_in,q_in,r_in)                    ! mom_calc.f90:16-47
    integer, intent(in) :: ndim_in, np_in, nsp_in
    integer, intent(in) :: nxgs_in, nxge_in, nygs_in, nyge_in, nzgs_in, nzge_in, nys_in, nye_in, nzs_in, nze_in
    real(8), intent(in) :: delx_in, delt_in, c_in, q_in(nsp_in), r_in(nsp_in)
    ndim = ndim_in; np = np_in; nsp = nsp_in
    nxgs = nxgs_in; nxge = nxge_in; nygs = nygs_in; nyge = nyge_in; nzgs = nzgs_in; nzge = nzge_in
    nys = nys_in; nye = nye_in; nzs = nzs_in; nze = nze_in
    last_nxs = nxgs; last_nxe = nxge
  end subroutine mom_calc__init

  subroutine mom_calc__accl(gp,up,uf,cumcnt,nxs,nxe)                           ! mom_calc.f90:49-216
    integer, intent(in)  :: nxs, nxe
    integer, intent(in)  :: cumcnt(nxgs:nxge+1,nys:nye,nzs:nze,nsp)
    real(8), intent(in)  :: up(ndim,np,nys:nye,nzs:nze,nsp)
    real(8), intent(in)  :: uf(6,nxgs-2:nxge+2,nys-2:nye+2,nzs-2:nze+2)
    real(8), intent(out) :: gp(ndim,np,nys:nye,nzs:nze,nsp)
    last_nxs = nxs; last_nxe = nxe           ! the re-centred momenta never leave the device (wm_mom_calc)
  end subroutine mom_calc__accl

  subroutine mom_calc__nvt(mom,up,np2)                                         ! mom_calc.f90:219-332 (+ the fold of bc__mom)
    integer, intent(in)    :: np2(nys:nye,nzs:nze,nsp)
    real(8), intent(in)    :: up(ndim,np,nys:nye,nzs:nze,nsp)
    real(8), intent(inout), target :: mom(7,nxgs-1:nxge+1,nys-1:nye+1,nzs-1:nze+1,1:nsp)
    integer(c_int) :: ierr
    ierr = wm_mom_calc(ctx, int(last_nxs,c_int), int(last_nxe,c_int), c_loc(mom))
    call wm_check(ierr, 'mom_calc__nvt')
  end subroutine mom_calc__nvt

end module mom_calc
! boundary_periodic__mom / boundary_reconnection__mom / boundary_shock__mom: add to the three boundary modules above
!   subroutine boundary_*__mom(mom); real(8), intent(inout) :: mom(...); end subroutine   (no-op: already folded)

!-----------------------------------------------------------------------------------------------------------
! The shock driver's particle source on the device (include/wuming_b200.h: wm_shock_inject / wm_shock_relocate).
! inject() and relocate() live in the DRIVER (2d/proj/shock/app.f90:615-852, 3d/proj/shock/app.f90:644-906), not in a
! library module, so this one is an edit of app.f90 rather than a module swap: the driver keeps its integer bookkeeping
! (nlinj_grid: app.f90:711-743, get_global_cumsum / ncinj_grid: :769-781) and replaces the two particle loops, the np2 /
! cumcnt updates and the uf columns (:786-849, :637-689) by one call each:
!
!     id_first(j,isp) = ncinj_grid(j) + nptotal(isp)                    ! 3-D: id_first(j,k,isp)
!     call shock_source__inject(nxe, nlinj_grid, id_first, it)
!     ...
!     nxe = nxe + 1
!     id_first(j,isp) = (j-nygs)*n0 + nptotal(isp)                      ! 3-D: ((nyge-nygs+1)*(k-nzgs)+(j-nygs))*n0 + nptotal(isp)
!     call shock_source__relocate(nxe, id_first, it)
!
! np2 / cumcnt / up / uf then change on the device only; the host copies are refreshed at the driver's output cadence by
! wm_shim_sync_to_host, exactly like the rest of the resident state.
!-----------------------------------------------------------------------------------------------------------
module shock_source
  use iso_c_binding
  use wuming_b200_c
  implicit none
  private
  public :: shock_source__init, shock_source__inject, shock_source__relocate
  type(wm_shock_params), save :: sprm
  interface
    function wm_shock_inject(c, p, nxe, nlinj, id_first, epoch) bind(c, name='wm_shock_inject') result(ierr)
      import; type(c_ptr), value :: c; type(wm_shock_params), intent(in) :: p; integer(c_int), value :: nxe
      type(c_ptr), value :: nlinj, id_first; integer(c_long_long), value :: epoch; integer(c_int) :: ierr
    end function
    function wm_shock_relocate(c, p, nxe_new, id_first, epoch) bind(c, name='wm_shock_relocate') result(ierr)
      import; type(c_ptr), value :: c; type(wm_shock_params), intent(in) :: p; integer(c_int), value :: nxe_new
      type(c_ptr), value :: id_first; integer(c_long_long), value :: epoch; integer(c_int) :: ierr
    end function
  end interface
contains

  subroutine shock_source__init(n0, v0, v_thi, v_the, b0, theta_bn, phi_bn, l_damp_ini, seed)
    integer, intent(in)    :: n0
    real(8), intent(in)    :: v0, v_thi, v_the, b0, theta_bn, phi_bn, l_damp_ini
    integer(8), intent(in) :: seed
    sprm%n0 = n0; sprm%v0 = v0; sprm%v_thi = v_thi; sprm%v_the = v_the; sprm%b0 = b0
    sprm%theta_bn = theta_bn; sprm%phi_bn = phi_bn; sprm%l_damp_ini = l_damp_ini; sprm%seed = seed
  end subroutine shock_source__init

  subroutine shock_source__inject(nxe, nlinj_grid, id_first, it)          ! app.f90:786-849
    integer, intent(in)            :: nxe, it
    integer, intent(in), target    :: nlinj_grid(*)                         ! (nys:nye[,nzs:nze])
    integer(8), intent(in), target :: id_first(*)                           ! (nys:nye[,nzs:nze],nsp)
    integer(c_int) :: ierr
    ierr = wm_shock_inject(ctx, sprm, int(nxe,c_int), c_loc(nlinj_grid), c_loc(id_first), int(it,c_long_long))
    call wm_check(ierr, 'inject')
  end subroutine shock_source__inject

  subroutine shock_source__relocate(nxe_new, id_first, it)                 ! app.f90:637-689
    integer, intent(in)            :: nxe_new, it
    integer(8), intent(in), target :: id_first(*)
    integer(c_int) :: ierr
    ierr = wm_shock_relocate(ctx, sprm, int(nxe_new,c_int), c_loc(id_first), int(it,c_long_long))
    call wm_check(ierr, 'relocate')
  end subroutine shock_source__relocate

end module shock_source
