!> ISO_C_BINDING interface of libfv3tracer.so (include/fv3tracer.h) and the drop-in bodies that keep the reference's
!! argument lists.  Written against NOAA-EMC/fv3atm -> atmos_cubed_sphere:
!!   tracer_2d      model/fv_tracer2d.F90:324-345   (callers: model/fv_dynamics.F90:696-702)
!!   mapn_tracer    model/fv_mapz.F90:1386-1402     (caller : model/fv_mapz.F90:411-412, inside the OpenMP j loop :250-261)
!! Precision: `real` is 8 bytes in the 64-bit build (-fdefault-real-8) and 4 bytes with -D32BIT; the preprocessor
!! symbol OVERLOAD_R4 that the reference's CMake defines for the 32-bit build selects the fv3t_f32_* symbols.
!! NOT compiled in this repository's image (no Fortran compiler); see INTEGRATION.md.
module fv3tracer_iface_mod
  use iso_c_binding
  implicit none
  private
  public :: fv3t_dims, fv3t_grid, fv3t_create, fv3t_tracer_2d, fv3t_remap_tracers, fv3t_mapn_tracer, fv3t_last_error_msg

#ifdef OVERLOAD_R4
  integer, parameter, public :: fv3t_real = c_float
#else
  integer, parameter, public :: fv3t_real = c_double
#endif

  type, bind(C) :: fv3t_dims
    integer(c_int) :: npx, npz, nq_max, ntiles
    integer(c_int) :: tile_id(6)
  end type fv3t_dims

  type, bind(C) :: fv3t_grid
    type(c_ptr) :: area, rarea, dx, dy, dxa, dya, sin_sg
  end type fv3t_grid

  interface
#ifdef OVERLOAD_R4
#define FV3T_SYM(name) 'fv3t_f32_'//name
    integer(c_int) function fv3t_create(ctx, dims, grid, device, stream) bind(C, name='fv3t_f32_create')
#else
    integer(c_int) function fv3t_create(ctx, dims, grid, device, stream) bind(C, name='fv3t_f64_create')
#endif
      import :: c_ptr, c_int, fv3t_dims, fv3t_grid
      type(c_ptr), intent(out) :: ctx
      type(fv3t_dims), intent(in) :: dims
      type(fv3t_grid), intent(in) :: grid
      integer(c_int), value :: device
      type(c_ptr), value :: stream
    end function fv3t_create

#ifdef OVERLOAD_R4
    integer(c_int) function fv3t_tracer_2d(ctx, q, dp1, mfx, mfy, cx, cy, nq, hord, q_split, nord_tr, trdm, lim_fac, &
                                           nsplt_out, ksplt_out) bind(C, name='fv3t_f32_tracer_2d')
#else
    integer(c_int) function fv3t_tracer_2d(ctx, q, dp1, mfx, mfy, cx, cy, nq, hord, q_split, nord_tr, trdm, lim_fac, &
                                           nsplt_out, ksplt_out) bind(C, name='fv3t_f64_tracer_2d')
#endif
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      real(fv3t_real), intent(inout) :: q(*), dp1(*), mfx(*), mfy(*), cx(*), cy(*)
      integer(c_int), value :: nq, hord, q_split, nord_tr
      real(fv3t_real), value :: trdm, lim_fac
      type(c_ptr), value :: nsplt_out, ksplt_out
    end function fv3t_tracer_2d

#ifdef OVERLOAD_R4
    integer(c_int) function fv3t_remap_tracers(ctx, pe, ak, bk, ptop, q, delp, nq, kord_tr, fill) &
        bind(C, name='fv3t_f32_remap_tracers')
#else
    integer(c_int) function fv3t_remap_tracers(ctx, pe, ak, bk, ptop, q, delp, nq, kord_tr, fill) &
        bind(C, name='fv3t_f64_remap_tracers')
#endif
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      real(fv3t_real), intent(in) :: pe(*), ak(*), bk(*)
      real(fv3t_real), value :: ptop
      real(fv3t_real), intent(inout) :: q(*), delp(*)
      integer(c_int), value :: nq
      integer(c_int), intent(in) :: kord_tr(*)
      integer(c_int), value :: fill
    end function fv3t_remap_tracers

#ifdef OVERLOAD_R4
    integer(c_int) function fv3t_mapn_tracer(ctx, nq, km, pe1, pe2, q1, dp2, kord, j, i1, i2, isd, ied, jsd, jed, q_min, fill) &
        bind(C, name='fv3t_f32_mapn_tracer')
#else
    integer(c_int) function fv3t_mapn_tracer(ctx, nq, km, pe1, pe2, q1, dp2, kord, j, i1, i2, isd, ied, jsd, jed, q_min, fill) &
        bind(C, name='fv3t_f64_mapn_tracer')
#endif
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      integer(c_int), value :: nq, km, j, i1, i2, isd, ied, jsd, jed, fill
      real(fv3t_real), intent(in) :: pe1(*), pe2(*), dp2(*)
      real(fv3t_real), intent(inout) :: q1(*)
      integer(c_int), intent(in) :: kord(*)
      real(fv3t_real), value :: q_min
    end function fv3t_mapn_tracer

    function fv3t_last_error() bind(C, name='fv3t_last_error') result(msg)
      import :: c_ptr
      type(c_ptr) :: msg
    end function fv3t_last_error
  end interface

contains

  !> C string of the library's last error -> Fortran string (for mpp_error(FATAL, ...))
  function fv3t_last_error_msg() result(s)
    character(len=512) :: s
    character(kind=c_char), pointer :: cs(:)
    integer :: n
    s = ' '
    call c_f_pointer(fv3t_last_error(), cs, [512])
    do n = 1, 512
      if (cs(n) == c_null_char) exit
      s(n:n) = cs(n)
    end do
  end function fv3t_last_error_msg

end module fv3tracer_iface_mod


!> bind(C) interfaces of the device-resident / building-block entries (include/fv3tracer.h), 64-bit symbols shown; the
!! 32-bit build binds the fv3t_f32_* names exactly as fv3tracer_iface_mod does.
module fv3tracer_blocks_mod
  use iso_c_binding
  use fv3tracer_iface_mod, only: fv3t_real
  implicit none
  integer(c_int), parameter :: FV3T_Q = 0, FV3T_DP1 = 1, FV3T_MFX = 2, FV3T_MFY = 3, FV3T_CX = 4, FV3T_CY = 5, FV3T_PE = 6, &
                               FV3T_DELP = 7
  interface
    integer(c_int) function fv3t_upload(ctx, field, host, nq) bind(C, name='fv3t_f64_upload')
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      integer(c_int), value :: field, nq
      real(fv3t_real), intent(in) :: host(*)
    end function
    integer(c_int) function fv3t_download(ctx, field, host, nq) bind(C, name='fv3t_f64_download')
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      integer(c_int), value :: field, nq
      real(fv3t_real), intent(inout) :: host(*)
    end function
    integer(c_int) function fv3t_set_vertical(ctx, ak, bk, ptop) bind(C, name='fv3t_f64_set_vertical')
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      real(fv3t_real), intent(in) :: ak(*), bk(*)
      real(fv3t_real), value :: ptop
    end function
    integer(c_int) function fv3t_tracer_2d_begin(ctx, nq, q_split, cmax_local) bind(C, name='fv3t_f64_tracer_2d_begin')
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      integer(c_int), value :: nq, q_split
      real(fv3t_real), intent(out) :: cmax_local(*)
    end function
    integer(c_int) function fv3t_tracer_2d_set_cmax(ctx, cmax_global, q_split, nsplt_out) bind(C, name='fv3t_f64_tracer_2d_set_cmax')
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      real(fv3t_real), intent(in) :: cmax_global(*)
      integer(c_int), value :: q_split
      integer(c_int), intent(out) :: nsplt_out
    end function
    integer(c_int) function fv3t_halo_pack(ctx, it, local_tile, edge, dev_buf) bind(C, name='fv3t_f64_halo_pack')
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx, dev_buf
      integer(c_int), value :: it, local_tile, edge
    end function
    integer(c_int) function fv3t_halo_unpack(ctx, it, local_tile, edge, dev_buf) bind(C, name='fv3t_f64_halo_unpack')
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx, dev_buf
      integer(c_int), value :: it, local_tile, edge
    end function
    integer(c_int) function fv3t_tracer_2d_substep(ctx, it, hord, lim_fac) bind(C, name='fv3t_f64_tracer_2d_substep')
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      integer(c_int), value :: it, hord
      real(fv3t_real), value :: lim_fac
    end function
    integer(c_int) function fv3t_tracer_2d_finish(ctx) bind(C, name='fv3t_f64_tracer_2d_finish')
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx
    end function
    integer(c_int) function fv3t_remap_prepare(ctx) bind(C, name='fv3t_f64_remap_prepare')
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx
    end function
    integer(c_int) function fv3t_remap_tracers_resident(ctx, nq, kord_tr, fill) bind(C, name='fv3t_f64_remap_tracers_resident')
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx
      integer(c_int), value :: nq, fill
      integer(c_int), intent(in) :: kord_tr(*)
    end function
    integer(c_int) function fv3t_neighbor(ctx, global_tile, edge, nbr_tile, nbr_edge, rotated) bind(C, name='fv3t_neighbor')
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx
      integer(c_int), value :: global_tile, edge
      integer(c_int), intent(out) :: nbr_tile, nbr_edge, rotated
    end function
    integer(c_size_t) function fv3t_halo_strip_elems(ctx) bind(C, name='fv3t_halo_strip_elems')
      import :: c_ptr, c_size_t
      type(c_ptr), value :: ctx
    end function
  end interface
end module fv3tracer_blocks_mod


!> Drop-in body for fv_tracer2d_mod::tracer_2d (same dummy list as model/fv_tracer2d.F90:324-345) for the usual
!! decomposition of one MPI rank per tile (layout = 1,1); finer layouts need sub-tile contexts (SURVEY.md 8e, next round).
!! The sub-step loop of the reference (:496-566) is kept on the host so that the q halo is exchanged where the reference
!! completes it (:499): the 3-cell edge strips are packed on the device, already rotated into the neighbour's index
!! order (fv3t_*_halo_pack), exchanged with MPI (device pointers with CUDA-aware MPI, or staged through fv3t_strip_to_host),
!! and scattered by fv3t_*_halo_unpack.  cmax is reduced with mp_reduce_max exactly as at :433.
subroutine tracer_2d_b200(q, dp1, mfx, mfy, cx, cy, gridstruct, bd, domain, npx, npy, npz,   &
                          nq,  hord, q_split, dt, id_divg, q_pack, dp1_pack, nord_tr, trdm, lim_fac)
  use iso_c_binding
  use fv3tracer_iface_mod
  use fv3tracer_blocks_mod   ! bind(C) interfaces of fv3t_*_upload/download/tracer_2d_begin/set_cmax/halo_pack/halo_unpack/
                             ! tracer_2d_substep/tracer_2d_finish/fv3t_neighbor/fv3t_halo_strip_elems (same pattern as above)
  use fv_arrays_mod,   only: fv_grid_type, fv_grid_bounds_type
  use fv_mp_mod,       only: mp_reduce_max
  use mpp_domains_mod, only: domain2d, group_halo_update_type
  use mpp_mod,         only: mpp_error, FATAL
  implicit none
  type(fv_grid_bounds_type), intent(IN) :: bd
  integer, intent(IN) :: npx, npy, npz, nq, hord, nord_tr, q_split, id_divg
  real   , intent(IN) :: dt, trdm, lim_fac
  type(group_halo_update_type), intent(inout) :: q_pack, dp1_pack
  real   , intent(INOUT) :: q(bd%isd:bd%ied,bd%jsd:bd%jed,npz,nq)
  real   , intent(INOUT) :: dp1(bd%isd:bd%ied,bd%jsd:bd%jed,npz)
  real   , intent(INOUT) :: mfx(bd%is:bd%ie+1,bd%js:bd%je,  npz)
  real   , intent(INOUT) :: mfy(bd%is:bd%ie  ,bd%js:bd%je+1,npz)
  real   , intent(INOUT) ::  cx(bd%is:bd%ie+1,bd%jsd:bd%jed  ,npz)
  real   , intent(INOUT) ::  cy(bd%isd:bd%ied,bd%js :bd%je +1,npz)
  type(fv_grid_type), intent(IN), target :: gridstruct
  type(domain2d), intent(INOUT) :: domain

  type(c_ptr), save :: ctx = c_null_ptr
  type(fv3t_dims) :: dims
  type(fv3t_grid) :: g
  real, allocatable, target, save :: sin_sg5(:,:,:)
  real :: cmax(npz)
  integer(c_int) :: rc, nsplt, it

  if (trdm > 1.e-4) call mpp_error(FATAL, 'tracer_2d: tracer damping (trdm2 > 1e-4) is not on the accelerated path')
  if (.not. c_associated(ctx)) then
    dims%npx = npx; dims%npz = npz; dims%nq_max = nq; dims%ntiles = 1
    dims%tile_id = 0; dims%tile_id(1) = this_rank_tile(domain)   ! 1..6, from mpp_get_tile_id
    allocate(sin_sg5(bd%isd:bd%ied, bd%jsd:bd%jed, 5))
    sin_sg5 = gridstruct%sin_sg(:,:,1:5)                         ! the path reads sub-cell positions 1..5 only
    g%area  = c_loc(gridstruct%area);  g%rarea = c_loc(gridstruct%rarea)
    g%dx    = c_loc(gridstruct%dx);    g%dy    = c_loc(gridstruct%dy)
    g%dxa   = c_loc(gridstruct%dxa);   g%dya   = c_loc(gridstruct%dya)
    g%sin_sg = c_loc(sin_sg5)
    call check(fv3t_create(ctx, dims, g, local_gpu(), c_null_ptr))
  end if
  ! a host that keeps q, dp1, mfx, mfy, cx, cy resident on the device between dyn_core and here skips these six uploads
  call check(fv3t_upload(ctx, FV3T_Q, q, nq));     call check(fv3t_upload(ctx, FV3T_DP1, dp1, nq))
  call check(fv3t_upload(ctx, FV3T_MFX, mfx, nq)); call check(fv3t_upload(ctx, FV3T_MFY, mfy, nq))
  call check(fv3t_upload(ctx, FV3T_CX, cx, nq));   call check(fv3t_upload(ctx, FV3T_CY, cy, nq))
  call check(fv3t_tracer_2d_begin(ctx, nq, q_split, cmax))      ! fv_tracer2d.F90:387-427
  if (q_split == 0) call mp_reduce_max(cmax, npz)               ! :433
  call check(fv3t_tracer_2d_set_cmax(ctx, cmax, q_split, nsplt)) ! :432-486
  do it = 1, nsplt
    call exchange_q_halo(ctx, it, domain)                       ! :499  (pack -> MPI -> unpack, see header comment)
    call check(fv3t_tracer_2d_substep(ctx, it, hord, lim_fac))  ! :503-556
  end do
  call check(fv3t_tracer_2d_finish(ctx))
  call check(fv3t_download(ctx, FV3T_Q, q, nq));   call check(fv3t_download(ctx, FV3T_DP1, dp1, nq))
  if (nsplt /= 1) then                                          ! the caller sees the 1/ksplt-scaled arrays (:463-481)
    call check(fv3t_download(ctx, FV3T_MFX, mfx, nq)); call check(fv3t_download(ctx, FV3T_MFY, mfy, nq))
    call check(fv3t_download(ctx, FV3T_CX, cx, nq));   call check(fv3t_download(ctx, FV3T_CY, cy, nq))
  end if
contains
  subroutine check(rc)
    integer(c_int), intent(in) :: rc
    if (rc /= 0) call mpp_error(FATAL, 'tracer_2d: '//trim(fv3t_last_error_msg()))
  end subroutine check
end subroutine tracer_2d_b200
