!> Drop-in bodies that keep the reference's dummy argument lists and forward to libfv3tracer.so:
!!   tracer_2d      ACS/model/fv_tracer2d.F90:324-345   (caller ACS/model/fv_dynamics.F90:700-702)
!!   tracer_2d_1L   ACS/model/fv_tracer2d.F90:92-113    (caller fv_dynamics.F90:696-698, z_tracer)
!!   mapn_tracer    ACS/model/fv_mapz.F90:1386-1402     (caller fv_mapz.F90:411-412, inside the OpenMP j loop :250-261)
!! for the decomposition with ONE MPI rank per cubed-sphere tile (layout = 1,1: bd%is = bd%js = 1, bd%ie = bd%je = npx-1), the
!! only decomposition the library's contexts cover (include/fv3tracer.h; finer layouts are SURVEY.md 8e "sub-tile contexts").
!! Everything it calls exists: the bind(C) interfaces are in fv3tracer_iface.F90, the FMS calls are the ones the reference itself
!! uses (mpp_pe, mpp_root_pe, mpp_send, mpp_recv, mpp_sync_self: fv_mp_mod.F90; mp_reduce_max: fv_mp_mod.F90:1948).
!! NOT compiled in this repository's image (no Fortran compiler, no FMS / MPI): see INTEGRATION.md.
module fv3tracer_dropin_mod
  use iso_c_binding
  use fv3tracer_iface_mod
  use fv_arrays_mod,   only: fv_grid_type, fv_grid_bounds_type
  use fv_mp_mod,       only: mp_reduce_max
  use mpp_domains_mod, only: domain2d, group_halo_update_type, mpp_get_tile_id
  use mpp_mod,         only: mpp_error, FATAL, mpp_pe, mpp_root_pe, mpp_send, mpp_recv, mpp_sync_self
  implicit none
  private
  public :: tracer_2d_b200, tracer_2d_1L_b200, mapn_tracer_b200

  type(c_ptr), save :: ctx = c_null_ptr          ! one context per MPI rank / GPU, created on first use
  integer, save :: my_tile = 0
  real, allocatable, target, save :: sin_sg5(:,:,:)

contains

  subroutine check(rc, who)
    integer(c_int), intent(in) :: rc
    character(len=*), intent(in) :: who
    if (rc /= 0) call mpp_error(FATAL, who//': '//trim(fv3t_last_error_msg()))   ! the reference has no status returns
  end subroutine check

  !> Context of this rank's tile: uploads the fv_grid_type members the path reads (fv_arrays.F90:81-88,157).
  subroutine ensure_context(gridstruct, bd, domain, npx, npz, nq)
    type(fv_grid_type), intent(in), target :: gridstruct
    type(fv_grid_bounds_type), intent(in) :: bd
    type(domain2d), intent(inout) :: domain
    integer, intent(in) :: npx, npz, nq
    type(fv3t_dims) :: dims
    type(fv3t_grid) :: g
    integer :: tiles(1), ngpu
    if (c_associated(ctx)) return
    if (bd%is /= 1 .or. bd%js /= 1 .or. bd%ie /= npx-1 .or. bd%je /= npx-1) &
      call mpp_error(FATAL, 'fv3tracer: the accelerated path needs layout = 1,1 (one rank per tile)')
    tiles = mpp_get_tile_id(domain)
    my_tile = tiles(1)
    dims%npx = npx; dims%npz = npz; dims%nq_max = nq; dims%ntiles = 1
    dims%tile_id = 0; dims%tile_id(1) = my_tile
    allocate(sin_sg5(bd%isd:bd%ied, bd%jsd:bd%jed, 5))
    sin_sg5 = gridstruct%sin_sg(:,:,1:5)                          ! the path reads sub-cell positions 1..5 only
    g%area  = c_loc(gridstruct%area);  g%rarea = c_loc(gridstruct%rarea)
    g%dx    = c_loc(gridstruct%dx);    g%dy    = c_loc(gridstruct%dy)
    g%dxa   = c_loc(gridstruct%dxa);   g%dya   = c_loc(gridstruct%dya)
    g%sin_sg = c_loc(sin_sg5)
    ngpu = max(1, fv3t_device_count())
    call check(fv3t_create(ctx, dims, g, mod(mpp_pe() - mpp_root_pe(), ngpu), c_null_ptr), 'fv3t_create')
  end subroutine ensure_context

  !> The q halo update the reference completes at fv_tracer2d.F90:499 / :314: for each of the four tile edges the 3-cell strip
  !! the neighbour needs is packed on the device ALREADY ROTATED into the neighbour's index order (contact table,
  !! fv_mp_mod.F90:581-629), sent, and the strip received from that neighbour is scattered into this tile's halo.  With one
  !! rank per tile and the pe list in tile order, the rank of tile t is mpp_root_pe() + t - 1.
  subroutine exchange_q_halo(it)
    integer, intent(in) :: it
    real(fv3t_real), allocatable :: sbuf(:,:), rbuf(:,:)
    integer(c_int) :: nbr_tile, nbr_edge, rotated
    integer :: e, n, peer(0:3)
    n = int(fv3t_halo_strip_elems(ctx))
    allocate(sbuf(n,0:3), rbuf(n,0:3))
    do e = 0, 3
      call check(fv3t_neighbor(ctx, int(my_tile, c_int), int(e, c_int), nbr_tile, nbr_edge, rotated), 'fv3t_neighbor')
      peer(e) = mpp_root_pe() + nbr_tile - 1
      call check(fv3t_halo_pack_host(ctx, int(it, c_int), 0_c_int, int(e, c_int), sbuf(:,e)), 'halo_pack')
      call mpp_send(sbuf(:,e), n, peer(e), tag=100+int(nbr_edge))   ! tagged with the RECEIVER's edge
    end do
    do e = 0, 3
      call mpp_recv(rbuf(:,e), n, peer(e), tag=100+e)
      call check(fv3t_halo_unpack_host(ctx, int(it, c_int), 0_c_int, int(e, c_int), rbuf(:,e)), 'halo_unpack')
    end do
    call mpp_sync_self()
    deallocate(sbuf, rbuf)
  end subroutine exchange_q_halo

  subroutine tracer_2d_any(q, dp1, mfx, mfy, cx, cy, gridstruct, bd, domain, npx, npz, nq, hord, q_split, trdm, lim_fac, one_level)
    type(fv_grid_bounds_type), intent(IN) :: bd
    integer, intent(IN) :: npx, npz, nq, hord, q_split
    real   , intent(IN) :: trdm, lim_fac
    logical, intent(IN) :: one_level
    real   , intent(INOUT) :: q(bd%isd:bd%ied,bd%jsd:bd%jed,npz,nq)
    real   , intent(INOUT) :: dp1(bd%isd:bd%ied,bd%jsd:bd%jed,npz)
    real   , intent(INOUT) :: mfx(bd%is:bd%ie+1,bd%js:bd%je,  npz)
    real   , intent(INOUT) :: mfy(bd%is:bd%ie  ,bd%js:bd%je+1,npz)
    real   , intent(INOUT) ::  cx(bd%is:bd%ie+1,bd%jsd:bd%jed  ,npz)
    real   , intent(INOUT) ::  cy(bd%isd:bd%ied,bd%js :bd%je +1,npz)
    type(fv_grid_type), intent(IN), target :: gridstruct
    type(domain2d), intent(INOUT) :: domain
    real :: cmax(npz)
    integer(c_int) :: nsplt, it
    if (trdm > 1.e-4) call mpp_error(FATAL, 'tracer_2d: tracer damping (trdm2 > 1e-4, deln_flux) is not on the accelerated path')
    call ensure_context(gridstruct, bd, domain, npx, npz, nq)
    ! a host that keeps these arrays resident on the device between dyn_core and here skips the six uploads
    call check(fv3t_upload(ctx, FV3T_Q, q, int(nq, c_int)), 'upload');     call check(fv3t_upload(ctx, FV3T_DP1, dp1, int(nq, c_int)), 'upload')
    call check(fv3t_upload(ctx, FV3T_MFX, mfx, int(nq, c_int)), 'upload'); call check(fv3t_upload(ctx, FV3T_MFY, mfy, int(nq, c_int)), 'upload')
    call check(fv3t_upload(ctx, FV3T_CX, cx, int(nq, c_int)), 'upload');   call check(fv3t_upload(ctx, FV3T_CY, cy, int(nq, c_int)), 'upload')
    call check(fv3t_tracer_2d_begin(ctx, int(nq, c_int), int(q_split, c_int), cmax), 'tracer_2d_begin')   ! fv_tracer2d.F90:387-427
    if (q_split == 0) call mp_reduce_max(cmax, npz)                                                       ! :433 / :225
    call check(fv3t_tracer_2d_set_cmax(ctx, cmax, int(q_split, c_int), nsplt), 'tracer_2d_set_cmax')      ! :432-486
    do it = 1, nsplt
      call exchange_q_halo(int(it))                                                                       ! :499 / :314
      call check(fv3t_tracer_2d_substep(ctx, it, int(hord, c_int), real(lim_fac, fv3t_real)), 'tracer_2d_substep')  ! :503-556
    end do
    call check(fv3t_tracer_2d_finish(ctx), 'tracer_2d_finish')
    call check(fv3t_download(ctx, FV3T_Q, q, int(nq, c_int)), 'download')
    call check(fv3t_download(ctx, FV3T_DP1, dp1, int(nq, c_int)), 'download')
    if (nsplt /= 1) then                                             ! the caller sees the 1/ksplt-scaled arrays (:463-481)
      call check(fv3t_download(ctx, FV3T_MFX, mfx, int(nq, c_int)), 'download'); call check(fv3t_download(ctx, FV3T_MFY, mfy, int(nq, c_int)), 'download')
      call check(fv3t_download(ctx, FV3T_CX, cx, int(nq, c_int)), 'download');   call check(fv3t_download(ctx, FV3T_CY, cy, int(nq, c_int)), 'download')
    end if
    ! NOTE one_level: the building blocks advance dp1 as tracer_2d does; tracer_2d_1L's dp1 scratch differs for levels with
    ! ksplt(k) < nsplt (fv_tracer2d.F90:305 vs :547).  dp1 is not read again after the tracer advection (fv_dynamics.F90:704ff),
    ! and a single-context host gets the 1L post-state from fv3t_tracer_2d_1l.
    if (one_level) continue
  end subroutine tracer_2d_any

  subroutine tracer_2d_b200(q, dp1, mfx, mfy, cx, cy, gridstruct, bd, domain, npx, npy, npz,   &
                            nq,  hord, q_split, dt, id_divg, q_pack, dp1_pack, nord_tr, trdm, lim_fac)
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
    call tracer_2d_any(q, dp1, mfx, mfy, cx, cy, gridstruct, bd, domain, npx, npz, nq, hord, q_split, trdm, lim_fac, .false.)
  end subroutine tracer_2d_b200

  subroutine tracer_2d_1L_b200(q, dp1, mfx, mfy, cx, cy, gridstruct, bd, domain, npx, npy, npz,   &
                               nq,  hord, q_split, dt, id_divg, q_pack, dp1_pack, nord_tr, trdm, lim_fac)
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
    call tracer_2d_any(q, dp1, mfx, mfy, cx, cy, gridstruct, bd, domain, npx, npz, nq, hord, 0, trdm, lim_fac, .true.)
  end subroutine tracer_2d_1L_b200

  !> mapn_tracer with the reference's own dummy list (fv_mapz.F90:1386-1402).  Called from the OpenMP loop over j of
  !! Lagrangian_to_Eulerian with thread-private pe1, pe2, dp2: the library entry is thread-safe (serialised internally) and
  !! consumes pe2 / dp2 as given.  A host that hoists the j loop calls fv3t_remap_tracers once instead (INTEGRATION.md).
  subroutine mapn_tracer_b200(nq, km, pe1, pe2, q1, dp2, kord, j, i1, i2, isd, ied, jsd, jed, q_min, fill)
    integer, intent(in):: km, nq, j, i1, i2, isd, ied, jsd, jed
    integer, intent(in):: kord(nq)
    real, intent(in):: pe1(i1:i2,km+1), pe2(i1:i2,km+1), dp2(i1:i2,km), q_min
    logical, intent(in):: fill
    real, intent(inout):: q1(isd:ied,jsd:jed,km,nq)
    integer(c_int) :: kord_c(nq)
    if (.not. c_associated(ctx)) call mpp_error(FATAL, 'mapn_tracer: no fv3tracer context (tracer_2d creates it)')
    kord_c = int(kord, c_int)
    call check(fv3t_mapn_tracer(ctx, int(nq, c_int), int(km, c_int), pe1, pe2, q1, dp2, kord_c, int(j, c_int), int(i1, c_int), &
                                int(i2, c_int), int(isd, c_int), int(ied, c_int), int(jsd, c_int), int(jed, c_int),              &
                                real(q_min, fv3t_real), merge(1_c_int, 0_c_int, fill)), 'mapn_tracer')
  end subroutine mapn_tracer_b200

end module fv3tracer_dropin_mod
