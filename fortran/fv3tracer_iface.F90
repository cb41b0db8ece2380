!> ISO_C_BINDING interface of libfv3tracer.so (include/fv3tracer.h): one module, both precisions.
!! `real` is 8 bytes in the 64-bit build (-fdefault-real-8) and 4 bytes with -D32BIT; the preprocessor symbol OVERLOAD_R4
!! that the reference's CMake defines for the 32-bit build (ACS/CMakeLists.txt:27) selects the fv3t_f32_* symbol set -- for
!! EVERY entry, through the single prefix constant below (bind names are constant expressions: FV3T_P//'create').
!! NOT compiled in this repository's image (no Fortran compiler of any kind); see INTEGRATION.md for the build line.
module fv3tracer_iface_mod
  use iso_c_binding
  implicit none
  private
  public :: fv3t_dims, fv3t_grid, fv3t_real, fv3t_last_error_msg, fv3t_device_count, fv3t_neighbor, fv3t_halo_strip_elems
  public :: fv3t_create, fv3t_destroy, fv3t_tracer_2d, fv3t_tracer_2d_1l, fv3t_remap_tracers, fv3t_mapn_tracer, fv3t_tracer_step
  public :: fv3t_upload, fv3t_download, fv3t_set_vertical, fv3t_tracer_2d_begin, fv3t_tracer_2d_set_cmax, fv3t_halo_local
  public :: fv3t_halo_pack_host, fv3t_halo_unpack_host, fv3t_tracer_2d_substep, fv3t_tracer_2d_finish
  public :: fv3t_tracer_2d_resident, fv3t_remap_tracers_resident, fv3t_remap_prepare
  public :: fv3t_set_damping, fv3t_map_scalar, fv3t_map1_ppm, fv3t_map_field, fv3t_fv_tp_2d
  public :: fv3t_halo_list_create, fv3t_halo_list_count, fv3t_halo_local_table, fv3t_halo_gather, fv3t_halo_scatter
  public :: FV3T_Q, FV3T_DP1, FV3T_MFX, FV3T_MFY, FV3T_CX, FV3T_CY, FV3T_PE, FV3T_DELP

#ifdef OVERLOAD_R4
  integer, parameter :: fv3t_real = c_float
  character(len=*), parameter :: FV3T_P = 'fv3t_f32_'
#else
  integer, parameter :: fv3t_real = c_double
  character(len=*), parameter :: FV3T_P = 'fv3t_f64_'
#endif
  integer(c_int), parameter :: FV3T_Q = 0, FV3T_DP1 = 1, FV3T_MFX = 2, FV3T_MFY = 3, FV3T_CX = 4, FV3T_CY = 5, FV3T_PE = 6, &
                               FV3T_DELP = 7

  type, bind(C) :: fv3t_dims
    integer(c_int) :: npx, npz, nq_max, ntiles
    integer(c_int) :: tile_id(6)
    !> sub-tile contexts (layout > 1,1: bd%is..ie a proper sub-range of the tile): sub_layout = L, block (sub_bi, sub_bj) of every
    !! resident sub-domain, npx = LOCAL extent + 1; 0 = whole tiles
    integer(c_int) :: sub_layout = 0
    integer(c_int) :: sub_bi(6) = 0, sub_bj(6) = 0
  end type fv3t_dims

  type, bind(C) :: fv3t_grid
    type(c_ptr) :: area, rarea, dx, dy, dxa, dya, sin_sg
  end type fv3t_grid

  interface
    integer(c_int) function fv3t_create(ctx, dims, grid, device, stream) bind(C, name=FV3T_P//'create')
      import :: c_ptr, c_int, fv3t_dims, fv3t_grid
      type(c_ptr), intent(out) :: ctx
      type(fv3t_dims), intent(in) :: dims
      type(fv3t_grid), intent(in) :: grid
      integer(c_int), value :: device
      type(c_ptr), value :: stream
    end function
    integer(c_int) function fv3t_destroy(ctx) bind(C, name='fv3t_destroy')
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx
    end function
    integer(c_int) function fv3t_device_count() bind(C, name='fv3t_device_count')
      import :: c_int
    end function
    integer(c_int) function fv3t_tracer_2d(ctx, q, dp1, mfx, mfy, cx, cy, nq, hord, q_split, nord_tr, trdm, lim_fac, &
                                           nsplt_out, ksplt_out) bind(C, name=FV3T_P//'tracer_2d')
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      real(fv3t_real), intent(inout) :: q(*), dp1(*), mfx(*), mfy(*), cx(*), cy(*)
      integer(c_int), value :: nq, hord, q_split, nord_tr
      real(fv3t_real), value :: trdm, lim_fac
      type(c_ptr), value :: nsplt_out, ksplt_out
    end function
    integer(c_int) function fv3t_tracer_2d_1l(ctx, q, dp1, mfx, mfy, cx, cy, nq, hord, q_split, nord_tr, trdm, lim_fac, &
                                              nsplt_out, ksplt_out) bind(C, name=FV3T_P//'tracer_2d_1L')
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      real(fv3t_real), intent(inout) :: q(*), dp1(*), mfx(*), mfy(*), cx(*), cy(*)
      integer(c_int), value :: nq, hord, q_split, nord_tr
      real(fv3t_real), value :: trdm, lim_fac
      type(c_ptr), value :: nsplt_out, ksplt_out
    end function
    integer(c_int) function fv3t_remap_tracers(ctx, pe, ak, bk, ptop, q, delp, nq, kord_tr, fill) bind(C, name=FV3T_P//'remap_tracers')
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      real(fv3t_real), intent(in) :: pe(*), ak(*), bk(*)
      real(fv3t_real), value :: ptop
      real(fv3t_real), intent(inout) :: q(*), delp(*)
      integer(c_int), value :: nq, fill
      integer(c_int), intent(in) :: kord_tr(*)
    end function
    integer(c_int) function fv3t_mapn_tracer(ctx, nq, km, pe1, pe2, q1, dp2, kord, j, i1, i2, isd, ied, jsd, jed, q_min, fill) &
        bind(C, name=FV3T_P//'mapn_tracer')
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      integer(c_int), value :: nq, km, j, i1, i2, isd, ied, jsd, jed, fill
      real(fv3t_real), intent(in) :: pe1(*), pe2(*), dp2(*)
      real(fv3t_real), intent(inout) :: q1(*)
      integer(c_int), intent(in) :: kord(*)
      real(fv3t_real), value :: q_min
    end function
    integer(c_int) function fv3t_tracer_step(ctx, q, dp1, mfx, mfy, cx, cy, pe, ak, bk, ptop, delp, nq, hord, q_split, lim_fac, &
                                             kord_tr, fill, nsplt_out) bind(C, name=FV3T_P//'tracer_step')
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      real(fv3t_real), intent(inout) :: q(*), dp1(*), mfx(*), mfy(*), cx(*), cy(*), delp(*)
      real(fv3t_real), intent(in) :: pe(*), ak(*), bk(*)
      real(fv3t_real), value :: ptop, lim_fac
      integer(c_int), value :: nq, hord, q_split, fill
      integer(c_int), intent(in) :: kord_tr(*)
      type(c_ptr), value :: nsplt_out
    end function
    integer(c_int) function fv3t_upload(ctx, field, host, nq) bind(C, name=FV3T_P//'upload')
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      integer(c_int), value :: field, nq
      real(fv3t_real), intent(in) :: host(*)
    end function
    integer(c_int) function fv3t_download(ctx, field, host, nq) bind(C, name=FV3T_P//'download')
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      integer(c_int), value :: field, nq
      real(fv3t_real), intent(inout) :: host(*)
    end function
    integer(c_int) function fv3t_set_vertical(ctx, ak, bk, ptop) bind(C, name=FV3T_P//'set_vertical')
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      real(fv3t_real), intent(in) :: ak(*), bk(*)
      real(fv3t_real), value :: ptop
    end function
    integer(c_int) function fv3t_tracer_2d_begin(ctx, nq, q_split, cmax_local) bind(C, name=FV3T_P//'tracer_2d_begin')
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      integer(c_int), value :: nq, q_split
      real(fv3t_real), intent(out) :: cmax_local(*)
    end function
    integer(c_int) function fv3t_tracer_2d_set_cmax(ctx, cmax_global, q_split, nsplt_out) bind(C, name=FV3T_P//'tracer_2d_set_cmax')
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      real(fv3t_real), intent(in) :: cmax_global(*)
      integer(c_int), value :: q_split
      integer(c_int), intent(out) :: nsplt_out
    end function
    integer(c_int) function fv3t_halo_local(ctx, it) bind(C, name=FV3T_P//'halo_local')
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx
      integer(c_int), value :: it
    end function
    integer(c_int) function fv3t_halo_pack_host(ctx, it, local_tile, edge, host_buf) bind(C, name=FV3T_P//'halo_pack_host')
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      integer(c_int), value :: it, local_tile, edge
      real(fv3t_real), intent(out) :: host_buf(*)
    end function
    integer(c_int) function fv3t_halo_unpack_host(ctx, it, local_tile, edge, host_buf) bind(C, name=FV3T_P//'halo_unpack_host')
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      integer(c_int), value :: it, local_tile, edge
      real(fv3t_real), intent(in) :: host_buf(*)
    end function
    integer(c_int) function fv3t_tracer_2d_substep(ctx, it, hord, lim_fac) bind(C, name=FV3T_P//'tracer_2d_substep')
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      integer(c_int), value :: it, hord
      real(fv3t_real), value :: lim_fac
    end function
    integer(c_int) function fv3t_tracer_2d_finish(ctx) bind(C, name=FV3T_P//'tracer_2d_finish')
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx
    end function
    integer(c_int) function fv3t_tracer_2d_resident(ctx, nq, hord, q_split, lim_fac, nsplt_out) bind(C, name=FV3T_P//'tracer_2d_resident')
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      integer(c_int), value :: nq, hord, q_split
      real(fv3t_real), value :: lim_fac
      integer(c_int), intent(out) :: nsplt_out
    end function
    integer(c_int) function fv3t_remap_tracers_resident(ctx, nq, kord_tr, fill) bind(C, name=FV3T_P//'remap_tracers_resident')
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx
      integer(c_int), value :: nq, fill
      integer(c_int), intent(in) :: kord_tr(*)
    end function
    integer(c_int) function fv3t_remap_prepare(ctx) bind(C, name=FV3T_P//'remap_prepare')
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx
    end function
    integer(c_int) function fv3t_set_damping(ctx, del6_u, del6_v, da_min, nord_tr, trdm) bind(C, name=FV3T_P//'set_damping')
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      type(c_ptr), value :: del6_u, del6_v        ! c_loc of gridstruct%del6_u / del6_v, or c_null_ptr to keep the metrics
      real(fv3t_real), value :: da_min, trdm
      integer(c_int), value :: nord_tr
    end function
    integer(c_int) function fv3t_map_scalar(ctx, q, qs, iv, kord, q_min) bind(C, name=FV3T_P//'map_scalar')
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      real(fv3t_real), intent(inout) :: q(*)
      type(c_ptr), value :: qs                     ! c_loc(qs) for iv = -2, else c_null_ptr
      integer(c_int), value :: iv, kord
      real(fv3t_real), value :: q_min
    end function
    integer(c_int) function fv3t_map1_ppm(ctx, q, qs, iv, kord) bind(C, name=FV3T_P//'map1_ppm')
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      real(fv3t_real), intent(inout) :: q(*)
      type(c_ptr), value :: qs
      integer(c_int), value :: iv, kord
    end function
    integer(c_int) function fv3t_map_field(ctx, q, qs, iv, kord, q_min, use_cs) bind(C, name=FV3T_P//'map_field')
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      real(fv3t_real), intent(inout) :: q(*)
      type(c_ptr), value :: qs
      integer(c_int), value :: iv, kord, use_cs
      real(fv3t_real), value :: q_min
    end function
    !> fv_tp_2d (tp_core.F90:110-133); the Fortran optionals arrive as c_null_ptr / nord < 0
    integer(c_int) function fv3t_fv_tp_2d(ctx, nlev, q, crx, cry, hord, fx, fy, xfx, yfx, ra_x, ra_y, lim_fac, mfx, mfy, mass, nord, &
                                          damp_c) bind(C, name=FV3T_P//'fv_tp_2d')
      import :: c_ptr, c_int, fv3t_real
      type(c_ptr), value :: ctx
      integer(c_int), value :: nlev, hord, nord
      real(fv3t_real), intent(inout) :: q(*)
      real(fv3t_real), intent(in) :: crx(*), cry(*), xfx(*), yfx(*), ra_x(*), ra_y(*)
      real(fv3t_real), intent(out) :: fx(*), fy(*)
      real(fv3t_real), value :: lim_fac, damp_c
      type(c_ptr), value :: mfx, mfy, mass
    end function
    !> generic halo exchange by gather list (sub-tile contexts); dev_buf is a DEVICE address (CUDA-aware MPI / NCCL)
    integer(c_int) function fv3t_halo_list_create(ctx, offsets, count, list) bind(C, name='fv3t_halo_list_create')
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx
      integer(c_int), intent(in) :: offsets(*)
      integer(c_int), value :: count
      integer(c_int), intent(out) :: list
    end function
    integer(c_int) function fv3t_halo_list_count(ctx, list) bind(C, name='fv3t_halo_list_count')
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx
      integer(c_int), value :: list
    end function
    integer(c_int) function fv3t_halo_local_table(ctx, dst, src, len) bind(C, name='fv3t_halo_local_table')
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx
      integer(c_int), intent(in) :: dst(*), src(*)
      integer(c_int), value :: len
    end function
    integer(c_int) function fv3t_halo_gather(ctx, it, local_tile, list, dev_buf, buf_stride) bind(C, name=FV3T_P//'halo_gather')
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx, dev_buf
      integer(c_int), value :: it, local_tile, list, buf_stride
    end function
    integer(c_int) function fv3t_halo_scatter(ctx, it, local_tile, list, dev_buf, buf_stride) bind(C, name=FV3T_P//'halo_scatter')
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx, dev_buf
      integer(c_int), value :: it, local_tile, list, buf_stride
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
    function fv3t_last_error() bind(C, name='fv3t_last_error') result(msg)
      import :: c_ptr
      type(c_ptr) :: msg
    end function
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
