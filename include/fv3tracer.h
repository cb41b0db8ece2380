/* fv3tracer.h -- C-ABI of the B200-native FV3 tracer-transport path (libfv3tracer.so).
 *
 * Drop-in boundary for ONE hot path of NOAA-EMC/fv3atm (submodule atmos_cubed_sphere, "ACS/"):
 *   - sub-cycled horizontal tracer advection   ACS/model/fv_tracer2d.F90:324-569 (tracer_2d; tracer_2d_1L :92-321 is
 *     numerically identical for q_split = 0, trdm2 = 0) -> ACS/model/tp_core.F90:110-249 (fv_tp_2d), xppm/yppm/
 *     copy_corners/pert_ppm (:253-1236)
 *   - vertical Lagrangian-to-Eulerian tracer remap  ACS/model/fv_mapz.F90:261-273,343-368,407-426 (tracer part of
 *     Lagrangian_to_Eulerian) -> mapn_tracer (:1386-1499) / map1_q2 (:1502-1592), scalar_profile (:1691-2096),
 *     cs_limiters (:2501-2576), ppm_profile/ppm_limiters (:2580-2916), fillz (ACS/model/fv_fill.F90:51-156)
 *
 * Conventions
 *   - Plain C: raw pointers + sizes.  Every array has EXACTLY the Fortran layout of the reference's dummy argument
 *     (column-major, i contiguous) with the global-domain bounds is=js=1, ie=je=npx-1, ng=3:
 *         q   (isd:ied, jsd:jed, npz, nq)      dp1, delp (isd:ied, jsd:jed, npz)
 *         cx  (is:ie+1, jsd:jed, npz)          cy  (isd:ied, js:je+1, npz)
 *         mfx (is:ie+1, js:je,  npz)           mfy (is:ie,  js:je+1, npz)
 *         pe  (is-1:ie+1, npz+1, js-1:je+1)    ak, bk (npz+1)
 *     A context holds `ntiles` cubed-sphere tiles (all six for a whole mosaic on one GPU); multi-tile arguments are
 *     the per-tile Fortran arrays stacked tile-major (tile stride = size of one tile's array).
 *   - Two symbol sets, fv3t_f64_* and fv3t_f32_*, mirror the reference's 64-bit and 32-bit (-D32BIT) dycore builds
 *     (ACS/CMakeLists.txt:27); declared below by the FV3T_DECLARE macro with REAL = double / float.
 *   - All functions return 0 on success, non-zero on error (fv3t_last_error() gives the message).  The reference has
 *     no status returns (mpp_error(FATAL) aborts, fv_tracer2d.F90:76); the Fortran shim maps non-zero to that.
 *   - `stream` arguments are cudaStream_t passed as void* (NULL = default stream).  Host-array entry points are
 *     synchronous; *_resident entry points only enqueue work on the context's stream.
 *   - There is no CPU fallback: creating a context without a CUDA device fails.
 */
#ifndef FV3TRACER_H
#define FV3TRACER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fv3t_ctx fv3t_ctx;

/* fv_grid_bounds_type + the scalar members the path needs (ACS/model/fv_arrays.F90:1178-1186, :181-208) */
typedef struct fv3t_dims {
  int npx;        /* cells per tile edge + 1 (npx == npy on the cubed sphere)                   */
  int npz;        /* number of levels (<= 128 in this build)                                      */
  int nq_max;     /* capacity: largest nq any call will pass                                     */
  int ntiles;     /* tiles resident in this context, 1..6                                        */
  int tile_id[6]; /* global tile numbers (1..6) of the resident tiles, in storage order          */
  /* Sub-tile contexts -- a rank that owns PART of a tile (bd%is..ie a proper sub-range of 1..npx-1, layout > 1,1;
     fv_arrays.F90:1178-1186).  sub_layout = 0 or 1: whole tiles.  sub_layout = L >= 2: every resident "tile" t is the square
     sub-domain (sub_bi[t], sub_bj[t]) (0-based block column / row) of the L x L decomposition of tile tile_id[t]; npx is then
     the LOCAL extent + 1 (global npx = L * (npx - 1) + 1), all array shapes are the local ones, and the same tile may appear
     several times.  The tile-edge formulas of xppm / yppm are applied on the sides that lie on a tile edge only and
     copy_corners' views at true cube corners only (gridstruct%sw_corner .., fv_arrays.F90:181); every other halo cell --
     the diagonal blocks included -- must be delivered by the halo exchange (fv3t_halo_list_create, fv3t_*_halo_gather /
     _scatter, fv3t_halo_local_table).  Sub-tile contexts are driven through the tracer_2d building blocks and the remap. */
  int sub_layout;
  int sub_bi[6], sub_bj[6];
} fv3t_dims;

/* Device-resident fields addressable through fv3t_*_upload/download/device_ptr */
enum fv3t_field {
  FV3T_Q = 0,   /* tracers (two ping-pong buffers; outside a tracer_2d call every level lives in the current one) */
  FV3T_DP1 = 1, /* delp before dyn_core (updated in place between sub-steps, fv_tracer2d.F90:547-553)        */
  FV3T_MFX = 2,
  FV3T_MFY = 3,
  FV3T_CX = 4,
  FV3T_CY = 5,
  FV3T_PE = 6,  /* Lagrangian interface pressure, (i,k,j) order                                              */
  FV3T_DELP = 7 /* Eulerian delp written by the remap (fv_mapz.F90:364-368)                                   */
};

const char* fv3t_last_error(void);
int fv3t_device_count(void);

#define FV3T_DECLARE(P, REAL)                                                                                           \
  /* The fv_grid_type members the path reads (ACS/model/fv_arrays.F90:81-88,157): `real` copies, tile-major,         \
     extents area/rarea/dxa/dya (isd:ied,jsd:jed), dx (isd:ied,jsd:jed+1), dy (isd:ied+1,jsd:jed),                     \
     sin_sg (isd:ied,jsd:jed,5) = sub-cell positions 1..5 of sin_sg(:,:,9).  Replaces gridstruct in the                \
     reference's tracer_2d / fv_tp_2d argument lists. */                                                               \
  typedef struct fv3t_##P##_grid {                                                                                      \
    const REAL *area, *rarea, *dx, *dy, *dxa, *dya, *sin_sg;                                                            \
  } fv3t_##P##_grid;                                                                                                    \
                                                                                                                        \
  /* Allocate device mirrors and upload the grid.  Replaces nothing in the reference (state is Fortran-owned there);   \
     corresponds to the lifetime of fv_atmos_type (fv_arrays.F90:1386). */                                             \
  int fv3t_##P##_create(fv3t_ctx** ctx, const fv3t_dims* dims, const fv3t_##P##_grid* grid, int device, void* stream);  \
                                                                                                                        \
  /* tracer_2d(q, dp1, mfx, mfy, cx, cy, gridstruct, bd, domain, npx, npy, npz, nq, hord, q_split, dt, id_divg,        \
               q_pack, dp1_pack, nord_tr, trdm, lim_fac)            ACS/model/fv_tracer2d.F90:324-345                  \
     Host arrays in, host arrays out; reproduces the reference's side effects: q updated on the compute domain,       \
     cx, cy, mfx, mfy scaled by 1/ksplt(k) when nsplt /= 1 (:463-481), dp1 advanced between sub-steps (:549).         \
     The halo update of q that the reference completes at :499 is performed internally for the resident tiles          \
     (requires ntiles == 6).  nord_tr / trdm: tracer damping (deln_flux on the first sub-step, :487-494, 527-532) when     \
     trdm > 1e-4; needs the damping metrics of fv3t_*_set_damping, nord_tr in 0..2.                                    \
     nsplt_out / ksplt_out[npz] (optional) return the sub-step counts (:441,:457). */                                  \
  int fv3t_##P##_tracer_2d(fv3t_ctx* ctx, REAL* q, REAL* dp1, REAL* mfx, REAL* mfy, REAL* cx, REAL* cy, int nq,         \
                           int hord, int q_split, int nord_tr, REAL trdm, REAL lim_fac, int* nsplt_out, int* ksplt_out); \
                                                                                                                        \
  /* Tracer damping = deln_flux (ACS/model/tp_core.F90:1239-1387) as fv_tp_2d calls it for tracers (:229-234): the members of     \
     fv_grid_type it reads, del6_u (isd:ied, jsd:jed+1), del6_v (isd:ied+1, jsd:jed) (fv_arrays.F90:124) and da_min (:183),      \
     tile-major, plus the settings used by the device-resident entries (fv3t_*_tracer_2d takes nord_tr / trdm per call).          \
     del6_u = del6_v = NULL keeps the metrics and only changes the settings.  trdm <= 1e-4 switches the damping off.            \
     The damping fluxes are applied as a correction of the advected field: same terms as the reference, to rounding. */          \
  int fv3t_##P##_set_damping(fv3t_ctx* ctx, const REAL* del6_u, const REAL* del6_v, REAL da_min, int nord_tr, REAL trdm);   \
                                                                                                                        \
  /* tracer_2d_1L(q, dp1, mfx, mfy, cx, cy, gridstruct, bd, domain, npx, npy, npz, nq, hord, q_split, dt, id_divg, q_pack,        \
                  dp1_pack, nord_tr, trdm, lim_fac)                  ACS/model/fv_tracer2d.F90:92-113, called for z_tracer        \
     (fv_dynamics.F90:696-698).  Same arguments and the same q, cx, cy, mfx, mfy post-state as tracer_2d (the per-level         \
     sub-step counts of :201-202 are the ksplt(k) of tracer_2d); the dp1 post-state differs: a level's dp1 is advanced          \
     between ITS OWN sub-steps only (:305), i.e. ksplt(k)-1 times.  q_split is not used by the routine. */                      \
  int fv3t_##P##_tracer_2d_1L(fv3t_ctx* ctx, REAL* q, REAL* dp1, REAL* mfx, REAL* mfy, REAL* cx, REAL* cy, int nq,         \
                              int hord, int q_split, int nord_tr, REAL trdm, REAL lim_fac, int* nsplt_out, int* ksplt_out); \
                                                                                                                        \
  /* Tracer part of Lagrangian_to_Eulerian for all rows js..je at once (the j loop of fv_mapz.F90:261 hoisted):        \
     pe1 = pe(:,:,j); pe2 = ak + bk*pe(:,km+1,j); dp2; delp <- dp2 (:263-272,350-368); then mapn_tracer (nq > 5) or   \
     map1_q2 + fillz per tracer (:410-426).  q and delp are updated in place on the compute domain. */                 \
  int fv3t_##P##_remap_tracers(fv3t_ctx* ctx, const REAL* pe, const REAL* ak, const REAL* bk, REAL ptop, REAL* q,       \
                               REAL* delp, int nq, const int* kord_tr, int fill);                                       \
                                                                                                                        \
  /* mapn_tracer(nq, km, pe1, pe2, q1, dp2, kord, j, i1, i2, isd, ied, jsd, jed, q_min, fill)   fv_mapz.F90:1386-1402  \
     Row-granular compatibility entry with the reference's own argument list (one tile, one row j; pe1, pe2           \
     (i1:i2, km+1), dp2 (i1:i2, km), q1 (isd:ied, jsd:jed, km, nq)).  pe2 and dp2 are consumed AS GIVEN (any target      \
     grid, not only ak + bk*ps) and the profile is always scalar_profile, whatever nq -- it is mapn_tracer, not the    \
     nq-dependent dispatch of Lagrangian_to_Eulerian.  Thread-safe (the reference calls it from an OpenMP loop over j,  \
     fv_mapz.F90:250-261); serialised internally.  Prefer fv3t_*_remap_tracers. */                 \
  int fv3t_##P##_mapn_tracer(fv3t_ctx* ctx, int nq, int km, const REAL* pe1, const REAL* pe2, REAL* q1, const REAL* dp2, \
                             const int* kord, int j, int i1, int i2, int isd, int ied, int jsd, int jed, REAL q_min,   \
                             int fill);                                                                                 \
                                                                                                                        \
  /* map_scalar(km, pe1, q1, qs, kn, pe2, q2, i1, i2, j, ibeg, iend, jbeg, jend, iv, kord, q_min)  ACS/model/fv_mapz.F90:1199-1290  \
     map1_ppm  (km, pe1, q1, qs, kn, pe2, q2, i1, i2, j, ibeg, iend, jbeg, jend, iv, kord)         ACS/model/fv_mapz.F90:1293-1383  \
     for kn = km and all rows js..je of the resident tiles at once (the j loop of Lagrangian_to_Eulerian hoisted, as for the        \
     tracers): ONE scalar field q (isd:ied, jsd:jed, km) per tile is remapped in place from the resident Lagrangian pe              \
     (fv3t_*_upload(FV3T_PE)) onto pe2 = ak + bk*pe(km+1) (fv3t_*_set_vertical; fv_mapz.F90:263-272).  map_scalar builds the        \
     sub-grid profile with scalar_profile (:1691-2096), map1_ppm with cs_profile (:2098-2498) -- the routine through which the       \
     reference remaps pt, w, delz, u, v (:393-436, 610-660) -- and both with ppm_profile for kord <= 7.  iv: 0 positive definite,    \
     1 others, -1 winds, 2 (top layer flat), -2 (bottom interface value qs (isd:ied, jsd:jed) given; qs may be NULL otherwise).      \
     fv3t_*_map_field is the common entry (use_cs selects the profile).  The reference's own operation order, bit for bit. */        \
  int fv3t_##P##_map_scalar(fv3t_ctx* ctx, REAL* q, const REAL* qs, int iv, int kord, REAL q_min);                       \
  int fv3t_##P##_map1_ppm(fv3t_ctx* ctx, REAL* q, const REAL* qs, int iv, int kord);                                     \
  int fv3t_##P##_map_field(fv3t_ctx* ctx, REAL* q, const REAL* qs, int iv, int kord, REAL q_min, int use_cs);            \
                                                                                                                        \
  /* fv_tp_2d(q, crx, cry, npx, npy, hord, fx, fy, xfx, yfx, gridstruct, bd, ra_x, ra_y, lim_fac, mfx, mfy, mass, nord, damp_c)     \
     ACS/model/tp_core.F90:110-133 -- the 2-D transport operator as the dynamical core calls it directly (d_sw.F90 for delp, pt, w, \
     vorticity; fv_tracer2d.F90:510-531 for tracers), for `nlev` stacked 2-D fields per resident tile (tile-major, level fastest    \
     within a tile).  Shapes per field as in the reference for a rank that owns a whole tile: q (isd:ied, jsd:jed) INOUT -- its     \
     corner halos come back holding the dir = 1 copy_corners view, as in the reference (:189); crx, xfx (is:ie+1, jsd:jed);         \
     cry, yfx (isd:ied, js:je+1); ra_x (is:ie, jsd:jed); ra_y (isd:ied, js:je); fx (is:ie+1, js:je) and fy (is:ie, js:je+1) OUT.    \
     The Fortran optionals: mfx (is:ie+1, js:je) and mfy (is:ie, js:je+1) both given -> fluxes scaled by the mass fluxes (:209-220),\
     both NULL -> by xfx, yfx (:236-242); mass (isd:ied, jsd:jed) or NULL; nord < 0 = absent.  deln_flux (:1239-1387) is added when \
     damp_c > 1e-4 and nord is present (tracer branch: and mass is present), mass weighted in the tracer branch only; it needs      \
     fv3t_*_set_damping's metrics and nord <= 2.  The grid terms (area, dxa, dya, rarea, del6_u, del6_v, da_min) are the            \
     context's.  The reference's own operation order, bit for bit, for every hord. */                                              \
  int fv3t_##P##_fv_tp_2d(fv3t_ctx* ctx, int nlev, REAL* q, const REAL* crx, const REAL* cry, int hord, REAL* fx, REAL* fy,       \
                          const REAL* xfx, const REAL* yfx, const REAL* ra_x, const REAL* ra_y, REAL lim_fac, const REAL* mfx,     \
                          const REAL* mfy, const REAL* mass, int nord, REAL damp_c);                                              \
                                                                                                                        \
  /* tracer_2d immediately followed by the tracer remap (they are consecutive in ACS/model/fv_dynamics.F90:686-760), host    \
     arrays in and out, as ONE call: same arguments, same post-state as fv3t_*_tracer_2d + fv3t_*_remap_tracers (q, delp;     \
     dp1, cx, cy, mfx, mfy when nsplt /= 1).  Tracers are independent on this path, so they are pipelined one by one:         \
     upload of tracer i+1 | kernels of tracer i | download of tracer i-1, both PCIe directions busy at once.  Host arrays      \
     should be page-locked.  Needs ntiles == 6. */                                                                           \
  int fv3t_##P##_tracer_step(fv3t_ctx* ctx, REAL* q, REAL* dp1, REAL* mfx, REAL* mfy, REAL* cx, REAL* cy, const REAL* pe,   \
                             const REAL* ak, const REAL* bk, REAL ptop, REAL* delp, int nq, int hord, int q_split,          \
                             REAL lim_fac, const int* kord_tr, int fill, int* nsplt_out);                                    \
                                                                                                                        \
  /* Device-resident operation (north star: tracers, Courant numbers, mass fluxes and delp stay in HBM).  nq is the tile     \
     stride of the resident q: the resident calls must be given the count of the last upload(FV3T_Q) (anything else is an      \
     error); a host that fills q through fv3t_device_ptr lays it out as (isd:ied, jsd:jed, npz, nq) per tile itself. */        \
  int fv3t_##P##_upload(fv3t_ctx* ctx, int field, const REAL* host, int nq);                                            \
  int fv3t_##P##_download(fv3t_ctx* ctx, int field, REAL* host, int nq);                                                \
  int fv3t_##P##_set_vertical(fv3t_ctx* ctx, const REAL* ak, const REAL* bk, REAL ptop);                                \
  int fv3t_##P##_tracer_2d_resident(fv3t_ctx* ctx, int nq, int hord, int q_split, REAL lim_fac, int* nsplt_out);        \
  int fv3t_##P##_remap_tracers_resident(fv3t_ctx* ctx, int nq, const int* kord_tr, int fill);                           \
  /* Optional hint: the resident pe is final for this step (it is when dyn_core returns, before tracer_2d is called,    \
     ACS/model/fv_dynamics.F90:600-704), so the tracer-independent remap coefficients and delp may be computed on a     \
     side stream while tracer_2d runs.  Consumed by the next remap_tracers_resident; invalidated by uploading pe or      \
     set_vertical.  Do not write pe through fv3t_device_ptr between this call and the remap. */                          \
  int fv3t_##P##_remap_prepare(fv3t_ctx* ctx);                                                                           \
                                                                                                                        \
  /* Building blocks for a context that holds only some tiles (face sharding): the caller transports the packed       \
     edge strips between contexts (NCCL send/recv in this repo, MPI in a Fortran host) and reduces cmax.               \
       begin   : steps A-B of tracer_2d up to the local cmax(1:npz) (fv_tracer2d.F90:387-427), returned on the host    \
       set_cmax: the globally reduced cmax -> nsplt, ksplt(k) (:432-457)                                                \
       halo_pack / halo_unpack: pack the 3-cell edge strips each remote neighbour needs, already rotated into the      \
                 receiver's index order (mosaic contacts, ACS/tools/fv_mp_mod.F90:581-629) / scatter received strips  \
       substep : one pass of the `it` loop body (:503-556) for the resident tiles                                     \
       finish  : the in-place 1/ksplt scaling of cx, cy, mfx, mfy (:463-481) */                                        \
  int fv3t_##P##_tracer_2d_begin(fv3t_ctx* ctx, int nq, int q_split, REAL* cmax_local);                                 \
  int fv3t_##P##_tracer_2d_set_cmax(fv3t_ctx* ctx, const REAL* cmax_global, int q_split, int* nsplt_out);               \
  int fv3t_##P##_halo_local(fv3t_ctx* ctx, int it);                                                                     \
  int fv3t_##P##_halo_pack(fv3t_ctx* ctx, int it, int local_tile, int edge, REAL* dev_buf);                             \
  int fv3t_##P##_halo_unpack(fv3t_ctx* ctx, int it, int local_tile, int edge, const REAL* dev_buf);                     \
  /* the same with HOST buffers (3*(npx-1)*npz*nq elements): for a host whose MPI is not CUDA-aware -- the strip is packed on  \
     the device, copied out, exchanged by the host (mpp_send / mpp_recv in the Fortran shim), copied in and scattered */        \
  int fv3t_##P##_halo_pack_host(fv3t_ctx* ctx, int it, int local_tile, int edge, REAL* host_buf);                        \
  int fv3t_##P##_halo_unpack_host(fv3t_ctx* ctx, int it, int local_tile, int edge, const REAL* host_buf);                \
  /* Generic halo exchange by gather list (sub-tile contexts; works for whole tiles too): `list` names a set of plane offsets    \
     (row-major (j+2)*(npx+5) + (i+2)) registered with fv3t_halo_list_create.  gather: dev_buf[pl*buf_stride + e] = q_in(it)     \
     [tile local_tile][plane pl][offset e] for the planes pl = iq*npz + k of the resident tracers; scatter is the inverse.      \
     buf_stride = 0 stands for the length of the list; a larger stride lets several lists share one plane-major message.  Levels \
     whose ksplt(k) < it are skipped on both sides, like the strips above.  local_tile = -1: the list holds FLAT offsets         \
     local_tile * (npx+5)^2 + offset into the stack of resident tiles (one launch for the cells of several tiles). */          \
  int fv3t_##P##_halo_gather(fv3t_ctx* ctx, int it, int local_tile, int list, REAL* dev_buf, int buf_stride);            \
  int fv3t_##P##_halo_scatter(fv3t_ctx* ctx, int it, int local_tile, int list, const REAL* dev_buf, int buf_stride);     \
  int fv3t_##P##_tracer_2d_substep(fv3t_ctx* ctx, int it, int hord, REAL lim_fac);                                      \
  int fv3t_##P##_tracer_2d_finish(fv3t_ctx* ctx);

FV3T_DECLARE(f64, double)
FV3T_DECLARE(f32, float)

/* Precision-independent calls */
/* Registers `count` plane offsets (host array) as a gather / scatter list of this context; *list receives its handle. */
int fv3t_halo_list_create(fv3t_ctx* ctx, const int* offsets, int count, int* list);
int fv3t_halo_list_count(fv3t_ctx* ctx, int list);
/* Replaces the table fv3t_*_halo_local applies (halo cells whose source is resident in the SAME context): len pairs of flat
   offsets (local_tile * (npx+5) + j + 2) * (npx+5) + i + 2 into the stack of resident planes.  Whole-tile contexts build this
   table themselves from the cubed-sphere contact table; a sub-tile context starts with an empty one. */
int fv3t_halo_local_table(fv3t_ctx* ctx, const int* dst, const int* src, int len);
int fv3t_destroy(fv3t_ctx* ctx);
int fv3t_sync(fv3t_ctx* ctx);
/* Base of the device mirror.  FV3T_Q: the buffer that CURRENTLY holds the tracers -- tracer_2d (odd nsplt) and the remap
   write the other ping-pong buffer and flip it, so the pointer is valid only until the next tracer_2d / remap call.   */
void* fv3t_device_ptr(fv3t_ctx* ctx, int field);
size_t fv3t_halo_strip_elems(fv3t_ctx* ctx);     /* elements of one packed edge strip: 3*(npx-1)*npz*nq_cur  */
int fv3t_neighbor(fv3t_ctx* ctx, int global_tile, int edge, int* nbr_tile, int* nbr_edge, int* rotated);
uint64_t fv3t_kernel_launches(fv3t_ctx* ctx);    /* kernels launched by this context so far (bench accounting) */
/* CUDA-event timing of the kernels enqueued between the two calls on the context's stream */
int fv3t_timer_start(fv3t_ctx* ctx);
int fv3t_timer_stop_ms(fv3t_ctx* ctx, float* ms);
/* per-kernel-class accumulated device time since the last reset: 0 advect, 1 remap, 2 halo, 3 cmax, 4 scale    */
int fv3t_profile_enable(fv3t_ctx* ctx, int on);
int fv3t_profile_get_ms(fv3t_ctx* ctx, int kernel_class, float* total_ms, int* launches);

#ifdef __cplusplus
}
#endif
#endif /* FV3TRACER_H */
