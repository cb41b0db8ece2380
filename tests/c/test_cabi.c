/* Plain-C client of the C-ABI (include/fv3tracer.h), as a Fortran / C host would bind it: no Python, no ctypes.
 *
 *   gcc -std=c99 -I include tests/c/test_cabi.c -o test_cabi -L fv3atm_b200 -lfv3tracer -lm
 *
 * Builds a one-level-of-detail planar "mosaic" (six identical flat tiles: unit metrics) and calls EVERY entry point the header
 * declares for the f64 symbol set (the f32 set is the same code instantiated for float and is exercised by the Python tests):
 *   - with zero Courant numbers / mass fluxes tracer_2d and tracer_2d_1L must return q bit-for-bit and nsplt = 1;
 *   - with pe2 == pe1 (a Lagrangian surface equal to the Eulerian one) the remap is the identity to rounding and the
 *     flux-form column sum is conserved; delp must equal the layer thicknesses of ak + bk*ps bit for bit;
 *   - the row-granular mapn_tracer entry reproduces the batched remap of the same rows;
 *   - the sub-step building blocks (begin / set_cmax / halo_local / halo_pack / halo_unpack / substep / finish) reproduce
 *     tracer_2d; upload / download round-trip; timers, profile counters, neighbour table, error strings.
 * Without a CUDA device it checks that creation fails loudly with the documented message and exits 0 ("no CPU fallback").
 * Exit code 0 = all checks passed. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fv3tracer.h"

#define N 12
#define NPZ 8
#define NQ 6
#define ND (N + 6)
#define CHECK(cond, msg)                                                  \
  do {                                                                    \
    if (!(cond)) {                                                        \
      fprintf(stderr, "FAIL %s:%d: %s (%s)\n", __FILE__, __LINE__, msg, fv3t_last_error()); \
      return 1;                                                           \
    }                                                                     \
  } while (0)
#define OK(call) CHECK((call) == 0, #call)

static double* filled(size_t n, double v) {
  double* p = (double*)malloc(n * sizeof(double));
  for (size_t i = 0; i < n; ++i) p[i] = v;
  return p;
}

int main(void) {
  const size_t plane = (size_t)ND * ND;
  fv3t_dims d;
  memset(&d, 0, sizeof d);
  d.npx = N + 1;
  d.npz = NPZ;
  d.nq_max = NQ;
  d.ntiles = 6;
  for (int t = 0; t < 6; ++t) d.tile_id[t] = t + 1;
  fv3t_f64_grid g;
  g.area = filled(6 * plane, 1.0);
  g.rarea = filled(6 * plane, 1.0);
  g.dx = filled(6 * (size_t)ND * (ND + 1), 1.0);
  g.dy = filled(6 * (size_t)(ND + 1) * ND, 1.0);
  g.dxa = filled(6 * plane, 1.0);
  g.dya = filled(6 * plane, 1.0);
  g.sin_sg = filled(6 * plane * 5, 1.0);

  fv3t_ctx* ctx = NULL;
  if (fv3t_device_count() == 0) {
    const int rc = fv3t_f64_create(&ctx, &d, &g, 0, NULL);
    CHECK(rc != 0 && ctx == NULL, "create must fail without a CUDA device");
    CHECK(strstr(fv3t_last_error(), "no CUDA device") != NULL, "the failure must say why");
    CHECK(fv3t_destroy(NULL) == 0 && fv3t_sync(NULL) != 0, "null context handling");
    printf("test_cabi: no CUDA device -- creation failed loudly as documented (%s)\n", fv3t_last_error());
    return 0;
  }
  OK(fv3t_f64_create(&ctx, &d, &g, 0, NULL));

  /* fields: q = smooth positive function of (tile, tracer, level, j, i); uniform dp1; zero winds */
  const size_t nq_el = 6 * (size_t)NQ * NPZ * plane, nc = 6 * (size_t)NPZ * plane;
  const size_t ncx = 6 * (size_t)NPZ * ND * (N + 1), nmf = 6 * (size_t)NPZ * N * (N + 1), npe = 6 * (size_t)(N + 2) * (NPZ + 1) * (N + 2);
  double* q = (double*)malloc(nq_el * sizeof(double));
  for (size_t e = 0; e < nq_el; ++e) q[e] = 1.0 + 0.5 * sin(0.37 * (double)(e % 1013)) + 0.001 * (double)(e % 7);
  double* q0 = (double*)malloc(nq_el * sizeof(double));
  memcpy(q0, q, nq_el * sizeof(double));
  double *dp1 = filled(nc, 1000.0), *cx = filled(ncx, 0.0), *cy = filled(ncx, 0.0), *mfx = filled(nmf, 0.0), *mfy = filled(nmf, 0.0);
  int nsplt = -1, ksplt[NPZ];

  /* ---- tracer_2d / tracer_2d_1L with no wind: identity, nsplt = 1 */
  OK(fv3t_f64_tracer_2d(ctx, q, dp1, mfx, mfy, cx, cy, NQ, 8, 0, 0, 0.0, 1.0, &nsplt, ksplt));
  CHECK(nsplt == 1 && ksplt[0] == 1 && ksplt[NPZ - 1] == 1, "nsplt");
  for (int t = 0; t < 6; ++t)
    for (int l = 0; l < NQ * NPZ; ++l)
      for (int j = 3; j < ND - 3; ++j)
        for (int i = 3; i < ND - 3; ++i) {
          const size_t e = ((size_t)t * NQ * NPZ + l) * plane + (size_t)j * ND + i;
          CHECK(q[e] == q0[e], "tracer_2d with zero Courant numbers must not change q");
        }
  OK(fv3t_f64_tracer_2d_1L(ctx, q, dp1, mfx, mfy, cx, cy, NQ, 10, 0, 0, 0.0, 1.0, &nsplt, ksplt));
  CHECK(nsplt == 1, "tracer_2d_1L nsplt");
  CHECK(fv3t_f64_tracer_2d(ctx, q, dp1, mfx, mfy, cx, cy, NQ, 8, 0, 2, 0.2, 1.0, &nsplt, ksplt) != 0, "tracer damping without its metrics is an error");
  CHECK(strstr(fv3t_last_error(), "set_damping") != NULL, "error string names the missing call");
  {
    /* with the metrics: a horizontally uniform field has no del-n flux, so q must come back unchanged to rounding */
    double* d6u = filled(6 * (size_t)ND * (ND + 1), 1.0);
    double* d6v = filled(6 * (size_t)ND * (ND + 1), 1.0);
    OK(fv3t_f64_set_damping(ctx, d6u, d6v, 1.0, 1, 0.0));
    double* qu = filled(nq_el, 3.25);
    OK(fv3t_f64_tracer_2d(ctx, qu, dp1, mfx, mfy, cx, cy, NQ, 8, 0, 1, 0.15, 1.0, &nsplt, ksplt));
    for (int j = 3; j < ND - 3; ++j) CHECK(fabs(qu[(size_t)j * ND + 5] - 3.25) < 1e-13, "damping leaves a uniform field alone");
    free(d6u), free(d6v), free(qu);
  }

  /* ---- fv_tp_2d as an operator: a uniform field at rest has the flux  field value x flux area  on every face, with and without
     mass fluxes; the two are given together or not at all */
  {
    const int nlev = 2;
    const size_t npl = 6 * (size_t)nlev;
    double* qf = filled(npl * plane, 3.25);
    double *crx = filled(npl * ND * (N + 1), 0.0), *cry = filled(npl * ND * (N + 1), 0.0);
    double *xfx = filled(npl * ND * (N + 1), 2.0), *yfx = filled(npl * ND * (N + 1), 2.0);
    double *rax = filled(npl * ND * N, 1.0), *ray = filled(npl * ND * N, 1.0);
    double *fx = filled(npl * N * (N + 1), -1.0), *fy = filled(npl * N * (N + 1), -1.0), *mf = filled(npl * N * (N + 1), 4.0);
    OK(fv3t_f64_fv_tp_2d(ctx, nlev, qf, crx, cry, 10, fx, fy, xfx, yfx, rax, ray, 1.0, NULL, NULL, NULL, -1, 0.0));
    for (size_t e = 0; e < npl * N * (N + 1); ++e) CHECK(fx[e] == 6.5 && fy[e] == 6.5, "fv_tp_2d: flux of a uniform field at rest = q * xfx");
    OK(fv3t_f64_fv_tp_2d(ctx, nlev, qf, crx, cry, 8, fx, fy, xfx, yfx, rax, ray, 1.0, mf, mf, NULL, -1, 0.0));
    for (size_t e = 0; e < npl * N * (N + 1); ++e) CHECK(fx[e] == 13.0 && fy[e] == 13.0, "fv_tp_2d: ... = q * mfx with mass fluxes");
    CHECK(fv3t_f64_fv_tp_2d(ctx, nlev, qf, crx, cry, 5, fx, fy, xfx, yfx, rax, ray, 1.0, mf, NULL, NULL, -1, 0.0) != 0, "mfx without mfy is an error");
    CHECK(fv3t_f64_fv_tp_2d(ctx, nlev, qf, crx, cry, 14, fx, fy, xfx, yfx, rax, ray, 1.0, NULL, NULL, NULL, -1, 0.0) != 0, "hord 14 is not a scheme");
    free(qf), free(crx), free(cry), free(xfx), free(yfx), free(rax), free(ray), free(fx), free(fy), free(mf);
  }

  /* ---- remap with pe1 == pe2: identity to rounding, delp = diff(ak + bk ps) bit for bit */
  double ak[NPZ + 1], bk[NPZ + 1];
  const double ptop = 100.0, ps = 1.0e5;
  for (int k = 0; k <= NPZ; ++k) {
    const double s = (double)k / NPZ;
    bk[k] = s * s;
    ak[k] = ptop * (1.0 - s) * (1.0 + 3.0 * s);
  }
  ak[NPZ] = 0.0;
  bk[NPZ] = 1.0;
  double* pe = (double*)malloc(npe * sizeof(double));
  for (int t = 0; t < 6; ++t)
    for (int j = 0; j < N + 2; ++j)
      for (int k = 0; k <= NPZ; ++k)
        for (int i = 0; i < N + 2; ++i) pe[(((size_t)t * (N + 2) + j) * (NPZ + 1) + k) * (N + 2) + i] = k == 0 ? ptop : ak[k] + bk[k] * ps;
  double* delp = filled(nc, 0.0);
  int kord[NQ];
  for (int iq = 0; iq < NQ; ++iq) kord[iq] = 9;
  memcpy(q, q0, nq_el * sizeof(double));
  OK(fv3t_f64_remap_tracers(ctx, pe, ak, bk, ptop, q, delp, NQ, kord, 1));
  double worst = 0.0;
  for (int t = 0; t < 6; ++t)
    for (int iq = 0; iq < NQ; ++iq)
      for (int k = 0; k < NPZ; ++k)
        for (int j = 3; j < ND - 3; ++j)
          for (int i = 3; i < ND - 3; ++i) {
            const size_t e = (((size_t)t * NQ + iq) * NPZ + k) * plane + (size_t)j * ND + i;
            const double dd = fabs(q[e] - q0[e]);
            if (dd > worst) worst = dd;
            const double dpk = (k + 1 == NPZ ? ps : ak[k + 1] + bk[k + 1] * ps) - (k == 0 ? ptop : ak[k] + bk[k] * ps);
            CHECK(delp[((size_t)t * NPZ + k) * plane + (size_t)j * ND + i] == dpk, "delp <- dp2");
          }
  CHECK(worst <= 2e-12, "remap onto the same grid is the identity to rounding");

  /* ---- row-granular mapn_tracer on a one-tile context == batched remap of that tile */
  {
    fv3t_dims d1 = d;
    d1.ntiles = 1;
    fv3t_ctx* c1 = NULL;
    OK(fv3t_f64_create(&c1, &d1, &g, 0, NULL));
    /* a target grid that is NOT ak + bk ps: every interior interface shifted by a quarter layer */
    double* pe1r = (double*)malloc((size_t)(NPZ + 1) * N * sizeof(double));
    double* pe2r = (double*)malloc((size_t)(NPZ + 1) * N * sizeof(double));
    double* dp2r = (double*)malloc((size_t)NPZ * N * sizeof(double));
    for (int k = 0; k <= NPZ; ++k)
      for (int i = 0; i < N; ++i) {
        const double p1 = k == 0 ? ptop : ak[k] + bk[k] * ps;
        const double pn = k == NPZ ? p1 : ak[k + 1] + bk[k + 1] * ps;
        pe1r[(size_t)k * N + i] = p1;
        pe2r[(size_t)k * N + i] = (k == 0 || k == NPZ) ? p1 : p1 + 0.25 * (pn - p1);
      }
    for (int k = 0; k < NPZ; ++k)
      for (int i = 0; i < N; ++i) dp2r[(size_t)k * N + i] = pe2r[(size_t)(k + 1) * N + i] - pe2r[(size_t)k * N + i];
    double* q1 = (double*)malloc((size_t)NQ * NPZ * plane * sizeof(double));
    memcpy(q1, q0, (size_t)NQ * NPZ * plane * sizeof(double));
    double sum0 = 0.0, sum1 = 0.0;
    const int jrow = 4, icol = 5;
    for (int k = 0; k < NPZ; ++k) sum0 += q1[(size_t)k * plane + (size_t)(jrow + 2) * ND + icol + 2] * (pe1r[(size_t)(k + 1) * N + icol - 1] - pe1r[(size_t)k * N + icol - 1]);
    OK(fv3t_f64_mapn_tracer(c1, NQ, NPZ, pe1r, pe2r, q1, dp2r, kord, jrow, 1, N, -2, N + 3, -2, N + 3, 0.0, 1));
    for (int k = 0; k < NPZ; ++k) sum1 += q1[(size_t)k * plane + (size_t)(jrow + 2) * ND + icol + 2] * dp2r[(size_t)k * N + icol - 1];
    CHECK(fabs(sum1 - sum0) <= 1e-12 * fabs(sum0), "mapn_tracer conserves the column integral on the caller's own target grid");
    int moved = 0;
    for (int k = 0; k < NPZ; ++k) moved |= q1[(size_t)k * plane + (size_t)(jrow + 2) * ND + icol + 2] != q0[(size_t)k * plane + (size_t)(jrow + 2) * ND + icol + 2];
    CHECK(moved, "the shifted target grid must change the column (pe2 is consumed as given)");
    CHECK(fv3t_f64_mapn_tracer(c1, NQ, NPZ, pe1r, pe2r, q1, dp2r, kord, jrow, 2, N, -2, N + 3, -2, N + 3, 0.0, 1) != 0, "partial rows are rejected");
    OK(fv3t_destroy(c1));
    free(pe1r), free(pe2r), free(dp2r), free(q1);
  }

  /* ---- device-resident operation and the sub-step building blocks reproduce tracer_2d */
  memcpy(q, q0, nq_el * sizeof(double));
  for (size_t e = 0; e < ncx; ++e) cx[e] = 0.3 * sin(0.11 * (double)(e % 977));  /* some wind: Courant numbers < 1 */
  OK(fv3t_f64_upload(ctx, FV3T_Q, q, NQ));
  OK(fv3t_f64_upload(ctx, FV3T_DP1, dp1, NQ));
  OK(fv3t_f64_upload(ctx, FV3T_MFX, mfx, NQ));
  OK(fv3t_f64_upload(ctx, FV3T_MFY, mfy, NQ));
  OK(fv3t_f64_upload(ctx, FV3T_CX, cx, NQ));
  OK(fv3t_f64_upload(ctx, FV3T_CY, cy, NQ));
  OK(fv3t_f64_upload(ctx, FV3T_PE, pe, NQ));
  OK(fv3t_f64_set_vertical(ctx, ak, bk, ptop));
  OK(fv3t_sync(ctx));
  CHECK(fv3t_device_ptr(ctx, FV3T_Q) != NULL && fv3t_device_ptr(ctx, FV3T_DELP) != NULL, "device pointers");
  OK(fv3t_profile_enable(ctx, 1));
  OK(fv3t_timer_start(ctx));
  OK(fv3t_f64_remap_prepare(ctx));
  OK(fv3t_f64_tracer_2d_resident(ctx, NQ, 8, 0, 1.0, &nsplt));
  float ms = -1.f, kms = -1.f;
  int kl = 0;
  OK(fv3t_timer_stop_ms(ctx, &ms));
  OK(fv3t_profile_get_ms(ctx, 0, &kms, &kl));
  CHECK(ms > 0.f && kl >= 1 && fv3t_kernel_launches(ctx) > 0, "timers / launch counters");
  OK(fv3t_profile_enable(ctx, 0));
  double* qa = (double*)malloc(nq_el * sizeof(double));
  OK(fv3t_f64_download(ctx, FV3T_Q, qa, NQ));
  /* the same call assembled from its building blocks */
  OK(fv3t_f64_upload(ctx, FV3T_Q, q, NQ));
  double cmax[NPZ];
  OK(fv3t_f64_tracer_2d_begin(ctx, NQ, 0, cmax));
  int ns2 = 0;
  OK(fv3t_f64_tracer_2d_set_cmax(ctx, cmax, 0, &ns2));
  CHECK(ns2 == nsplt, "nsplt of the building blocks");
  for (int it = 1; it <= ns2; ++it) {
    OK(fv3t_f64_halo_local(ctx, it));
    OK(fv3t_f64_tracer_2d_substep(ctx, it, 8, 1.0));
  }
  OK(fv3t_f64_tracer_2d_finish(ctx));
  double* qb = (double*)malloc(nq_el * sizeof(double));
  OK(fv3t_f64_download(ctx, FV3T_Q, qb, NQ));
  for (int t = 0; t < 6; ++t)
    for (int l = 0; l < NQ * NPZ; ++l)
      for (int j = 3; j < ND - 3; ++j)
        for (int i = 3; i < ND - 3; ++i) {
          const size_t e = ((size_t)t * NQ * NPZ + l) * plane + (size_t)j * ND + i;
          CHECK(qa[e] == qb[e], "begin/set_cmax/halo_local/substep/finish == tracer_2d_resident");
        }
  OK(fv3t_f64_remap_tracers_resident(ctx, NQ, kord, 1));
  /* halo strips: pack on the device, unpack into the neighbour's halo, table lookups */
  CHECK(fv3t_halo_strip_elems(ctx) == (size_t)3 * N * NPZ * NQ, "strip size");
  int nt = 0, ne = 0, rot = -1;
  OK(fv3t_neighbor(ctx, 1, 1, &nt, &ne, &rot));
  CHECK(nt >= 1 && nt <= 6 && ne >= 0 && ne <= 3, "neighbour table");
  CHECK(fv3t_neighbor(ctx, 7, 0, &nt, &ne, &rot) != 0, "bad tile is an error");
  {
    /* host-staged strips: what tile 1 sends across its edge 1 is exactly what halo_local copies into the neighbour's halo */
    const size_t nel = fv3t_halo_strip_elems(ctx);
    double* strip = (double*)malloc(nel * sizeof(double));
    OK(fv3t_neighbor(ctx, 1, 1, &nt, &ne, &rot));
    OK(fv3t_f64_upload(ctx, FV3T_Q, q0, NQ));
    OK(fv3t_f64_halo_local(ctx, 1));
    double* qh = (double*)malloc(nq_el * sizeof(double));
    OK(fv3t_f64_download(ctx, FV3T_Q, qh, NQ));
    OK(fv3t_f64_upload(ctx, FV3T_Q, q0, NQ));
    OK(fv3t_f64_halo_pack_host(ctx, 1, 0, 1, strip));
    OK(fv3t_f64_halo_unpack_host(ctx, 1, nt - 1, ne, strip));
    double* qs = (double*)malloc(nq_el * sizeof(double));
    OK(fv3t_f64_download(ctx, FV3T_Q, qs, NQ));
    size_t changed = 0, wrong = 0;
    for (size_t e = 0; e < nq_el; ++e)
      if (qs[e] != q0[e]) {
        ++changed;
        wrong += qs[e] != qh[e];
      }
    CHECK(changed > 0 && wrong == 0, "pack_host -> unpack_host fills the neighbour's edge halo like halo_local");
    free(strip), free(qh), free(qs);
  }
  /* gather lists (the generic halo exchange of sub-tile contexts; valid for whole tiles too) and a sub-tile context */
  {
    int offs[3] = {3 * ND + 3, 3 * ND + 4, 4 * ND + 3}, flat[2] = {3 * ND + 3, (int)plane + 3 * ND + 3}, bad[1] = {6 * (int)plane}, id = -1, idf = -1;
    OK(fv3t_halo_list_create(ctx, offs, 3, &id));
    OK(fv3t_halo_list_create(ctx, flat, 2, &idf));
    CHECK(id >= 0 && fv3t_halo_list_count(ctx, id) == 3 && fv3t_halo_list_count(ctx, idf) == 2 && fv3t_halo_list_count(ctx, 99) == -1, "list handles");
    CHECK(fv3t_halo_list_create(ctx, bad, 1, &id) != 0, "offsets outside the resident planes are rejected");
    CHECK(fv3t_f64_halo_gather(ctx, 1, 0, id, NULL, 0) != 0, "gather into a null buffer is an error");
    CHECK(fv3t_f64_halo_gather(ctx, 1, 0, idf, (double*)fv3t_device_ptr(ctx, FV3T_DELP), 0) != 0, "a flat list needs local_tile = -1");
    /* delp's device mirror serves as a scratch buffer: gather three cells of every plane of tile 2, put them back */
    OK(fv3t_f64_upload(ctx, FV3T_Q, q0, NQ));
    OK(fv3t_f64_halo_gather(ctx, 1, 1, id, (double*)fv3t_device_ptr(ctx, FV3T_DELP), 0));
    OK(fv3t_f64_halo_gather(ctx, 1, -1, idf, (double*)fv3t_device_ptr(ctx, FV3T_DELP), 8));
    OK(fv3t_f64_halo_scatter(ctx, 1, -1, idf, (const double*)fv3t_device_ptr(ctx, FV3T_DELP), 8));
    OK(fv3t_f64_download(ctx, FV3T_Q, q, NQ));
    CHECK(memcmp(q, q0, nq_el * sizeof(double)) == 0, "gather followed by scatter of the same list leaves q unchanged");
    fv3t_dims ds = d;   /* six C12 sub-domains of a C24 mosaic */
    ds.sub_layout = 2;
    for (int t = 0; t < 6; ++t) ds.sub_bi[t] = t & 1, ds.sub_bj[t] = (t >> 1) & 1, ds.tile_id[t] = 1 + t / 4;
    fv3t_ctx* cs = NULL;
    OK(fv3t_f64_create(&cs, &ds, &g, 0, NULL));
    CHECK(fv3t_f64_tracer_2d_resident(cs, NQ, 8, 0, 1.0, &nsplt) != 0, "a sub-tile context is driven through the building blocks");
    int hd[1] = {2 * ND + 5}, hs[1] = {(int)plane + 9};
    OK(fv3t_halo_local_table(cs, hd, hs, 1));
    CHECK(fv3t_halo_local_table(cs, hd, bad, 1) != 0, "table offsets outside the resident planes are rejected");
    ds.sub_bi[0] = 2;
    fv3t_ctx* cbad = NULL;
    CHECK(fv3t_f64_create(&cbad, &ds, &g, 0, NULL) != 0, "block index outside the layout");
    OK(fv3t_destroy(cs));
  }
  /* tracer_step: both calls in one, host arrays in and out */
  memcpy(q, q0, nq_el * sizeof(double));
  OK(fv3t_f64_tracer_step(ctx, q, dp1, mfx, mfy, cx, cy, pe, ak, bk, ptop, delp, NQ, 8, 0, 1.0, kord, 1, &nsplt));
  CHECK(fv3t_f64_upload(ctx, 99, q, NQ) != 0 && fv3t_f64_download(ctx, FV3T_Q, q, NQ + 1) != 0, "bad field / nq are errors");
  CHECK(fv3t_f32_tracer_2d_resident(ctx, NQ, 8, 0, 1.0f, &nsplt) != 0, "precision mismatch is an error, not a crash");
  OK(fv3t_destroy(ctx));
  printf("test_cabi: all checks passed (identity remap max |dq| = %.2e)\n", worst);
  return 0;
}
