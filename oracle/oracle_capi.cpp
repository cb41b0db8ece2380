// TEST INFRASTRUCTURE ONLY -- C entry points of the CPU oracle (see fv3_oracle_advect.hpp).
// Built by oracle/Makefile into oracle/liboracle.so; loaded with ctypes by tests/oracle_binding.py,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Never by the product.
#include <cstdint>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "fv3_oracle_advect.hpp"
#include "fv3_oracle_remap.hpp"

using namespace fv3oracle;

namespace {

template <class T>
void c_ppm_line(T* flux, const T* q1, const T* c, const T* dxa, int iord, int is, int ie, int isd, int npx, int edges, T lim_fac) {
  PpmScratch<T> w;
  ppm_line<T>(SLine<T>{flux, is, 1}, SLine<const T>{q1, isd, 1}, SLine<const T>{c, is, 1}, SLine<const T>{dxa, isd, 1}, iord, is,
              ie, npx, edges != 0, lim_fac, w);
}

template <class T>
void c_tracer_2d(int ntiles, int n, int npz, int nq, T* q, T* dp1, T* mfx, T* mfy, T* cx, T* cy, const T* area, const T* rarea,
                 const T* dx, const T* dy, const T* dxa, const T* dya, const T* sin_sg, const int64_t* halo_dst,
                 const int64_t* halo_src, int64_t halo_len, int hord, int q_split, T lim_fac, int* nsplt_out, int* ksplt_out,
                 T* cmax_out, const T* del6_u = nullptr, const T* del6_v = nullptr, T da_min = T(0), int nord_tr = 0, T trdm = T(0)) {
  const long nd = n + 6;
  std::vector<GridT<T>> g(ntiles);
  for (int t = 0; t < ntiles; ++t) {
    g[t].area = area + (long)t * nd * nd;
    g[t].rarea = rarea + (long)t * nd * nd;
    g[t].dx = dx + (long)t * nd * (nd + 1);
    g[t].dy = dy + (long)t * (nd + 1) * nd;
    g[t].dxa = dxa + (long)t * nd * nd;
    g[t].dya = dya + (long)t * nd * nd;
    g[t].sin_sg = sin_sg + (long)t * nd * nd * 5;
    if (del6_u) g[t].del6_u = del6_u + (long)t * nd * (nd + 1);
    if (del6_v) g[t].del6_v = del6_v + (long)t * (nd + 1) * nd;
    g[t].da_min = da_min;
  }
  Mosaic<T> m{ntiles, n, npz, nq, q, dp1, mfx, mfy, cx, cy, g.data(), halo_dst, halo_src, halo_len};
  tracer_2d_mosaic<T>(m, hord, q_split, lim_fac, nsplt_out, ksplt_out, cmax_out, nord_tr, trdm);
}

template <class T>
void c_tracer_2d_1l(int ntiles, int n, int npz, int nq, T* q, T* dp1, T* mfx, T* mfy, T* cx, T* cy, const T* area, const T* rarea,
                    const T* dx, const T* dy, const T* dxa, const T* dya, const T* sin_sg, const int64_t* halo_dst,
                    const int64_t* halo_src, int64_t halo_len, int hord, T lim_fac, int* nsplt_out, int* ksplt_out, T* cmax_out) {
  const long nd = n + 6;
  std::vector<GridT<T>> g(ntiles);
  for (int t = 0; t < ntiles; ++t) {
    g[t].area = area + (long)t * nd * nd;
    g[t].rarea = rarea + (long)t * nd * nd;
    g[t].dx = dx + (long)t * nd * (nd + 1);
    g[t].dy = dy + (long)t * (nd + 1) * nd;
    g[t].dxa = dxa + (long)t * nd * nd;
    g[t].dya = dya + (long)t * nd * nd;
    g[t].sin_sg = sin_sg + (long)t * nd * nd * 5;
  }
  Mosaic<T> m{ntiles, n, npz, nq, q, dp1, mfx, mfy, cx, cy, g.data(), halo_dst, halo_src, halo_len};
  tracer_2d_1L_mosaic<T>(m, hord, lim_fac, nsplt_out, ksplt_out, cmax_out);
}

template <class T>
void c_fv_tp_2d(int n, T* q, const T* crx, const T* cry, int hord, T* fx, T* fy, const T* xfx, const T* yfx, const T* ra_x,
                const T* ra_y, const T* area, const T* dxa, const T* dya, T lim_fac, const T* mfx, const T* mfy) {
  const Bounds bd = Bounds::tile(n);
  const long nxd = n + 6;
  GridT<T> g{};
  g.area = area;
  g.dxa = dxa;
  g.dya = dya;
  Tp2dScratch<T> w;
  w.size(bd);
  fv_tp_2d<T>(V2<T>{q, bd.isd, bd.jsd, nxd}, V2<const T>{crx, 1, bd.jsd, (long)n + 1}, V2<const T>{cry, bd.isd, 1, nxd}, n + 1,
              n + 1, hord, V2<T>{fx, 1, 1, (long)n + 1}, V2<T>{fy, 1, 1, (long)n}, V2<const T>{xfx, 1, bd.jsd, (long)n + 1},
              V2<const T>{yfx, bd.isd, 1, nxd}, g, bd, V2<const T>{ra_x, 1, bd.jsd, (long)n}, V2<const T>{ra_y, bd.isd, 1, nxd},
              lim_fac, mfx, mfy, w);
}

// fv_tp_2d with every optional argument: mfx/mfy (both or neither), mass, nord (< 0: absent), damp_c
template <class T>
void c_fv_tp_2d_full(int n, T* q, const T* crx, const T* cry, int hord, T* fx, T* fy, const T* xfx, const T* yfx, const T* ra_x,
                     const T* ra_y, const T* area, const T* dxa, const T* dya, const T* rarea, const T* del6_u, const T* del6_v,
                     T da_min, T lim_fac, const T* mfx, const T* mfy, const T* mass, int nord, T damp_c) {
  const Bounds bd = Bounds::tile(n);
  const long nxd = n + 6;
  GridT<T> g{};
  g.area = area;
  g.rarea = rarea;
  g.dxa = dxa;
  g.dya = dya;
  g.del6_u = del6_u;
  g.del6_v = del6_v;
  g.da_min = da_min;
  Tp2dScratch<T> w;
  w.size(bd);
  fv_tp_2d<T>(V2<T>{q, bd.isd, bd.jsd, nxd}, V2<const T>{crx, 1, bd.jsd, (long)n + 1}, V2<const T>{cry, bd.isd, 1, nxd}, n + 1,
              n + 1, hord, V2<T>{fx, 1, 1, (long)n + 1}, V2<T>{fy, 1, 1, (long)n}, V2<const T>{xfx, 1, bd.jsd, (long)n + 1},
              V2<const T>{yfx, bd.isd, 1, nxd}, g, bd, V2<const T>{ra_x, 1, bd.jsd, (long)n}, V2<const T>{ra_y, bd.isd, 1, nxd},
              lim_fac, mfx, mfy, w, mass, nord, damp_c);
}

template <class T> void c_copy_corners(T* q, int n, int dir) {
  const Bounds bd = Bounds::tile(n);
  GridT<T> g{};
  copy_corners<T>(V2<T>{q, bd.isd, bd.jsd, (long)n + 6}, n + 1, n + 1, dir, bd, g);
}

template <class T>
void c_remap_tracers(int ntiles, int n, int km, int nq, const T* pe, const T* ak, const T* bk, T ptop, T* q, T* delp,
                     const int* kord_tr, int fill) {
  const long nd = n + 6, plane = nd * nd;
  const long pe_tile = (long)(n + 2) * (km + 1) * (n + 2);
  for (int t = 0; t < ntiles; ++t)
    remap_tracers_tile<T>(n, km, nq, pe + t * pe_tile, ak, bk, ptop, q + (long)t * plane * km * nq, delp + (long)t * plane * km,
                          kord_tr, fill != 0);
}

// a4 in Fortran layout a4(4, km) for one column (0-based C: a4[4*k + c])
template <class T> void c_profile_col(int which, T* a4, const T* delp, int km, int iv, int kord, T qmin, T qs) {
  ColA4<T> A;
  A.size(km);
  ProfScratch<T> w;
  w.size(km);
  std::vector<T> dp(km + 2);
  for (int k = 1; k <= km; ++k) {
    dp[k] = delp[k - 1];
    for (int c = 1; c <= 4; ++c) A(c, k) = a4[4 * (k - 1) + (c - 1)];
  }
  if (which == 0)
    cs_profile_col<T>(true, qs, A, dp.data(), km, iv, kord, qmin, w);
  else if (which == 1)
    cs_profile_col<T>(false, qs, A, dp.data(), km, iv, kord, qmin, w);
  else
    ppm_profile_col<T>(A, dp.data(), km, iv, kord, w);
  for (int k = 1; k <= km; ++k)
    for (int c = 1; c <= 4; ++c) a4[4 * (k - 1) + (c - 1)] = A(c, k);
}

// q(km, nq) Fortran layout for one column: q[k + km*ic]
template <class T> void c_fillz_col(int km, int nq, T* q, const T* dp) {
  std::vector<T> d(km + 2), dm;
  for (int k = 1; k <= km; ++k) d[k] = dp[k - 1];
  auto Q = [&](int k, int ic) -> T& { return q[(k - 1) + (long)km * (ic - 1)]; };
  fillz_col<T>(km, nq, Q, d.data(), dm);
}

// one column of mapn_tracer (which=0) or map1_q2 per tracer (which=1); q(km,nq) in/out
template <class T>
void c_map_col(int which, int km, int nq, const T* pe1, const T* pe2, T* q, const int* kord, T q_min, int fill) {
  RemapScratch<T> w;
  w.size(km, nq);
  std::vector<T> p1(km + 3), p2(km + 3), dp2(km + 3), q2(km + 2);
  for (int k = 1; k <= km + 1; ++k) {
    p1[k] = pe1[k - 1];
    p2[k] = pe2[k - 1];
  }
  for (int k = 1; k <= km; ++k) dp2[k] = p2[k + 1] - p2[k];
  if (which == 0) {
    for (int iq = 0; iq < nq; ++iq)
      for (int k = 1; k <= km; ++k) w.a4[iq](1, k) = q[(k - 1) + (long)km * iq];
    mapn_tracer_col<T>(nq, km, p1.data(), p2.data(), dp2.data(), kord, q_min, fill != 0, w);
    for (int iq = 0; iq < nq; ++iq)
      for (int k = 1; k <= km; ++k) q[(k - 1) + (long)km * iq] = w.q2[(size_t)iq * (km + 2) + k];
  } else {
    for (int iq = 0; iq < nq; ++iq) {
      for (int k = 1; k <= km; ++k) w.a4[0](1, k) = q[(k - 1) + (long)km * iq];
      map1_q2_col<T>(km, p1.data(), w.a4[0], km, p2.data(), q2.data(), dp2.data(), 0, kord[iq], q_min, w);
      if (fill) {
        auto qq = [&](int k, int) -> T& { return q2[k]; };
        fillz_col<T>(km, 1, qq, dp2.data(), w.dm);
      }
      for (int k = 1; k <= km; ++k) q[(k - 1) + (long)km * iq] = q2[k];
    }
  }
}

// map_scalar (use_cs = 0, fv_mapz.F90:1199-1290) / map1_ppm (use_cs = 1, :1293-1383) for one column, kn = km
template <class T> void c_map_field_col(int use_cs, int km, const T* pe1, const T* pe2, T* q, int iv, int kord, T q_min, T qs) {
  RemapScratch<T> w;
  w.size(km, 1);
  std::vector<T> p1(km + 3), p2(km + 3), dp2(km + 3), q2(km + 2);
  for (int k = 1; k <= km + 1; ++k) {
    p1[k] = pe1[k - 1];
    p2[k] = pe2[k - 1];
  }
  for (int k = 1; k <= km; ++k) {
    dp2[k] = p2[k + 1] - p2[k];
    w.a4[0](1, k) = q[k - 1];
  }
  map1_q2_col<T>(km, p1.data(), w.a4[0], km, p2.data(), q2.data(), dp2.data(), iv, kord, q_min, w, use_cs == 0, qs);
  for (int k = 1; k <= km; ++k) q[k - 1] = q2[k];
}

}  // namespace

#define ORC_API(T, S)                                                                                                          \
  extern "C" void orc_##S##_ppm_line(T* flux, const T* q1, const T* c, const T* dxa, int iord, int is, int ie, int isd,     \
                                       int npx, int edges, T lim_fac) {                                                        \
    c_ppm_line<T>(flux, q1, c, dxa, iord, is, ie, isd, npx, edges, lim_fac);                                                   \
  }                                                                                                                            \
  extern "C" void orc_##S##_tracer_2d(int ntiles, int n, int npz, int nq, T* q, T* dp1, T* mfx, T* mfy, T* cx, T* cy,       \
                                        const T* area, const T* rarea, const T* dx, const T* dy, const T* dxa, const T* dya,  \
                                        const T* sin_sg, const int64_t* halo_dst, const int64_t* halo_src, int64_t halo_len,  \
                                        int hord, int q_split, T lim_fac, int* nsplt_out, int* ksplt_out, T* cmax_out) {      \
    c_tracer_2d<T>(ntiles, n, npz, nq, q, dp1, mfx, mfy, cx, cy, area, rarea, dx, dy, dxa, dya, sin_sg, halo_dst, halo_src,    \
                   halo_len, hord, q_split, lim_fac, nsplt_out, ksplt_out, cmax_out);                                          \
  }                                                                                                                            \
  extern "C" void orc_##S##_tracer_2d_damp(int ntiles, int n, int npz, int nq, T* q, T* dp1, T* mfx, T* mfy, T* cx, T* cy,  \
                                             const T* area, const T* rarea, const T* dx, const T* dy, const T* dxa,          \
                                             const T* dya, const T* sin_sg, const int64_t* halo_dst, const int64_t* halo_src, \
                                             int64_t halo_len, int hord, int q_split, T lim_fac, int* nsplt_out,             \
                                             int* ksplt_out, T* cmax_out, const T* del6_u, const T* del6_v, T da_min,         \
                                             int nord_tr, T trdm) {                                                           \
    c_tracer_2d<T>(ntiles, n, npz, nq, q, dp1, mfx, mfy, cx, cy, area, rarea, dx, dy, dxa, dya, sin_sg, halo_dst, halo_src,    \
                   halo_len, hord, q_split, lim_fac, nsplt_out, ksplt_out, cmax_out, del6_u, del6_v, da_min, nord_tr, trdm);   \
  }                                                                                                                            \
  extern "C" void orc_##S##_tracer_2d_1l(int ntiles, int n, int npz, int nq, T* q, T* dp1, T* mfx, T* mfy, T* cx, T* cy,    \
                                           const T* area, const T* rarea, const T* dx, const T* dy, const T* dxa,           \
                                           const T* dya, const T* sin_sg, const int64_t* halo_dst, const int64_t* halo_src, \
                                           int64_t halo_len, int hord, T lim_fac, int* nsplt_out, int* ksplt_out,            \
                                           T* cmax_out) {                                                                     \
    c_tracer_2d_1l<T>(ntiles, n, npz, nq, q, dp1, mfx, mfy, cx, cy, area, rarea, dx, dy, dxa, dya, sin_sg, halo_dst,         \
                      halo_src, halo_len, hord, lim_fac, nsplt_out, ksplt_out, cmax_out);                                    \
  }                                                                                                                            \
  extern "C" void orc_##S##_fv_tp_2d(int n, T* q, const T* crx, const T* cry, int hord, T* fx, T* fy, const T* xfx,         \
                                       const T* yfx, const T* ra_x, const T* ra_y, const T* area, const T* dxa, const T* dya, \
                                       T lim_fac, const T* mfx, const T* mfy) {                                                \
    c_fv_tp_2d<T>(n, q, crx, cry, hord, fx, fy, xfx, yfx, ra_x, ra_y, area, dxa, dya, lim_fac, mfx, mfy);                      \
  }                                                                                                                            \
  extern "C" void orc_##S##_fv_tp_2d_full(int n, T* q, const T* crx, const T* cry, int hord, T* fx, T* fy, const T* xfx,    \
                                            const T* yfx, const T* ra_x, const T* ra_y, const T* area, const T* dxa,           \
                                            const T* dya, const T* rarea, const T* del6_u, const T* del6_v, T da_min,          \
                                            T lim_fac, const T* mfx, const T* mfy, const T* mass, int nord, T damp_c) {        \
    c_fv_tp_2d_full<T>(n, q, crx, cry, hord, fx, fy, xfx, yfx, ra_x, ra_y, area, dxa, dya, rarea, del6_u, del6_v, da_min,      \
                       lim_fac, mfx, mfy, mass, nord, damp_c);                                                                 \
  }                                                                                                                            \
  extern "C" void orc_##S##_copy_corners(T* q, int n, int dir) { c_copy_corners<T>(q, n, dir); }                             \
  extern "C" void orc_##S##_remap_tracers(int ntiles, int n, int km, int nq, const T* pe, const T* ak, const T* bk, T ptop, \
                                            T* q, T* delp, const int* kord_tr, int fill) {                                     \
    c_remap_tracers<T>(ntiles, n, km, nq, pe, ak, bk, ptop, q, delp, kord_tr, fill);                                           \
  }                                                                                                                            \
  extern "C" void orc_##S##_profile_col(int which, T* a4, const T* delp, int km, int iv, int kord, T qmin, T qs) {          \
    c_profile_col<T>(which, a4, delp, km, iv, kord, qmin, qs);                                                                 \
  }                                                                                                                            \
  extern "C" void orc_##S##_fillz_col(int km, int nq, T* q, const T* dp) { c_fillz_col<T>(km, nq, q, dp); }                  \
  extern "C" void orc_##S##_map_col(int which, int km, int nq, const T* pe1, const T* pe2, T* q, const int* kord, T q_min,  \
                                      int fill) {                                                                              \
    c_map_col<T>(which, km, nq, pe1, pe2, q, kord, q_min, fill);                                                               \
  }                                                                                                                            \
  extern "C" void orc_##S##_map_field_col(int use_cs, int km, const T* pe1, const T* pe2, T* q, int iv, int kord, T q_min,    \
                                            T qs) {                                                                            \
    c_map_field_col<T>(use_cs, km, pe1, pe2, q, iv, kord, q_min, qs);                                                          \
  }

ORC_API(float, f32)
ORC_API(double, f64)

extern "C" void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
extern "C" int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
