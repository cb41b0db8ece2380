// fv3atm_b200: context, host-side orchestration and the C-ABI (include/fv3tracer.h) of libfv3tracer.so.
//
// Host logic restated from the reference's tracer_2d driver (atmos_cubed_sphere/model/fv_tracer2d.F90:
// 432-457 nsplt/ksplt, :496-566 sub-step loop) and from the j-loop dispatch of Lagrangian_to_Eulerian
// (model/fv_mapz.F90:407-426).  The cubed-sphere mosaic is the 12-contact table of tools/fv_mp_mod.F90:581-629.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/fv3tracer.h"
#include "fv3t_advect.cuh"
#include "fv3t_advect2.cuh"
#include "fv3t_remap.cuh"
#include "fv3t_remap2.cuh"
#include "fv3t_fast.h"
#include "fv3t_deln.cuh"
#include "fv3t_tp2d.cuh"

namespace {

thread_local std::string g_err;

int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) return fail("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)

// ---- mosaic topology ------------------------------------------------------------------------------
enum { EW = 0, EE = 1, ES = 2, EN = 3 };
struct EdgeMap {
  int nbr_tile, nbr_edge;  // 0-based tile
  int A[2][2], b[2];       // halo cell (i,j) of this tile -> cell (A(i,j)+b) of nbr_tile, 1-based cells
};

struct Line4 {
  int is, ie, js, je;
};

void edge_of(const Line4& l, int n, int& edge, int p0[2], int d[2], int out[2]) {
  p0[0] = l.is;
  p0[1] = l.js;
  d[0] = (l.ie > l.is) - (l.ie < l.is);
  d[1] = (l.je > l.js) - (l.je < l.js);
  if (l.is == l.ie) {
    edge = (l.is == n) ? EE : EW;
    out[0] = (edge == EE) ? 1 : -1;
    out[1] = 0;
  } else {
    edge = (l.js == n) ? EN : ES;
    out[0] = 0;
    out[1] = (edge == EN) ? 1 : -1;
  }
}

// maps[tile][edge] for the six-tile cubed sphere
void build_edge_maps(int n, EdgeMap maps[6][4]) {
  const int nx = n, ny = n;
  struct C {
    int t1, t2;
    Line4 l1, l2;
  };
  const C contacts[12] = {
      {1, 2, {nx, nx, 1, ny}, {1, 1, 1, ny}},   {1, 3, {1, nx, ny, ny}, {1, 1, ny, 1}},   {1, 5, {1, 1, 1, ny}, {nx, 1, ny, ny}},
      {1, 6, {1, nx, 1, 1}, {1, nx, ny, ny}},   {2, 3, {1, nx, ny, ny}, {1, nx, 1, 1}},   {2, 4, {nx, nx, 1, ny}, {nx, 1, 1, 1}},
      {2, 6, {1, nx, 1, 1}, {nx, nx, ny, 1}},   {3, 4, {nx, nx, 1, ny}, {1, 1, 1, ny}},   {3, 5, {1, nx, ny, ny}, {1, 1, ny, 1}},
      {4, 5, {1, nx, ny, ny}, {1, nx, 1, 1}},   {4, 6, {nx, nx, 1, ny}, {nx, 1, 1, 1}},   {5, 6, {nx, nx, 1, ny}, {1, 1, 1, ny}},
  };
  for (const C& c : contacts) {
    int e1, e2, p1[2], p2[2], d1[2], d2[2], o1[2], o2[2];
    edge_of(c.l1, n, e1, p1, d1, o1);
    edge_of(c.l2, n, e2, p2, d2, o2);
    for (int side = 0; side < 2; ++side) {
      const int ta = side ? c.t2 : c.t1, tb = side ? c.t1 : c.t2;
      const int ea = side ? e2 : e1, eb = side ? e1 : e2;
      const int* pa = side ? p2 : p1;
      const int* pb = side ? p1 : p2;
      const int* da = side ? d2 : d1;
      const int* db = side ? d1 : d2;
      const int* oa = side ? o2 : o1;
      const int* ob = side ? o1 : o2;
      EdgeMap& m = maps[ta - 1][ea];
      m.nbr_tile = tb - 1;
      m.nbr_edge = eb;
      const int inb[2] = {-ob[0], -ob[1]};
      for (int r = 0; r < 2; ++r)
        for (int cc = 0; cc < 2; ++cc) m.A[r][cc] = db[r] * da[cc] + inb[r] * oa[cc];
      for (int r = 0; r < 2; ++r) m.b[r] = pb[r] - inb[r] - (m.A[r][0] * pa[0] + m.A[r][1] * pa[1]);
    }
  }
}

// canonical order of the halo strip beyond `edge`: depth m = 1..3 outer, position s = 1..n inner
inline void halo_cell(int edge, int n, int m, int s, int& i, int& j) {
  switch (edge) {
    case EW: i = 1 - m; j = s; break;
    case EE: i = n + m; j = s; break;
    case ES: i = s; j = 1 - m; break;
    default: i = s; j = n + m; break;
  }
}

enum KClass { KC_ADVECT = 0, KC_REMAP = 1, KC_HALO = 2, KC_CMAX = 3, KC_SCALE = 4, KC_N = 5 };

template <class T> struct Impl {
  fv3t_dims d{};
  int n = 0, npz = 0, nqmax = 0, nt = 0, device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  // device state
  T *q[2] = {nullptr, nullptr}, *dp1 = nullptr, *mfx = nullptr, *mfy = nullptr, *cx = nullptr, *cy = nullptr, *pe = nullptr,
    *delp = nullptr;
  T *area = nullptr, *rarea = nullptr, *dx = nullptr, *dy = nullptr, *dxa = nullptr, *dya = nullptr, *sin_sg = nullptr;
  T *ak = nullptr, *bk = nullptr, *cmax_t = nullptr;
  T *xfs = nullptr, *yfs = nullptr;  // strict mode: xfx, yfx of tracer_2d step A (scratch, extents of cx / cy; allocated on first use)
  // fast mode (fv3t_advect3.cuh): tracer-independent per-level scratch in plane layout, allocated on first use
  fv3t::Pair<T>*X2 = nullptr, *Y2 = nullptr, *cab = nullptr;
  T *rrx = nullptr, *rry = nullptr;
  // multi-tracer TMA-staged advection (fv3t_advect5.cuh): padded-pitch scratch planes + their tensor maps
  fv3t::Pair<T>*X5 = nullptr, *Y5 = nullptr, *C5 = nullptr;
  T *RX5 = nullptr, *RY5 = nullptr, *MX5 = nullptr, *MY5 = nullptr, *AREA5 = nullptr, *RAREA5 = nullptr;
  fv3t::Adv5Maps maps5;
  bool maps5_ok = false;
  bool use5 = true;            // FV3T_ADV5=0 keeps the per-tracer k_advect4
  bool call5 = false;          // the current tracer_2d call runs k_advect5
  bool mode_1l = false;        // the current call is tracer_2d_1L: the dp1 post-state of fv_tracer2d.F90:305
  bool exact5 = false;         // ... its exact-arithmetic instantiation (schemes outside fast_hord_ok: bit-identical to the oracle)
  bool scale_pending = false;  // k_advect5 path: the in-place 1/ksplt scaling of cx, cy, mfx, mfy is applied by finish()
  fv3t::Pair<T>* P1 = nullptr;  // fast remap: spline / overlap coefficients per column (fv3t_remap3.cuh)
  T *GAM = nullptr, *RD1 = nullptr, *R2 = nullptr;
  unsigned char* coef4 = nullptr;  // lanes-over-levels remap (fv3t_remap4.cuh): per-column-group coefficient blocks
  unsigned char* neg4 = nullptr;   // per column and tracer: the remapped column holds a negative value (k_fillz4)
  bool use4 = false;               // FV3T_REMAP4=1 selects the experimental lanes-over-levels kernel (kord 9; slower on B200)
  cudaStream_t side = nullptr;            // remap_prepare: the coefficient kernel runs here, concurrently with tracer_2d
  cudaEvent_t ev_fork = nullptr, ev_coef = nullptr;
  bool coef_ready = false, coef_wanted = false;
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;  // tracer_step: copy streams of the per-tracer pipeline
  std::vector<cudaEvent_t> ev_up, ev_done;
  bool fast = true;        // FV3T_STRICT=1 selects the bit-exact kernels for everything
  bool prep_done = true;   // steps A/C of the current tracer_2d call have been run (done lazily by the first sub-step)
  bool call_fast = false;  // the current tracer_2d call runs the fast kernels
  T ptop = T(0);
  int *ksplt_d = nullptr, *par_d = nullptr, *cpy_d = nullptr, *kord_d = nullptr, *halo_dst = nullptr, *halo_src = nullptr;
  int cur = 0;  // outside tracer_2d every level of q lives in q[cur]; sub-step `it` reads q[(cur+it-1)&1] and writes the other
  int halo_len = 0;
  int* strip_idx[6][4] = {};  // per local tile / edge: device index lists (3n) for pack (src cells) and unpack (halo cells)
  int* strip_halo[6][4] = {};
  EdgeMap maps[6][4];
  int local_of[6];  // global tile (0-based) -> local slot or -1
  int nq_layout = 0;           // tracer count the resident q was laid out for by the last host upload (0: written through device pointers)
  int sub_L = 0;               // sub-tile context: layout L x L per tile (0: whole tiles); flags per resident sub-domain
  fv3t::A5Sub subflags[6];
  std::vector<int*> lists;     // gather / scatter lists (fv3t_halo_list_create)
  std::vector<int> list_len, list_max;
  // host state
  std::vector<int> ksplt, cpy;
  std::vector<T> cmax_h;
  int nsplt = 1, nq_cur = 0;
  bool have_vertical = false;
  uint64_t launches = 0;
  std::mutex row_mutex;
  T* row_buf = nullptr;  // staging for the row-granular mapn_tracer entry
  // tracer damping (deln_flux, fv3t_deln.cuh): fv_grid_type%del6_u / del6_v / da_min, the settings of the current call, scratch
  T *del6_u = nullptr, *del6_v = nullptr, *dfx2 = nullptr, *dfy2 = nullptr, *dd2 = nullptr;
  T da_min = T(0), damp_trdm = T(0);
  int damp_nord = 0;
  T* fld = nullptr;        // device staging of one scalar field (map_scalar / map1_ppm entries), sized like delp
  T* fld_qs = nullptr;     // its bottom boundary values (iv = -2)
  T* strip_buf = nullptr;  // device staging of one packed edge strip (halo_pack_host / halo_unpack_host)
  size_t strip_cap = 0;
  // timing
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool prof = false;
  float prof_ms[KC_N] = {};
  int prof_n[KC_N] = {};
  cudaEvent_t pe0 = nullptr, pe1 = nullptr;

  size_t plane() const { return (size_t)(n + 6) * (n + 6); }
  size_t sz_q(int nq) const { return plane() * npz * nq; }
  size_t sz_c() const { return plane() * npz; }
  size_t sz_cx() const { return (size_t)(n + 1) * (n + 6) * npz; }
  size_t sz_mf() const { return (size_t)(n + 1) * n * npz; }
  size_t sz_pe() const { return (size_t)(n + 2) * (npz + 1) * (n + 2); }
  size_t field_elems(int f, int nq) const {
    switch (f) {
      case FV3T_Q: return sz_q(nq) * nt;
      case FV3T_DP1:
      case FV3T_DELP: return sz_c() * nt;
      case FV3T_MFX:
      case FV3T_MFY: return sz_mf() * nt;
      case FV3T_CX:
      case FV3T_CY: return sz_cx() * nt;
      case FV3T_PE: return sz_pe() * nt;
    }
    return 0;
  }
  T* field_ptr(int f) {
    switch (f) {
      case FV3T_Q: return q[cur];
      case FV3T_DP1: return dp1;
      case FV3T_MFX: return mfx;
      case FV3T_MFY: return mfy;
      case FV3T_CX: return cx;
      case FV3T_CY: return cy;
      case FV3T_PE: return pe;
      case FV3T_DELP: return delp;
    }
    return nullptr;
  }

  void kbegin() {
    if (prof) cudaEventRecord(pe0, stream);
  }
  void kend(int kc) {
    ++launches;
    if (prof) {
      cudaEventRecord(pe1, stream);
      cudaEventSynchronize(pe1);
      float ms = 0;
      cudaEventElapsedTime(&ms, pe0, pe1);
      prof_ms[kc] += ms;
      prof_n[kc] += 1;
    }
  }

  int create(const fv3t_dims* dims, const T* const* g /*7 ptrs*/, int dev, void* strm);
  int destroy();
  int upload(int f, const T* h, int nq);
  int download(int f, T* h, int nq);
  int begin(int nq, int q_split, T* cmax_local);
  int set_cmax(const T* cmax_global, int q_split, int* nsplt_out);
  int halo_local(int it);
  int halo_pack(int it, int lt, int edge, T* buf, bool unpack);
  int substep(int it, int hord, T lim_fac);
  int prepare(int hord, bool allow5 = true);
  int apply_damping(int it, bool mf_scaled);
  int halo_list_move(int it, int lt, int list, T* buf, bool scatter, int stride);
  int halo_list_create(const int* offs, int count, int* list);
  int halo_local_table(const int* dst, const int* src, int len);
  int fv_tp_2d_host(int nlev, T* q, const T* crx, const T* cry, int hord, T* fx, T* fy, const T* xfx, const T* yfx, const T* ra_x,
                    const T* ra_y, T lim_fac, const T* mfx, const T* mfy, const T* mass, int nord, T damp_c);
  int alloc5();
  int finish();
  int tracer_2d_resident(int nq, int hord, int q_split, T lim_fac, int* nsplt_out);
  int remap_resident(int nq, const int* kord, int fill, int j_first, int j_count, const T* pe2_ext = nullptr, const T* dp2_ext = nullptr);
  int remap_prepare();
  int launch_coef_side();
  int tracer_step(T* hq, T* hdp1, T* hmfx, T* hmfy, T* hcx, T* hcy, const T* hpe, const T* hak, const T* hbk, T hptop, T* hdelp, int nq,
                  int hord, int q_split, T lim_fac, const int* kord, int fill, int* nsplt_out);
  int remap_alloc();
  bool remap4_ok() const { return fast && use4 && npz <= 127; }  // (and abs(kord) == 9, checked by the caller)
};

template <class T> int Impl<T>::create(const fv3t_dims* dims, const T* const* g, int dev, void* strm) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("fv3tracer: no CUDA device available (there is no CPU fallback)");
  if (dev < 0 || dev >= ndev) return fail("fv3tracer: device %d out of range (%d devices)", dev, ndev);
  d = *dims;
  n = dims->npx - 1;
  npz = dims->npz;
  nqmax = dims->nq_max;
  nt = dims->ntiles;
  device = dev;
  if (n < 8) return fail("fv3tracer: npx-1 = %d too small (need >= 8 cells per tile edge)", n);
  if (npz < 6 || npz > 128) return fail("fv3tracer: npz = %d outside the supported range 6..128", npz);
  if (nt < 1 || nt > 6) return fail("fv3tracer: ntiles = %d outside 1..6", nt);
  if (nqmax < 1) return fail("fv3tracer: nq_max must be >= 1");
  for (int t = 0; t < 6; ++t) local_of[t] = -1;
  sub_L = dims->sub_layout >= 2 ? dims->sub_layout : 0;
  for (int s = 0; s < nt; ++s) {
    const int gt = dims->tile_id[s] - 1;
    if (gt < 0 || gt > 5 || (!sub_L && local_of[gt] >= 0)) return fail("fv3tracer: bad tile_id[%d] = %d", s, dims->tile_id[s]);
    local_of[gt] = s;
    if (sub_L) {  // which sides of the resident square lie on a tile edge, which of its corners are cube corners (A5Sub)
      const int bi = dims->sub_bi[s], bj = dims->sub_bj[s];
      if (bi < 0 || bi >= sub_L || bj < 0 || bj >= sub_L) return fail("fv3tracer: sub-domain (%d, %d) outside the %d x %d layout", bi, bj, sub_L, sub_L);
      const bool w = bi == 0, e = bi == sub_L - 1, so = bj == 0, no = bj == sub_L - 1;
      subflags[s] = fv3t::A5Sub{!w, !e, !so, !no, (w && so ? 1 : 0) | (e && so ? 2 : 0) | (e && no ? 4 : 0) | (w && no ? 8 : 0)};
    }
  }
  CK(cudaSetDevice(dev));
  if (strm) {
    stream = (cudaStream_t)strm;
  } else {
    CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    own_stream = true;
  }
  build_edge_maps(n, maps);
  auto dalloc = [&](T** p, size_t elems) -> cudaError_t { return cudaMalloc((void**)p, elems * sizeof(T)); };
  CK(dalloc(&q[0], sz_q(nqmax) * nt));
  CK(dalloc(&q[1], sz_q(nqmax) * nt));
  CK(dalloc(&dp1, sz_c() * nt));
  CK(dalloc(&delp, sz_c() * nt));
  CK(dalloc(&mfx, sz_mf() * nt));
  CK(dalloc(&mfy, sz_mf() * nt));
  CK(dalloc(&cx, sz_cx() * nt));
  CK(dalloc(&cy, sz_cx() * nt));
  CK(dalloc(&pe, sz_pe() * nt));
  fast = !(getenv("FV3T_STRICT") && atoi(getenv("FV3T_STRICT")) != 0);
  use5 = !(getenv("FV3T_ADV5") && atoi(getenv("FV3T_ADV5")) == 0);
  use4 = getenv("FV3T_REMAP4") && atoi(getenv("FV3T_REMAP4")) != 0;
  CK(cudaMemsetAsync(q[0], 0, sz_q(nqmax) * nt * sizeof(T), stream));
  CK(cudaMemsetAsync(q[1], 0, sz_q(nqmax) * nt * sizeof(T), stream));
  CK(cudaMemsetAsync(delp, 0, sz_c() * nt * sizeof(T), stream));
  const size_t pl = plane();
  const size_t gsz[7] = {pl, pl, (size_t)(n + 6) * (n + 7), (size_t)(n + 7) * (n + 6), pl, pl, pl * 5};
  T** gp[7] = {&area, &rarea, &dx, &dy, &dxa, &dya, &sin_sg};
  for (int a = 0; a < 7; ++a) {
    CK(dalloc(gp[a], gsz[a] * nt));
    CK(cudaMemcpyAsync(*gp[a], g[a], gsz[a] * nt * sizeof(T), cudaMemcpyHostToDevice, stream));
  }
  CK(dalloc(&ak, npz + 1));
  CK(dalloc(&bk, npz + 1));
  CK(dalloc(&cmax_t, (size_t)nt * npz));
  CK(cudaMalloc((void**)&ksplt_d, npz * sizeof(int)));
  CK(cudaMalloc((void**)&par_d, npz * sizeof(int)));
  CK(cudaMalloc((void**)&cpy_d, npz * sizeof(int)));
  CK(cudaMalloc((void**)&kord_d, nqmax * sizeof(int)));
  CK(cudaMemsetAsync(par_d, 0, npz * sizeof(int), stream));
  ksplt.assign(npz, 1);
  cpy.assign(npz, 0);
  cmax_h.assign((size_t)nt * npz, T(0));

  // halo tables: local gathers (both tiles resident) + per-edge strips for remote exchange
  std::vector<int> hd, hs;
  const int nd = n + 6;
  for (int s = 0; s < nt && !sub_L; ++s) {  // (a sub-tile context gets its tables from the host: fv3t_halo_local_table, halo lists)
    const int gt = dims->tile_id[s] - 1;
    for (int e = 0; e < 4; ++e) {
      const EdgeMap& m = maps[gt][e];
      std::vector<int> my_halo(3 * n), my_src(3 * n);
      // cells of THIS tile that the neighbour across edge e needs, in the neighbour's canonical halo order
      const EdgeMap& back = maps[m.nbr_tile][m.nbr_edge];
      for (int mm = 1; mm <= 3; ++mm)
        for (int ss = 1; ss <= n; ++ss) {
          int i, j;
          halo_cell(e, n, mm, ss, i, j);
          my_halo[(mm - 1) * n + (ss - 1)] = (j + 2) * nd + (i + 2);
          const int ip = m.A[0][0] * i + m.A[0][1] * j + m.b[0];
          const int jp = m.A[1][0] * i + m.A[1][1] * j + m.b[1];
          if (local_of[m.nbr_tile] >= 0) {
            hd.push_back((s * nd + (j + 2)) * nd + (i + 2));
            hs.push_back((local_of[m.nbr_tile] * nd + (jp + 2)) * nd + (ip + 2));
          }
          int bi, bj;
          halo_cell(m.nbr_edge, n, mm, ss, bi, bj);
          const int si = back.A[0][0] * bi + back.A[0][1] * bj + back.b[0];
          const int sj = back.A[1][0] * bi + back.A[1][1] * bj + back.b[1];
          my_src[(mm - 1) * n + (ss - 1)] = (sj + 2) * nd + (si + 2);
        }
      CK(cudaMalloc((void**)&strip_halo[s][e], 3 * n * sizeof(int)));
      CK(cudaMalloc((void**)&strip_idx[s][e], 3 * n * sizeof(int)));
      CK(cudaMemcpyAsync(strip_halo[s][e], my_halo.data(), 3 * n * sizeof(int), cudaMemcpyHostToDevice, stream));
      CK(cudaMemcpyAsync(strip_idx[s][e], my_src.data(), 3 * n * sizeof(int), cudaMemcpyHostToDevice, stream));
      CK(cudaStreamSynchronize(stream));  // vectors go out of scope
    }
  }
  halo_len = (int)hd.size();
  if (halo_len) {
    CK(cudaMalloc((void**)&halo_dst, halo_len * sizeof(int)));
    CK(cudaMalloc((void**)&halo_src, halo_len * sizeof(int)));
    CK(cudaMemcpyAsync(halo_dst, hd.data(), halo_len * sizeof(int), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(halo_src, hs.data(), halo_len * sizeof(int), cudaMemcpyHostToDevice, stream));
  }
  CK(cudaEventCreate(&ev0));
  CK(cudaEventCreate(&ev1));
  CK(cudaEventCreate(&pe0));
  CK(cudaEventCreate(&pe1));
  CK(cudaStreamSynchronize(stream));
  return 0;
}

template <class T> int Impl<T>::destroy() {
  cudaSetDevice(device);
  cudaStreamSynchronize(stream);
  void* ptrs[] = {q[0], q[1], xfs, yfs, X2, Y2, cab, rrx, rry, X5, Y5, C5, RX5, RY5, MX5, MY5, AREA5, RAREA5, coef4, neg4, P1, GAM, RD1, R2, dp1, mfx, mfy, cx, cy, pe, delp, area, rarea, dx, dy, dxa, dya, sin_sg, ak, bk, cmax_t,
                  ksplt_d, par_d, cpy_d, kord_d, halo_dst, halo_src, row_buf, strip_buf, fld, fld_qs, del6_u, del6_v, dfx2, dfy2, dd2};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (int* l : lists)
    if (l) cudaFree(l);
  for (int s = 0; s < 6; ++s)
    for (int e = 0; e < 4; ++e) {
      if (strip_idx[s][e]) cudaFree(strip_idx[s][e]);
      if (strip_halo[s][e]) cudaFree(strip_halo[s][e]);
    }
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
  if (pe0) cudaEventDestroy(pe0);
  if (pe1) cudaEventDestroy(pe1);
  if (side) cudaStreamDestroy(side);
  if (s_h2d) cudaStreamDestroy(s_h2d);
  if (s_d2h) cudaStreamDestroy(s_d2h);
  for (cudaEvent_t e : ev_up) cudaEventDestroy(e);
  for (cudaEvent_t e : ev_done) cudaEventDestroy(e);
  if (ev_fork) cudaEventDestroy(ev_fork);
  if (ev_coef) cudaEventDestroy(ev_coef);
  if (own_stream) cudaStreamDestroy(stream);
  return 0;
}

template <class T> int Impl<T>::upload(int f, const T* h, int nq) {
  T* dptr = field_ptr(f);
  if (!dptr) return fail("fv3tracer: unknown field %d", f);
  if (f == FV3T_Q && (nq < 1 || nq > nqmax)) return fail("fv3tracer: nq = %d outside 1..nq_max = %d", nq, nqmax);
  CK(cudaSetDevice(device));
  CK(cudaMemcpyAsync(dptr, h, field_elems(f, nq) * sizeof(T), cudaMemcpyHostToDevice, stream));
  if (f == FV3T_Q) nq_cur = nq_layout = nq;
  if (f == FV3T_PE) coef_ready = coef_wanted = false;
  return 0;
}

template <class T> int Impl<T>::download(int f, T* h, int nq) {
  T* dptr = field_ptr(f);
  if (!dptr) return fail("fv3tracer: unknown field %d", f);
  CK(cudaSetDevice(device));
  if (f == FV3T_Q && (nq < 1 || nq > nqmax)) return fail("fv3tracer: nq = %d outside 1..nq_max = %d", nq, nqmax);
  CK(cudaMemcpyAsync(h, dptr, field_elems(f, nq) * sizeof(T), cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  return 0;
}

// steps A-B up to the local cmax (fv_tracer2d.F90:387-427); xfx/yfx are not materialised (see fv3t_advect.cuh)
template <class T> int Impl<T>::begin(int nq, int q_split, T* cmax_local) {
  if (nq < 1 || nq > nqmax) return fail("fv3tracer: nq = %d outside 1..nq_max = %d", nq, nqmax);
  // nq is the tile stride of q: a count other than the one the resident tracers were uploaded with would address other planes
  if (nq_layout && nq != nq_layout) return fail("fv3tracer: nq = %d, but the resident tracers were uploaded as %d per tile", nq, nq_layout);
  CK(cudaSetDevice(device));
  nq_cur = nq;
  if (q_split == 0) {
    kbegin();
    fv3t::k_cmax<T><<<nt * npz, 512, 0, stream>>>(cx, cy, sin_sg, cmax_t, n, npz);
    kend(KC_CMAX);
    CK(cudaMemcpyAsync(cmax_h.data(), cmax_t, sizeof(T) * nt * npz, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (cmax_local) {
      for (int k = 0; k < npz; ++k) {
        T c = cmax_h[k];
        for (int t = 1; t < nt; ++t) c = std::max(c, cmax_h[(size_t)t * npz + k]);
        cmax_local[k] = c;
      }
    }
  } else if (cmax_local) {
    for (int k = 0; k < npz; ++k) cmax_local[k] = T(0);
  }
  CK(cudaGetLastError());
  return 0;
}

// mp_reduce_max result -> nsplt, ksplt (fv_tracer2d.F90:432-457).  q_split /= 0: the reference reads an
// unset cmax for ksplt; defined here as ksplt(k) = nsplt.
template <class T> int Impl<T>::set_cmax(const T* cmax_global, int q_split, int* nsplt_out) {
  CK(cudaSetDevice(device));
  if (q_split == 0) {
    T c_global = cmax_global[0];
    if (npz != 1)
      for (int k = 1; k < npz; ++k) c_global = std::max(cmax_global[k], c_global);
    nsplt = (int)(T(1) + c_global);
  } else {
    nsplt = q_split;
  }
  for (int k = 0; k < npz; ++k) ksplt[k] = 1;
  if (nsplt != 1)
    for (int k = 0; k < npz; ++k) ksplt[k] = (q_split == 0) ? (int)(T(1) + cmax_global[k]) : nsplt;
  CK(cudaMemcpyAsync(ksplt_d, ksplt.data(), npz * sizeof(int), cudaMemcpyHostToDevice, stream));
  CK(cudaStreamSynchronize(stream));
  prep_done = false;  // steps A and C run with the first sub-step, when hord (strict / fast kernels) is known
  if (nsplt_out) *nsplt_out = nsplt;
  return 0;
}

template <class T> int Impl<T>::halo_local(int it) {
  if (!halo_len) return 0;
  CK(cudaSetDevice(device));
  dim3 grid((halo_len + 255) / 256, nq_cur * npz);
  kbegin();
  fv3t::k_halo_fill<T><<<grid, 256, 0, stream>>>(q[(cur + it - 1) & 1], halo_dst, halo_src, halo_len, n, npz, nq_cur, ksplt_d, it);
  kend(KC_HALO);
  CK(cudaGetLastError());
  return 0;
}

template <class T>
__global__ void k_strip(T* __restrict__ q, T* __restrict__ buf, const int* __restrict__ idx, int n, int npz, int nq,
                        const int* __restrict__ ksplt, int it, int unpack, int len, int stride) {
  const long plane = (long)(n + 6) * (n + 6);
  const int pl = blockIdx.y;  // iq*npz + kz
  if (it > ksplt[pl % npz]) return;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < len; e += gridDim.x * blockDim.x) {
    if (unpack)
      q[(long)pl * plane + idx[e]] = buf[(long)pl * stride + e];
    else
      buf[(long)pl * stride + e] = q[(long)pl * plane + idx[e]];
  }
}

// the same for a list of FLAT offsets into the stack of resident tiles (local_tile * plane + offset): one launch moves the cells
// of several resident tiles
template <class T>
__global__ void k_strip_flat(T* __restrict__ q, T* __restrict__ buf, const int* __restrict__ idx, int n, int npz, int nq,
                             const int* __restrict__ ksplt, int it, int unpack, int len, int stride) {
  const int plane = (n + 6) * (n + 6);
  const long tile_stride = (long)plane * npz * nq;
  const int pl = blockIdx.y;
  if (it > ksplt[pl % npz]) return;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < len; e += gridDim.x * blockDim.x) {
    const int o = idx[e];
    const long a = (long)(o / plane) * tile_stride + (long)pl * plane + o % plane;
    if (unpack)
      q[a] = buf[(long)pl * stride + e];
    else
      buf[(long)pl * stride + e] = q[a];
  }
}

template <class T> int Impl<T>::halo_pack(int it, int lt, int edge, T* buf, bool unpack) {
  if (lt < 0 || lt >= nt || edge < 0 || edge > 3) return fail("fv3tracer: bad tile/edge %d/%d", lt, edge);
  if (sub_L) return fail("fv3tracer: the edge-strip exchange is for whole-tile contexts; sub-tile contexts use the halo lists");
  CK(cudaSetDevice(device));
  dim3 grid((3 * n + 255) / 256, nq_cur * npz);
  T* qt = q[(cur + it - 1) & 1] + (size_t)lt * sz_q(nq_cur);
  kbegin();
  k_strip<T><<<grid, 256, 0, stream>>>(qt, buf, unpack ? strip_halo[lt][edge] : strip_idx[lt][edge], n, npz, nq_cur, ksplt_d, it,
                                       unpack ? 1 : 0, 3 * n, 3 * n);
  kend(KC_HALO);
  CK(cudaGetLastError());
  return 0;
}

template <class T> int Impl<T>::halo_list_create(const int* offs, int count, int* list) {
  if (!list || count < 0 || (count > 0 && !offs)) return fail("fv3tracer: halo_list_create: bad arguments");
  const long pl = (long)plane() * nt;
  int mx = 0;
  for (int e = 0; e < count; ++e) {
    if (offs[e] < 0 || offs[e] >= pl) return fail("fv3tracer: halo_list_create: offset %d outside the resident planes (%ld cells)", offs[e], pl);
    mx = std::max(mx, offs[e]);
  }
  CK(cudaSetDevice(device));
  int* d = nullptr;
  if (count) {
    CK(cudaMalloc((void**)&d, (size_t)count * sizeof(int)));
    CK(cudaMemcpy(d, offs, (size_t)count * sizeof(int), cudaMemcpyHostToDevice));
  }
  lists.push_back(d);
  list_len.push_back(count);
  list_max.push_back(mx);
  *list = (int)lists.size() - 1;
  return 0;
}

template <class T> int Impl<T>::halo_local_table(const int* dst, const int* src, int len) {
  if (len < 0 || (len > 0 && (!dst || !src))) return fail("fv3tracer: halo_local_table: bad arguments");
  const long total = (long)plane() * nt;
  for (int e = 0; e < len; ++e)
    if (dst[e] < 0 || dst[e] >= total || src[e] < 0 || src[e] >= total) return fail("fv3tracer: halo_local_table: offset outside the resident planes");
  CK(cudaSetDevice(device));
  CK(cudaStreamSynchronize(stream));
  if (halo_dst) cudaFree(halo_dst);
  if (halo_src) cudaFree(halo_src);
  halo_dst = halo_src = nullptr;
  halo_len = len;
  if (len) {
    CK(cudaMalloc((void**)&halo_dst, (size_t)len * sizeof(int)));
    CK(cudaMalloc((void**)&halo_src, (size_t)len * sizeof(int)));
    CK(cudaMemcpy(halo_dst, dst, (size_t)len * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(halo_src, src, (size_t)len * sizeof(int), cudaMemcpyHostToDevice));
  }
  return 0;
}

// generic gather (scatter) of the cells named by a registered list, every plane of the resident tracers of one resident tile
template <class T> int Impl<T>::halo_list_move(int it, int lt, int list, T* buf, bool scatter, int stride) {
  if (lt < -1 || lt >= nt) return fail("fv3tracer: bad local tile %d", lt);
  if (list < 0 || list >= (int)lists.size()) return fail("fv3tracer: unknown halo list %d", list);
  if (lt >= 0 && list_max[list] >= (int)plane()) return fail("fv3tracer: halo list %d holds flat offsets: use local_tile = -1", list);
  if (!buf) return fail("fv3tracer: halo_gather / halo_scatter: null buffer");
  if (nq_cur < 1) return fail("fv3tracer: no resident tracers");
  const int len = list_len[list];
  if (len == 0) return 0;
  if (stride == 0) stride = len;
  if (stride < len) return fail("fv3tracer: halo_gather / halo_scatter: buffer stride %d shorter than the list (%d)", stride, len);
  CK(cudaSetDevice(device));
  dim3 grid((len + 255) / 256, nq_cur * npz);
  T* qt = q[(cur + it - 1) & 1] + (size_t)(lt < 0 ? 0 : lt) * sz_q(nq_cur);
  kbegin();
  if (lt < 0)
    k_strip_flat<T><<<grid, 256, 0, stream>>>(qt, buf, lists[list], n, npz, nq_cur, ksplt_d, it, scatter ? 1 : 0, len, stride);
  else
    k_strip<T><<<grid, 256, 0, stream>>>(qt, buf, lists[list], n, npz, nq_cur, ksplt_d, it, scatter ? 1 : 0, len, stride);
  kend(KC_HALO);
  CK(cudaGetLastError());
  return 0;
}

// strip width: threads = W + 6 (3 halo columns on either side); pick the block size that wastes the fewest lanes
inline int pick_block(int n) {
  const int forced = getenv("FV3T_ADV_NT") ? atoi(getenv("FV3T_ADV_NT")) : 0;  // test / tuning knob
  if (forced >= 32 && forced <= 256 && forced % 32 == 0) return forced;
  int best = 32;
  double best_eff = -1.0;
  for (int nt = 32; nt <= 256; nt += 32) {
    const int w = nt - 6;
    const int strips = (n + w - 1) / w;
    const double eff = (double)n / ((double)strips * nt);
    if (eff > best_eff - 1e-9) {  // ties go to the larger block (fewer strips, fewer redundant halo columns per SM)
      best_eff = eff;
      best = nt;
    }
  }
  return best;
}

template <class T, int OI, int OO> int launch_advect(Impl<T>& c, fv3t::Adv2Params<T>& p) {
  const int NT = pick_block(c.n);
  p.W = NT - 6;
  const int strips = (c.n + p.W - 1) / p.W;
  dim3 grid(p.nq, strips, c.nt * c.npz);
  const size_t smem = (size_t)6 * NT * sizeof(T);
  c.kbegin();
  fv3t::k_advect2<T, OI, OO><<<grid, NT, smem, c.stream>>>(p);
  c.kend(KC_ADVECT);
  CK(cudaGetLastError());
  return 0;
}

// steps A and C of tracer_2d: xfx, yfx and the in-place 1/ksplt scaling of cx, cy, mfx, mfy (fv_tracer2d.F90:387-405,
// 449-486); the fast path additionally prepares 1/ra_x, 1/ra_y and the dp1/dp2 factors (fv3t_advect3.cuh, k_prep3)
template <class T> int Impl<T>::alloc5() {
  if (maps5_ok) return 0;
  const int nd = n + 6, PP = fv3t::a5_pitch(n);
  const size_t e = (size_t)nt * npz * nd * PP;
  auto zalloc = [&](void** p, size_t bytes) -> cudaError_t {
    if (*p) return cudaSuccess;
    cudaError_t r = cudaMalloc(p, bytes);
    if (r != cudaSuccess) return r;
    return cudaMemsetAsync(*p, 0, bytes, stream);  // the padding columns and the cells outside the valid faces stay zero
  };
  CK(zalloc((void**)&X5, e * sizeof(fv3t::Pair<T>)));
  CK(zalloc((void**)&Y5, e * sizeof(fv3t::Pair<T>)));
  CK(zalloc((void**)&C5, e * sizeof(fv3t::Pair<T>)));
  CK(zalloc((void**)&RX5, e * sizeof(T)));
  CK(zalloc((void**)&RY5, e * sizeof(T)));
  CK(zalloc((void**)&MX5, e * sizeof(T)));
  CK(zalloc((void**)&MY5, e * sizeof(T)));
  CK(zalloc((void**)&AREA5, (size_t)nt * nd * PP * sizeof(T)));
  CK(zalloc((void**)&RAREA5, (size_t)nt * nd * PP * sizeof(T)));
  CK(fv3t::fast_pad_plane<T>(AREA5, area, nd, PP, nt, stream));
  CK(fv3t::fast_pad_plane<T>(RAREA5, rarea, nd, PP, nt, stream));
  launches += 2;
  fv3t::Adv5Params<T> p{};
  p.X2 = X5;
  p.Y2 = Y5;
  p.CAB = C5;
  p.RX = RX5;
  p.RY = RY5;
  p.MFX = MX5;
  p.MFY = MY5;
  p.AREA = AREA5;
  p.RAREA = RAREA5;
  p.n = n;
  p.npz = npz;
  p.ntiles = nt;
  const cudaError_t r = fv3t::fast_advect5_maps<T>(&maps5, p, nt * npz);
  if (r != cudaSuccess) return fail("fv3tracer: cuTensorMapEncodeTiled failed (%s)", cudaGetErrorString(r));
  maps5_ok = true;
  return 0;
}

template <class T> int Impl<T>::prepare(int hord, bool allow5) {
  call_fast = fast && fv3t::fast_hord_ok(hord);
  // k_advect5 shares the staged level fields among the tracers of a CTA and takes one CTA per SM: with fewer than four resident
  // tracers (small tracer groups of a sharded run) the per-tracer CTAs of k_advect4 / k_advect2 keep more warps in flight.
  // Schemes outside fast_hord_ok run its exact-arithmetic instantiation (FV3T_STRICT=1 keeps everything on k_advect2).
  call5 = fast && use5 && allow5 && fv3t::adv5_hord_ok(hord) && (nq_cur >= 4 || sub_L);
  if (sub_L && !call5)
    return fail("fv3tracer: sub-tile contexts run k_advect5 only (not with FV3T_STRICT=1 / FV3T_ADV5=0, not through tracer_step)");
  exact5 = call5 && !call_fast;
  auto dalloc = [&](void** p, size_t bytes) -> cudaError_t { return *p ? cudaSuccess : cudaMalloc(p, bytes); };
  if (call5) {
    const int rc = alloc5();
    if (rc) return rc;
    fv3t::Prep5Params<T> pp{cx, cy, mfx, mfy, dp1, fv3t::GridDev<T>{area, rarea, dx, dy, dxa, dya, sin_sg}, X5, Y5, C5, RX5, RY5, MX5, MY5,
                            ksplt_d, n, npz, nt, 0, nt * npz, 1, 1, exact5 ? 1 : 0};
    kbegin();
    CK(fv3t::fast_prep5<T>(pp, stream));
    kend(KC_SCALE);
    scale_pending = nsplt != 1;
  } else if (call_fast) {
    const size_t e = sz_c() * nt;
    CK(dalloc((void**)&X2, e * sizeof(fv3t::Pair<T>)));
    CK(dalloc((void**)&Y2, e * sizeof(fv3t::Pair<T>)));
    CK(dalloc((void**)&cab, e * sizeof(fv3t::Pair<T>)));
    CK(dalloc((void**)&rrx, e * sizeof(T)));
    CK(dalloc((void**)&rry, e * sizeof(T)));
    fv3t::Prep3Params<T> pp{cx, cy, mfx, mfy, dp1, fv3t::GridDev<T>{area, rarea, dx, dy, dxa, dya, sin_sg}, X2, Y2, cab, rrx, rry,
                            ksplt_d, n, npz, nt};
    kbegin();
    CK(fv3t::fast_prep3<T>(pp, stream));
    kend(KC_SCALE);
    if (nsplt != 1) {
      kbegin();
      CK(fv3t::fast_scale3<T>(cx, cy, mfx, mfy, ksplt_d, n, npz, nt, stream));
      kend(KC_SCALE);
    }
  } else {
    CK(dalloc((void**)&xfs, sz_cx() * nt * sizeof(T)));
    CK(dalloc((void**)&yfs, sz_cx() * nt * sizeof(T)));
    dim3 grid(16, nt * npz);
    kbegin();
    fv3t::k_prep<T><<<grid, 256, 0, stream>>>(cx, cy, mfx, mfy, xfs, yfs, fv3t::GridDev<T>{area, rarea, dx, dy, dxa, dya, sin_sg},
                                              ksplt_d, n, npz, nt, nsplt != 1 ? 1 : 0);
    kend(KC_SCALE);
    CK(cudaGetLastError());
  }
  prep_done = true;
  return 0;
}

// one pass of the `it` loop body (fv_tracer2d.F90:503-556) for the resident tiles
template <class T> int Impl<T>::substep(int it, int hord, T lim_fac) {
  CK(cudaSetDevice(device));
  if (!prep_done) {
    const int rc = prepare(hord);
    if (rc) return rc;
  }
  if (call5) {
    if (!fv3t::adv5_hord_ok(hord)) return fail("fv3tracer: hord_tr changed between the sub-steps of one tracer_2d call");
    if (it > 1) {  // dp1 <- dp2 of sub-step it-1 (fv_tracer2d.F90:547-553), then dp1/dp2, 0.5*rarea/dp2 of this sub-step
      fv3t::Prep5Params<T> pp{cx, cy, mfx, mfy, dp1, fv3t::GridDev<T>{area, rarea, dx, dy, dxa, dya, sin_sg}, X5, Y5, C5, RX5, RY5, MX5, MY5,
                              ksplt_d, n, npz, nt, 0, nt * npz, it, 0, exact5 ? 1 : 0, mode_1l ? 1 : 0};
      kbegin();
      CK(fv3t::fast_prep5<T>(pp, stream));
      kend(KC_SCALE);
    }
    fv3t::Adv5Params<T> p{};
    p.qin = q[(cur + it - 1) & 1];
    p.qout = q[(cur + it) & 1];
    p.X2 = X5;
    p.Y2 = Y5;
    p.CAB = C5;
    p.RX = RX5;
    p.RY = RY5;
    p.MFX = MX5;
    p.MFY = MY5;
    p.AREA = AREA5;
    p.RAREA = RAREA5;
    p.dxa = dxa;
    p.dya = dya;
    p.ksplt = ksplt_d;
    p.n = n;
    p.npz = npz;
    p.nq = nq_cur;
    p.ntiles = nt;
    p.it = it;
    p.lev0 = 0;
    p.iq0 = 0;
    p.nql = nq_cur;
    p.lim_fac = lim_fac;
    if (coef_wanted && it == 1 && !prof) {
      const int rcs = launch_coef_side();
      if (rcs) return rcs;
    }
    kbegin();
    if (sub_L) {  // the instantiations that look at the edge / corner flags of the resident sub-domains
      fv3t::Adv5ParamsSub<T> ps{};
      static_cast<fv3t::Adv5Params<T>&>(ps) = p;
      for (int t = 0; t < nt; ++t) ps.sub[t] = subflags[t];
      if (exact5)
        CK(fv3t::exact_advect5_sub<T>(ps, maps5, hord, nt * npz, stream));
      else
        CK(fv3t::fast_advect5_sub<T>(ps, maps5, hord, nt * npz, stream));
    } else if (exact5) {
      CK(fv3t::exact_advect5<T>(p, maps5, hord, nt * npz, stream));
    } else {
      CK(fv3t::fast_advect5<T>(p, maps5, hord, nt * npz, stream));
    }
    kend(KC_ADVECT);
    return apply_damping(it, false);  // mfx, mfy are scaled by finish() on this path
  }
  if (call_fast) {
    if (!fv3t::fast_hord_ok(hord)) return fail("fv3tracer: hord_tr changed between the sub-steps of one tracer_2d call");
    if (it > 1) {  // dp1 <- dp2 of sub-step it-1 (fv_tracer2d.F90:547-553), then dp1/dp2, 0.5*rarea/dp2 of this sub-step
      fv3t::Cab3Params<T> cp{dp1, mfx, mfy, rarea, cab, ksplt_d, n, npz, it, mode_1l ? 1 : 0};
      kbegin();
      CK(fv3t::fast_cab3<T>(cp, nt, stream));
      kend(KC_SCALE);
    }
    fv3t::Adv3Params<T> p;
    p.qin = q[(cur + it - 1) & 1];
    p.qout = q[(cur + it) & 1];
    p.X2 = X2;
    p.Y2 = Y2;
    p.rrx = rrx;
    p.rry = rry;
    p.cab = cab;
    p.mfx = mfx;
    p.mfy = mfy;
    p.area = area;
    p.dxa = dxa;
    p.dya = dya;
    p.ksplt = ksplt_d;
    p.n = n;
    p.npz = npz;
    p.nq = nq_cur;
    p.ntiles = nt;
    p.it = it;
    p.W = 0;
    p.lim_fac = lim_fac;
    if (coef_wanted && it == 1 && !prof) {
      const int rcs = launch_coef_side();
      if (rcs) return rcs;
    }
    kbegin();
    // 64-thread CTAs (58-column strips): measured fastest on B200 at C768 (131 ms vs 143 / 147 / 153 ms for 96 / 128 / 160
    // threads, profiles/r01_advect3_block_sweep.txt): three barriers per row step cost least when a CTA is two warps
    const int forced = getenv("FV3T_ADV_NT") ? atoi(getenv("FV3T_ADV_NT")) : 0;
    const int NT = (forced >= 32 && forced <= 256 && forced % 32 == 0) ? forced : 64;
    CK(fv3t::fast_advect3<T>(p, hord, NT, stream));
    kend(KC_ADVECT);
    return apply_damping(it, true);
  }
  fv3t::Adv2Params<T> p;
  p.qin = q[(cur + it - 1) & 1];
  p.qout = q[(cur + it) & 1];
  p.dp1 = dp1;
  p.mfx = mfx;
  p.mfy = mfy;
  p.cx = cx;
  p.cy = cy;
  p.xfs = xfs;
  p.yfs = yfs;
  p.g = fv3t::GridDev<T>{area, rarea, dx, dy, dxa, dya, sin_sg};
  p.ksplt = ksplt_d;
  p.n = n;
  p.npz = npz;
  p.nq = nq_cur;
  p.ntiles = nt;
  p.it = it;
  p.W = 0;
  p.lim_fac = lim_fac;
  int rc;
  switch (hord) {
    case 8: rc = launch_advect<T, 8, 8>(*this, p); break;
    case 10: rc = launch_advect<T, 8, 10>(*this, p); break;  // ord_in = 8 when hord == 10 (tp_core.F90:157-161)
    case 9: rc = launch_advect<T, 9, 9>(*this, p); break;
    case 7: rc = launch_advect<T, 7, 7>(*this, p); break;
    case 11: rc = launch_advect<T, 11, 11>(*this, p); break;
    case 12: rc = launch_advect<T, 12, 12>(*this, p); break;
    case 13: rc = launch_advect<T, 13, 13>(*this, p); break;
    case 5: rc = launch_advect<T, 5, 5>(*this, p); break;
    case -5: rc = launch_advect<T, -5, -5>(*this, p); break;
    case 6: rc = launch_advect<T, 6, 6>(*this, p); break;
    case 1: rc = launch_advect<T, 1, 1>(*this, p); break;
    case 2: rc = launch_advect<T, 2, 2>(*this, p); break;
    case 3: rc = launch_advect<T, 3, 3>(*this, p); break;
    case 4: rc = launch_advect<T, 4, 4>(*this, p); break;
    default: return fail("fv3tracer: hord_tr = %d is not a scheme of xppm/yppm", hord);
  }
  if (rc) return rc;
  if ((rc = apply_damping(it, true))) return rc;
  // dp1 <- dp2 between sub-steps (fv_tracer2d.F90:547-553; tests the GLOBAL nsplt)
  if (it != nsplt) {
    dim3 grid(8, nt * npz);
    kbegin();
    fv3t::k_dp1_update<T><<<grid, 256, 0, stream>>>(dp1, mfx, mfy, rarea, ksplt_d, n, npz, it, mode_1l ? 1 : 0);
    kend(KC_SCALE);
    CK(cudaGetLastError());
  }
  return 0;
}

// Tracer damping of the first sub-step (fv_tracer2d.F90:487-494, 527-532 -> tp_core.F90:229-234 -> deln_flux :1239-1387), applied
// to the field the advection kernel has just written; see fv3t_deln.cuh.
template <class T> int Impl<T>::apply_damping(int it, bool mf_scaled) {
  if (it != 1 || !(damp_trdm > T(1.e-4))) return 0;
  if (sub_L) return fail("fv3tracer: tracer damping is not available in sub-tile contexts");
  if (!del6_u) return fail("fv3tracer: tracer damping needs the damping metrics (fv3t_*_set_damping)");
  if (nt != 6) return fail("fv3tracer: tracer damping needs all six tiles resident (the dp1 halo is filled locally)");
  if (damp_nord < 0 || damp_nord > 2) return fail("fv3tracer: nord_tr = %d outside 0..2 (deln_flux needs nord + 1 <= ng halo cells)", damp_nord);
  const int nd = n + 6;
  const size_t planes = (size_t)nt * npz;
  auto dalloc = [&](void** p, size_t bytes) -> cudaError_t { return *p ? cudaSuccess : cudaMalloc(p, bytes); };
  CK(dalloc((void**)&dfx2, planes * nd * (nd + 1) * sizeof(T)));
  CK(dalloc((void**)&dfy2, planes * nd * (nd + 1) * sizeof(T)));
  CK(dalloc((void**)&dd2, planes * nd * nd * sizeof(T)));
  // mass = dp1 with its edge halos (complete_group_halo_update(dp1_pack), fv_tracer2d.F90:487-494)
  if (halo_len) {
    dim3 gh((halo_len + 255) / 256, npz);
    kbegin();
    fv3t::k_halo_fill<T><<<gh, 256, 0, stream>>>(dp1, halo_dst, halo_src, halo_len, n, npz, 1, ksplt_d, 1);
    kend(KC_HALO);
  }
  const T damp = (T)std::pow((double)(damp_trdm * da_min), (double)(damp_nord + 1));
  const T* qin = q[(cur + it - 1) & 1];
  T* qout = q[(cur + it) & 1];
  dim3 grid(64, (unsigned)planes);
  for (int iq = 0; iq < nq_cur; ++iq) {
    fv3t::DelnParams<T> p{qin + (size_t)iq * sz_c(), dfx2, dfy2, dd2, del6_u, del6_v, rarea, n, npz, damp_nord, damp_nord, 1,
                          (long)sz_q(nq_cur)};
    kbegin();
    fv3t::k_deln_flux<T><<<grid, 256, 0, stream>>>(p);
    kend(KC_ADVECT);
    for (int s = 1; s <= damp_nord; ++s) {
      p.nt = damp_nord - s;
      p.first = 0;
      p.src = dd2;
      p.src_tile_stride = (long)sz_c();
      kbegin();
      fv3t::k_deln_div<T><<<grid, 256, 0, stream>>>(p);
      kend(KC_ADVECT);
      kbegin();
      fv3t::k_deln_flux<T><<<grid, 256, 0, stream>>>(p);
      kend(KC_ADVECT);
    }
    fv3t::DelnApplyParams<T> a{qout + (size_t)iq * sz_c(), dfx2, dfy2, dp1, mfx, mfy, rarea, ksplt_d, (long)sz_q(nq_cur), n, npz,
                               mf_scaled ? 1 : 0, damp};
    kbegin();
    fv3t::k_deln_apply<T><<<grid, 256, 0, stream>>>(a);
    kend(KC_ADVECT);
  }
  CK(cudaGetLastError());
  return 0;
}

// ---- fv_tp_2d as an operator (tp_core.F90:110-249; fv3t_tp2d.cuh): host arrays in, fluxes out -----------------------------
template <class T> static cudaError_t tp2d_flux(int ord, const fv3t::Tp2dFlux<T>& p, dim3 grid, cudaStream_t st) {
  switch (ord) {
#define FV3T_TP2D_CASE(O) case O: fv3t::k_tp2d_flux<T, O><<<grid, 128, 0, st>>>(p); break;
    FV3T_TP2D_CASE(1) FV3T_TP2D_CASE(2) FV3T_TP2D_CASE(3) FV3T_TP2D_CASE(4) FV3T_TP2D_CASE(5) FV3T_TP2D_CASE(-5) FV3T_TP2D_CASE(6)
    FV3T_TP2D_CASE(7) FV3T_TP2D_CASE(8) FV3T_TP2D_CASE(9) FV3T_TP2D_CASE(10) FV3T_TP2D_CASE(11) FV3T_TP2D_CASE(12) FV3T_TP2D_CASE(13)
#undef FV3T_TP2D_CASE
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

template <class T>
int Impl<T>::fv_tp_2d_host(int nlev, T* q, const T* crx, const T* cry, int hord, T* fx, T* fy, const T* xfx, const T* yfx,
                           const T* ra_x, const T* ra_y, T lim_fac, const T* mfx_h, const T* mfy_h, const T* mass_h, int nord,
                           T damp_c) {
  if (!((hord >= 1 && hord <= 13) || hord == -5)) return fail("fv3tracer: hord = %d is not a scheme of xppm/yppm", hord);
  if (nlev < 1) return fail("fv3tracer: fv_tp_2d: nlev = %d", nlev);
  if (sub_L) return fail("fv3tracer: fv_tp_2d: not available in sub-tile contexts");
  if (!q || !crx || !cry || !fx || !fy || !xfx || !yfx || !ra_x || !ra_y) return fail("fv3tracer: fv_tp_2d: null array");
  if ((mfx_h == nullptr) != (mfy_h == nullptr)) return fail("fv3tracer: fv_tp_2d: mfx and mfy must be given together (tp_core.F90:209)");
  const bool tracer = mfx_h != nullptr;
  // tp_core.F90:229-234 (tracer branch: nord, damp_c AND mass) / :243-248 (nord, damp_c)
  const bool damp_on = nord >= 0 && damp_c > T(1.e-4) && (tracer ? mass_h != nullptr : true);
  const bool with_mass = tracer && damp_on;
  if (damp_on) {
    if (!del6_u) return fail("fv3tracer: fv_tp_2d: damping needs the damping metrics (fv3t_*_set_damping)");
    if (nord > 2) return fail("fv3tracer: fv_tp_2d: nord = %d outside 0..2 (deln_flux needs nord + 1 <= ng halo cells)", nord);
  }
  CK(cudaSetDevice(device));
  const size_t nd = n + 6, P = (size_t)nt * nlev;
  const size_t s_pl = P * nd * nd, s_xf = P * (n + 1) * nd, s_ra = P * n * nd, s_f = P * (n + 1) * n, s_dx = P * nd * (nd + 1);
  // one allocation, carved: q, q_mid, crx, xfx, fx2, cry, yfx, fy2, ra_x, ra_y, fx, fy [, mfx, mfy] [, mass] [, dfx2, dfy2, d2]
  size_t total = 2 * s_pl + 6 * s_xf + 2 * s_ra + 2 * s_f + (tracer ? 2 * s_f : 0) + (with_mass ? s_pl : 0) +
                 (damp_on ? 2 * s_dx + s_pl : 0);
  T* base = nullptr;
  CK(cudaMalloc((void**)&base, total * sizeof(T)));
  struct Free {
    T* p;
    ~Free() { cudaFree(p); }
  } guard{base};
  T* cur_p = base;
  auto take = [&](size_t ne) { T* r = cur_p; cur_p += ne; return r; };
  T *d_q = take(s_pl), *d_mid = take(s_pl), *d_crx = take(s_xf), *d_xfx = take(s_xf), *d_fx2 = take(s_xf), *d_cry = take(s_xf),
    *d_yfx = take(s_xf), *d_fy2 = take(s_xf), *d_rax = take(s_ra), *d_ray = take(s_ra), *d_fx = take(s_f), *d_fy = take(s_f);
  T *d_mfx = tracer ? take(s_f) : nullptr, *d_mfy = tracer ? take(s_f) : nullptr, *d_mass = with_mass ? take(s_pl) : nullptr;
  T *d_dfx = damp_on ? take(s_dx) : nullptr, *d_dfy = damp_on ? take(s_dx) : nullptr, *d_d2 = damp_on ? take(s_pl) : nullptr;
  auto up = [&](T* d, const T* h, size_t ne) { return cudaMemcpyAsync(d, h, ne * sizeof(T), cudaMemcpyHostToDevice, stream); };
  CK(up(d_q, q, s_pl));
  CK(up(d_crx, crx, s_xf));
  CK(up(d_xfx, xfx, s_xf));
  CK(up(d_cry, cry, s_xf));
  CK(up(d_yfx, yfx, s_xf));
  CK(up(d_rax, ra_x, s_ra));
  CK(up(d_ray, ra_y, s_ra));
  if (tracer) {
    CK(up(d_mfx, mfx_h, s_f));
    CK(up(d_mfy, mfy_h, s_f));
  }
  if (with_mass) CK(up(d_mass, mass_h, s_pl));

  const int ord_in = hord == 10 ? 8 : hord, ord_ou = hord;  // tp_core.F90:157-162
  const int nd_i = (int)nd;
  dim3 grid(64, (unsigned)P);
  const long pl_x = (long)(n + 1) * nd_i, pl_f = (long)(n + 1) * n;
  fv3t::Tp2dFlux<T> f{};
  f.n = n;
  f.nlev = nlev;
  f.lim_fac = lim_fac;
  // fy2 = yppm(q, dir-2 view), i = isd..ied
  f.src = d_q, f.cour = d_cry, f.flux = d_fy2, f.dxa = dya, f.c_plane = f.f_plane = pl_x, f.c_pitch = f.f_pitch = nd_i;
  f.c_l0 = f.f_l0 = -2, f.dir = 1, f.view = 2, f.l_lo = -2, f.l_hi = n + 3;
  CK(tp2d_flux<T>(ord_in, f, grid, stream));
  fv3t::Tp2dMid<T> m{d_q, d_fy2, d_yfx, area, d_ray, d_mid, n, nlev, 1};
  fv3t::k_tp2d_mid<T><<<grid, 256, 0, stream>>>(m);
  // fx = xppm(q_i), j = 1..n
  f.src = d_mid, f.cour = d_crx, f.flux = d_fx, f.dxa = dxa, f.c_plane = pl_x, f.f_plane = pl_f, f.c_pitch = f.f_pitch = n + 1;
  f.c_l0 = -2, f.f_l0 = 1, f.dir = 0, f.view = 0, f.l_lo = 1, f.l_hi = n;
  CK(tp2d_flux<T>(ord_ou, f, grid, stream));
  // fx2 = xppm(q, dir-1 view), j = jsd..jed
  f.src = d_q, f.flux = d_fx2, f.f_plane = pl_x, f.f_l0 = -2, f.view = 1, f.l_lo = -2, f.l_hi = n + 3;
  CK(tp2d_flux<T>(ord_in, f, grid, stream));
  m = fv3t::Tp2dMid<T>{d_q, d_fx2, d_xfx, area, d_rax, d_mid, n, nlev, 0};
  fv3t::k_tp2d_mid<T><<<grid, 256, 0, stream>>>(m);
  // fy = yppm(q_j), i = 1..n
  f.src = d_mid, f.cour = d_cry, f.flux = d_fy, f.dxa = dya, f.c_plane = pl_x, f.f_plane = pl_f, f.c_pitch = nd_i, f.f_pitch = n;
  f.c_l0 = -2, f.f_l0 = 1, f.dir = 1, f.view = 0, f.l_lo = 1, f.l_hi = n;
  CK(tp2d_flux<T>(ord_ou, f, grid, stream));
  fv3t::Tp2dAvg<T> av{d_fx, d_fy, d_fx2, d_fy2, tracer ? d_mfx : d_xfx, tracer ? d_mfy : d_yfx, n, tracer ? 1 : 0};
  fv3t::k_tp2d_avg<T><<<grid, 256, 0, stream>>>(av);
  if (damp_on) {
    const T damp = (T)std::pow((double)(damp_c * da_min), (double)(nord + 1));
    const T* operand = d_q;
    if (!with_mass) {
      fv3t::k_tp2d_scale<T><<<1184, 256, 0, stream>>>(d_d2, d_q, damp, (long)s_pl);
      operand = d_d2;
    }
    fv3t::DelnParams<T> dp{operand, d_dfx, d_dfy, d_d2, del6_u, del6_v, rarea, n, nlev, nord, nord, 1, (long)(nlev * nd * nd)};
    fv3t::k_deln_flux<T><<<grid, 256, 0, stream>>>(dp);
    for (int s = 1; s <= nord; ++s) {
      dp.nt = nord - s;
      dp.first = 0;
      dp.src = d_d2;
      fv3t::k_deln_div<T><<<grid, 256, 0, stream>>>(dp);
      fv3t::k_deln_flux<T><<<grid, 256, 0, stream>>>(dp);
    }
    fv3t::Tp2dDampAdd<T> da{d_fx, d_fy, d_dfx, d_dfy, with_mass ? d_mass : nullptr, damp, n};
    fv3t::k_tp2d_damp_add<T><<<grid, 256, 0, stream>>>(da);
  }
  fv3t::k_tp2d_corners<T><<<(unsigned)P, 64, 0, stream>>>(d_q, n);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(fx, d_fx, s_f * sizeof(T), cudaMemcpyDeviceToHost, stream));
  CK(cudaMemcpyAsync(fy, d_fy, s_f * sizeof(T), cudaMemcpyDeviceToHost, stream));
  CK(cudaMemcpyAsync(q, d_q, s_pl * sizeof(T), cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  return 0;
}

template <class T> int Impl<T>::finish() {
  CK(cudaSetDevice(device));
  if (scale_pending) {  // the post-state the caller sees (fv_tracer2d.F90:463-481); k_prep5 read the unscaled arrays
    kbegin();
    CK(fv3t::fast_scale3<T>(cx, cy, mfx, mfy, ksplt_d, n, npz, nt, stream));
    kend(KC_SCALE);
    scale_pending = false;
  }
  // level k has been advanced ksplt(k) times and lives in q[(cur + ksplt(k)) & 1]; gather every level in q[(cur + nsplt) & 1]
  const int fin = (cur + nsplt) & 1;
  if (nsplt != 1) {
    bool any = false;
    for (int k = 0; k < npz; ++k) {
      cpy[k] = (ksplt[k] ^ nsplt) & 1;
      any |= cpy[k] != 0;
    }
    if (any) {
      CK(cudaMemcpyAsync(cpy_d, cpy.data(), npz * sizeof(int), cudaMemcpyHostToDevice, stream));
      CK(cudaStreamSynchronize(stream));  // cpy is reused by the next call
      for (int t = 0; t < nt; ++t) {
        const long total = (long)sz_q(nq_cur);
        kbegin();
        fv3t::k_copy_levels<T><<<1184, 256, 0, stream>>>(q[fin] + (size_t)t * total, q[fin ^ 1] + (size_t)t * total, cpy_d, (long)plane(),
                                                         npz, total);
        kend(KC_SCALE);
      }
    }
  }
  cur = fin;
  CK(cudaGetLastError());
  return 0;
}

template <class T> int Impl<T>::tracer_2d_resident(int nq, int hord, int q_split, T lim_fac, int* nsplt_out) {
  if (nt != 6 || sub_L) return fail("fv3tracer: tracer_2d needs all six whole tiles resident (ntiles = %d); use the *_begin/halo/substep calls", nt);
  std::vector<T> cm(npz);
  int rc = begin(nq, q_split, cm.data());
  if (rc) return rc;
  rc = set_cmax(cm.data(), q_split, nsplt_out);
  if (rc) return rc;
  for (int it = 1; it <= nsplt; ++it) {
    rc = halo_local(it);
    if (rc) return rc;
    rc = substep(it, hord, lim_fac);
    if (rc) return rc;
  }
  return finish();
}

template <class T, int G, bool MAPN> int launch_remap2(Impl<T>& c, const fv3t::Remap2Params<T>& p) {
  const int cols = c.n * p.j_count;
  dim3 grid((cols + 127) / 128, c.nt, (p.nq + G - 1) / G);
  // tuning knob: dynamic shared memory requested only to cap the resident CTAs per SM (the per-thread column scratch lives
  // in local memory, and the resident footprint must stay inside the 126 MB L2)
  static const int cap_smem = getenv("FV3T_REMAP_SMEM") ? atoi(getenv("FV3T_REMAP_SMEM")) : 0;
  if (cap_smem > 48 * 1024) {
    cudaFuncSetAttribute(fv3t::k_remap2<T, G, MAPN, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap_smem);
    cudaFuncSetAttribute(fv3t::k_remap2<T, G, MAPN, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap_smem);
  }
  c.kbegin();
  if (c.npz <= 64)
    fv3t::k_remap2<T, G, MAPN, 64><<<grid, 128, cap_smem, c.stream>>>(p);
  else
    fv3t::k_remap2<T, G, MAPN, 128><<<grid, 128, cap_smem, c.stream>>>(p);
  c.kend(KC_REMAP);
  CK(cudaGetLastError());
  return 0;
}

// Tracer part of Lagrangian_to_Eulerian for rows js+j_first .. (fv_mapz.F90:261-273, 343-368, 407-426): reads q[cur],
// writes q[cur^1]; a whole-tile call (j_count == n) makes the written buffer the current one.
template <class T> int Impl<T>::remap_resident(int nq, const int* kord, int fill, int j_first, int j_count, const T* pe2_ext, const T* dp2_ext) {
  if (nq < 1 || nq > nqmax) return fail("fv3tracer: nq = %d outside 1..nq_max = %d", nq, nqmax);
  if (nq_layout && nq != nq_layout) return fail("fv3tracer: nq = %d, but the resident tracers were uploaded as %d per tile", nq, nq_layout);
  if (!have_vertical && !pe2_ext) return fail("fv3tracer: set_vertical(ak, bk, ptop) has not been called");
  CK(cudaSetDevice(device));
  CK(cudaMemcpyAsync(kord_d, kord, nq * sizeof(int), cudaMemcpyHostToDevice, stream));
  // fv_mapz.F90:410: mapn_tracer for nq > 5, else map1_q2 + fillz per tracer; the row-granular mapn_tracer entry (pe2_ext set)
  // IS mapn_tracer whatever nq
  const bool mapn = nq > 5 || pe2_ext != nullptr;
  bool need_ppm = false;     // map1_q2 sends kord <= 7 to ppm_profile (fv_mapz.F90:1544-1548)
  if (!mapn)
    for (int iq = 0; iq < nq; ++iq) need_ppm |= kord[iq] <= 7;
  int rc = 0;
  if (coef_ready) CK(cudaStreamWaitEvent(stream, ev_coef, 0));  // remap_prepare's side-stream kernel also writes delp
  bool fast_ok = fast && mapn && j_count == n && npz <= 128 && !pe2_ext;
  const int ak0 = kord[0] < 0 ? -kord[0] : kord[0];
  for (int iq = 0; iq < nq; ++iq) {
    const int a = kord[iq] < 0 ? -kord[iq] : kord[iq];
    fast_ok = fast_ok && fv3t::fast_kord_ok(a) && (a == ak0 || (a <= 8 && ak0 <= 8) || (a >= 17 && ak0 >= 17));
  }
  if (fast_ok && remap4_ok() && ak0 == 9) {
    if (!coef4) CK(cudaMalloc((void**)&coef4, fv3t::remap4_coef_bytes<T>(n, nt)));
    if (!neg4) CK(cudaMalloc((void**)&neg4, (size_t)nt * nqmax * n * n));
    fv3t::Remap4Params<T> p{q[cur], q[cur ^ 1], pe, ak, bk, delp, coef4, neg4, ptop, n, npz, nq, nt, fill, 0, nq, 0, 0};
    if (!coef_ready) {  // otherwise computed ahead by remap_prepare on the side stream (waited for above)
      kbegin();
      CK(fv3t::fast_remap_coef4<T>(p, stream));
      kend(KC_SCALE);
    }
    kbegin();
    CK(fv3t::fast_remap4<T>(p, ak0, stream));
    kend(KC_REMAP);
  } else if (fast_ok) {
    int rca = remap_alloc();
    if (rca) return rca;
    fv3t::Remap3Params<T> p{q[cur], q[cur ^ 1], pe, ak, bk, delp, P1, GAM, RD1, R2, ptop, n, npz, nq, nt, fill};
    if (!coef_ready) {  // otherwise computed ahead by remap_prepare on the side stream (waited for above)
      kbegin();
      CK(fv3t::fast_remap_coef3<T>(p, stream));
      kend(KC_SCALE);
    }
    kbegin();
    CK(fv3t::fast_remap3<T>(p, ak0, stream));
    kend(KC_REMAP);
  } else if (need_ppm) {
    fv3t::RemapParams<T> p;
    p.q0 = q[cur];
    p.q1 = q[cur];
    p.qout = q[cur ^ 1];
    p.par = par_d;  // all zero
    p.pe = pe;
    p.ak = ak;
    p.bk = bk;
    p.delp = delp;
    p.kord = kord_d;
    p.ptop = ptop;
    p.n = n;
    p.km = npz;
    p.nq = nq;
    p.ntiles = nt;
    p.fill = fill;
    p.j_first = j_first;
    p.j_count = j_count;
    const int cols = n * j_count;
    dim3 grid((cols + 63) / 64, nt);
    kbegin();
    if (npz <= 64)
      fv3t::k_remap<T, 64><<<grid, 64, 0, stream>>>(p);
    else
      fv3t::k_remap<T, 128><<<grid, 64, 0, stream>>>(p);
    kend(KC_REMAP);
    CK(cudaGetLastError());
  } else {
    fv3t::Remap2Params<T> p;
    p.qsrc = q[cur];
    p.qdst = q[cur ^ 1];
    p.pe = pe;
    p.ak = ak;
    p.bk = bk;
    p.delp = delp;
    p.kord = kord_d;
    p.ptop = ptop;
    p.n = n;
    p.km = npz;
    p.nq = nq;
    p.ntiles = nt;
    p.fill = fill;
    p.j_first = j_first;
    p.j_count = j_count;
    p.pe2_ext = pe2_ext;
    p.dp2_ext = dp2_ext;
    static const int gsel = getenv("FV3T_REMAP_G") ? atoi(getenv("FV3T_REMAP_G")) : 1;  // tuning knob (tracers per thread)
    if (gsel == 1)
      rc = mapn ? launch_remap2<T, 1, true>(*this, p) : launch_remap2<T, 1, false>(*this, p);
    else if (gsel == 2)
      rc = mapn ? launch_remap2<T, 2, true>(*this, p) : launch_remap2<T, 2, false>(*this, p);
    else
      rc = mapn ? launch_remap2<T, 3, true>(*this, p) : launch_remap2<T, 3, false>(*this, p);
  }
  coef_ready = coef_wanted = false;
  if (rc) return rc;
  if (j_count == n) cur ^= 1;
  nq_cur = nq;
  return 0;
}

template <class T> int Impl<T>::remap_alloc() {
  auto dalloc = [&](void** p, size_t bytes) -> cudaError_t { return *p ? cudaSuccess : cudaMalloc(p, bytes); };
  const size_t e1 = plane() * (npz + 1) * nt;
  CK(dalloc((void**)&P1, e1 * sizeof(fv3t::Pair<T>)));
  CK(dalloc((void**)&GAM, e1 * sizeof(T)));
  CK(dalloc((void**)&RD1, sz_c() * nt * sizeof(T)));
  CK(dalloc((void**)&R2, 2 * sz_c() * nt * sizeof(T)));  // Pair{1/dp2, pe2(k+1)} for k_remap5, 1/dp2 for k_remap3
  return 0;
}

// tracer_2d followed by the tracer remap for HOST arrays, as one call (the two are consecutive in fv_dynamics.F90:686-760).
// Tracers never interact on this path, so after the tracer-independent fields are on the device the tracers are processed
// one at a time in a three-stage pipeline -- upload of tracer iq+1 (copy stream 1) | halo + advection + remap of tracer iq
// (context stream) | download of tracer iq-1 (copy stream 2) -- which keeps both PCIe directions busy at once and hides the
// kernels behind the copies.  Post-state as fv3t_*_tracer_2d + fv3t_*_remap_tracers: q, delp always; dp1, cx, cy, mfx, mfy
// only change when nsplt /= 1, and that (rare) case runs the unpipelined path.  Host arrays should be page-locked.
template <class T>
int Impl<T>::tracer_step(T* hq, T* hdp1, T* hmfx, T* hmfy, T* hcx, T* hcy, const T* hpe, const T* hak, const T* hbk, T hptop,
                         T* hdelp, int nq, int hord, int q_split, T lim_fac, const int* kord, int fill, int* nsplt_out) {
  if (sub_L) return fail("fv3tracer: tracer_step is for whole-tile contexts");
  if (nt != 6) return fail("fv3tracer: tracer_step needs all six tiles resident (ntiles = %d)", nt);
  if (nq < 1 || nq > nqmax) return fail("fv3tracer: nq = %d outside 1..nq_max = %d", nq, nqmax);
  CK(cudaSetDevice(device));
  int rc;
  CK(cudaMemcpyAsync(ak, hak, (npz + 1) * sizeof(T), cudaMemcpyHostToDevice, stream));
  CK(cudaMemcpyAsync(bk, hbk, (npz + 1) * sizeof(T), cudaMemcpyHostToDevice, stream));
  ptop = hptop;
  have_vertical = true;
  coef_ready = coef_wanted = false;
  if ((rc = upload(FV3T_DP1, hdp1, nq))) return rc;
  if ((rc = upload(FV3T_MFX, hmfx, nq))) return rc;
  if ((rc = upload(FV3T_MFY, hmfy, nq))) return rc;
  if ((rc = upload(FV3T_CX, hcx, nq))) return rc;
  if ((rc = upload(FV3T_CY, hcy, nq))) return rc;
  if ((rc = upload(FV3T_PE, hpe, nq))) return rc;
  std::vector<T> cm(npz);
  if ((rc = begin(nq, q_split, cm.data()))) return rc;
  if ((rc = set_cmax(cm.data(), q_split, nsplt_out))) return rc;
  bool fast_remap = fast && nq > 5 && npz <= 128;
  const int ak0 = kord[0] < 0 ? -kord[0] : kord[0];
  for (int iq = 0; iq < nq; ++iq) {
    const int a = kord[iq] < 0 ? -kord[iq] : kord[iq];
    fast_remap = fast_remap && fv3t::fast_kord_ok(a) && (a == ak0 || (a <= 8 && ak0 <= 8) || (a >= 17 && ak0 >= 17));
  }
  if (nsplt != 1 || !fast || !fv3t::fast_hord_ok(hord) || !fast_remap || prof) {
    // unpipelined: the sub-steps advance dp1 in place for all tracers at once, the strict kernels work on all tracers
    if ((rc = upload(FV3T_Q, hq, nq))) return rc;
    for (int it = 1; it <= nsplt; ++it) {
      if ((rc = halo_local(it))) return rc;
      if ((rc = substep(it, hord, lim_fac))) return rc;
    }
    if ((rc = finish())) return rc;
    if ((rc = remap_resident(nq, kord, fill, 0, n))) return rc;
    if ((rc = download(FV3T_Q, hq, nq))) return rc;
    if ((rc = download(FV3T_DELP, hdelp, nq))) return rc;
    if ((rc = download(FV3T_DP1, hdp1, nq))) return rc;
    if (nsplt != 1) {
      if ((rc = download(FV3T_MFX, hmfx, nq))) return rc;
      if ((rc = download(FV3T_MFY, hmfy, nq))) return rc;
      if ((rc = download(FV3T_CX, hcx, nq))) return rc;
      if ((rc = download(FV3T_CY, hcy, nq))) return rc;
    }
    return 0;
  }
  // ---- pipelined path (nsplt == 1, fast kernels) ------------------------------------------------------------------------
  if (!s_h2d) {
    CK(cudaStreamCreateWithFlags(&s_h2d, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s_d2h, cudaStreamNonBlocking));
  }
  while ((int)ev_up.size() < nq) {
    cudaEvent_t a, b;
    CK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    ev_up.push_back(a);
    ev_done.push_back(b);
  }
  nq_cur = nq_layout = nq;
  if ((rc = prepare(hord))) return rc;  // k_prep5 / k_prep3 (nsplt == 1: nothing is scaled in place)
  if ((rc = remap_alloc())) return rc;
  const int c0 = cur;                   // per tracer: advect q[c0] -> q[c0^1], remap q[c0^1] -> q[c0]
  fv3t::Remap3Params<T> rp{q[c0 ^ 1], q[c0], pe, ak, bk, delp, P1, GAM, RD1, R2, ptop, n, npz, nq, nt, fill};
  CK(fv3t::fast_remap_coef3<T>(rp, stream));
  ++launches;
  const size_t chunk = plane() * npz;   // one (tile, tracer) block of q
  const int forced = getenv("FV3T_ADV_NT") ? atoi(getenv("FV3T_ADV_NT")) : 0;
  const int NT = (forced >= 32 && forced <= 256 && forced % 32 == 0) ? forced : 64;
  for (int iq = 0; iq < nq; ++iq) {
    for (int t = 0; t < nt; ++t) {
      const size_t off = ((size_t)t * nq + iq) * chunk;
      CK(cudaMemcpyAsync(q[c0] + off, hq + off, chunk * sizeof(T), cudaMemcpyHostToDevice, s_h2d));
    }
    CK(cudaEventRecord(ev_up[iq], s_h2d));
    CK(cudaStreamWaitEvent(stream, ev_up[iq], 0));
    if (halo_len) {
      dim3 grid((halo_len + 255) / 256, npz);
      fv3t::k_halo_fill<T><<<grid, 256, 0, stream>>>(q[c0], halo_dst, halo_src, halo_len, n, npz, nq, ksplt_d, 1, iq * npz);
      ++launches;
    }
    if (call5) {
      fv3t::Adv5Params<T> p{};
      p.qin = q[c0];
      p.qout = q[c0 ^ 1];
      p.X2 = X5;
      p.Y2 = Y5;
      p.CAB = C5;
      p.RX = RX5;
      p.RY = RY5;
      p.MFX = MX5;
      p.MFY = MY5;
      p.AREA = AREA5;
      p.RAREA = RAREA5;
      p.dxa = dxa;
      p.dya = dya;
      p.ksplt = ksplt_d;
      p.n = n;
      p.npz = npz;
      p.nq = nq;
      p.ntiles = nt;
      p.it = 1;
      p.lev0 = 0;
      p.iq0 = iq;
      p.nql = 1;
      p.lim_fac = lim_fac;
      CK(fv3t::fast_advect5<T>(p, maps5, hord, nt * npz, stream));
    } else {
      fv3t::Adv3Params<T> p;
      p.qin = q[c0];
      p.qout = q[c0 ^ 1];
      p.X2 = X2;
      p.Y2 = Y2;
      p.rrx = rrx;
      p.rry = rry;
      p.cab = cab;
      p.mfx = mfx;
      p.mfy = mfy;
      p.area = area;
      p.dxa = dxa;
      p.dya = dya;
      p.ksplt = ksplt_d;
      p.n = n;
      p.npz = npz;
      p.nq = nq;
      p.ntiles = nt;
      p.it = 1;
      p.W = 0;
      p.lim_fac = lim_fac;
      p.iq0 = iq;
      p.nql = 1;
      CK(fv3t::fast_advect3<T>(p, hord, NT, stream));
    }
    rp.iq0 = iq;
    rp.nql = 1;
    CK(fv3t::fast_remap3<T>(rp, ak0, stream));
    launches += 2;
    CK(cudaEventRecord(ev_done[iq], stream));
    CK(cudaStreamWaitEvent(s_d2h, ev_done[iq], 0));
    for (int t = 0; t < nt; ++t) {
      const size_t off = ((size_t)t * nq + iq) * chunk;
      CK(cudaMemcpyAsync(hq + off, q[c0] + off, chunk * sizeof(T), cudaMemcpyDeviceToHost, s_d2h));
    }
    if (iq == 0) CK(cudaMemcpyAsync(hdelp, delp, sz_c() * nt * sizeof(T), cudaMemcpyDeviceToHost, s_d2h));  // written by k_remap_coef3
  }
  CK(cudaStreamSynchronize(s_d2h));
  CK(cudaStreamSynchronize(stream));
  CK(cudaGetLastError());
  return 0;  // cur is unchanged: every tracer went q[c0] -> q[c0^1] -> q[c0]
}

// The Lagrangian pe is final when dyn_core returns, i.e. BEFORE tracer_2d is called (fv_dynamics.F90:600-704): the
// tracer-independent remap coefficients (and delp <- dp2) can therefore be computed on a side stream while the advection
// runs.  Valid until pe, ak, bk or ptop change; the next remap_tracers_resident consumes the result.
template <class T> int Impl<T>::remap_prepare() {
  if (!have_vertical) return fail("fv3tracer: set_vertical(ak, bk, ptop) has not been called");
  if (!fast || npz > 128 || use4) return 0;  // strict kernels compute everything themselves; the experimental k_remap4 has its own coefficients
  coef_wanted = true;                // launched by the next tracer_2d sub-step, right behind its bandwidth-bound preparation
  return 0;
}

// k_remap_coef3 on the side stream, ordered after everything enqueued so far on the context's stream.  Called between the
// (HBM-bound) preparation kernels of tracer_2d and its (issue-bound) advection kernel, so that the HBM-bound coefficient
// kernel overlaps the advection instead of competing with k_cmax / k_prep3 for bandwidth.
template <class T> int Impl<T>::launch_coef_side() {
  coef_wanted = false;
  if (!side) {
    CK(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ev_coef, cudaEventDisableTiming));
  }
  CK(cudaEventRecord(ev_fork, stream));
  CK(cudaStreamWaitEvent(side, ev_fork, 0));
  {
    int rc = remap_alloc();
    if (rc) return rc;
    fv3t::Remap3Params<T> p{q[cur], q[cur ^ 1], pe, ak, bk, delp, P1, GAM, RD1, R2, ptop, n, npz, 0, nt, 1};
    CK(fv3t::fast_remap_coef3<T>(p, side));
  }
  ++launches;
  CK(cudaEventRecord(ev_coef, side));
  coef_ready = true;
  return 0;
}

}  // namespace

// ---- the opaque handle ------------------------------------------------------------------------------
struct fv3t_ctx {
  int prec;  // 4 or 8
  Impl<float>* f32;
  Impl<double>* f64;
};

extern "C" const char* fv3t_last_error(void) { return g_err.c_str(); }
extern "C" int fv3t_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

#define IMPL(ctx, P) ((ctx) && (ctx)->P ? (ctx)->P : nullptr)
#define NEED(ctx, P)                                                                                   \
  auto* I = IMPL(ctx, P);                                                                              \
  if (!I) return fail("fv3tracer: null context or precision mismatch (context is %s)", (ctx) ? ((ctx)->prec == 8 ? "f64" : "f32") : "null")

#define FV3T_DEFINE(P, REAL)                                                                                                   \
  extern "C" int fv3t_##P##_create(fv3t_ctx** ctx, const fv3t_dims* dims, const fv3t_##P##_grid* grid, int device,             \
                                   void* stream) {                                                                             \
    if (!ctx || !dims || !grid) return fail("fv3tracer: null argument");                                                       \
    auto* I = new Impl<REAL>();                                                                                                \
    const REAL* g[7] = {grid->area, grid->rarea, grid->dx, grid->dy, grid->dxa, grid->dya, grid->sin_sg};                      \
    const int rc = I->create(dims, g, device, stream);                                                                         \
    if (rc) {                                                                                                                  \
      I->destroy();                                                                                                            \
      delete I;                                                                                                                \
      return rc;                                                                                                               \
    }                                                                                                                          \
    auto* c = new fv3t_ctx{(int)sizeof(REAL), nullptr, nullptr};                                                               \
    c->P = I;                                                                                                                  \
    *ctx = c;                                                                                                                  \
    return 0;                                                                                                                  \
  }                                                                                                                            \
  extern "C" int fv3t_##P##_upload(fv3t_ctx* ctx, int field, const REAL* host, int nq) {                                       \
    NEED(ctx, P);                                                                                                              \
    return I->upload(field, host, nq);                                                                                         \
  }                                                                                                                            \
  extern "C" int fv3t_##P##_download(fv3t_ctx* ctx, int field, REAL* host, int nq) {                                           \
    NEED(ctx, P);                                                                                                              \
    return I->download(field, host, nq);                                                                                       \
  }                                                                                                                            \
  extern "C" int fv3t_##P##_set_vertical(fv3t_ctx* ctx, const REAL* ak, const REAL* bk, REAL ptop) {                           \
    NEED(ctx, P);                                                                                                              \
    CK(cudaSetDevice(I->device));                                                                                              \
    CK(cudaMemcpyAsync(I->ak, ak, (I->npz + 1) * sizeof(REAL), cudaMemcpyHostToDevice, I->stream));                            \
    CK(cudaMemcpyAsync(I->bk, bk, (I->npz + 1) * sizeof(REAL), cudaMemcpyHostToDevice, I->stream));                            \
    CK(cudaStreamSynchronize(I->stream));                                                                                      \
    I->ptop = ptop;                                                                                                            \
    I->coef_ready = I->coef_wanted = false;                                                                                    \
    I->have_vertical = true;                                                                                                   \
    return 0;                                                                                                                  \
  }                                                                                                                            \
  extern "C" int fv3t_##P##_tracer_2d_resident(fv3t_ctx* ctx, int nq, int hord, int q_split, REAL lim_fac, int* nsplt_out) {   \
    NEED(ctx, P);                                                                                                              \
    return I->tracer_2d_resident(nq, hord, q_split, lim_fac, nsplt_out);                                                       \
  }                                                                                                                            \
  extern "C" int fv3t_##P##_remap_tracers_resident(fv3t_ctx* ctx, int nq, const int* kord_tr, int fill) {                      \
    NEED(ctx, P);                                                                                                              \
    return I->remap_resident(nq, kord_tr, fill, 0, I->n);                                                                      \
  }                                                                                                                            \
  extern "C" int fv3t_##P##_tracer_step(fv3t_ctx* ctx, REAL* q, REAL* dp1, REAL* mfx, REAL* mfy, REAL* cx, REAL* cy,              \
                                        const REAL* pe, const REAL* ak, const REAL* bk, REAL ptop, REAL* delp, int nq, int hord,  \
                                        int q_split, REAL lim_fac, const int* kord_tr, int fill, int* nsplt_out) {               \
    NEED(ctx, P);                                                                                                              \
    if (!q || !dp1 || !mfx || !mfy || !cx || !cy || !pe || !ak || !bk || !delp || !kord_tr) return fail("fv3tracer: null argument"); \
    return I->tracer_step(q, dp1, mfx, mfy, cx, cy, pe, ak, bk, ptop, delp, nq, hord, q_split, lim_fac, kord_tr, fill, nsplt_out); \
  }                                                                                                                            \
  extern "C" int fv3t_##P##_remap_prepare(fv3t_ctx* ctx) {                                                                    \
    NEED(ctx, P);                                                                                                              \
    return I->remap_prepare();                                                                                                 \
  }                                                                                                                            \
  extern "C" int fv3t_##P##_tracer_2d(fv3t_ctx* ctx, REAL* q, REAL* dp1, REAL* mfx, REAL* mfy, REAL* cx, REAL* cy, int nq,     \
                                      int hord, int q_split, int nord_tr, REAL trdm, REAL lim_fac, int* nsplt_out,            \
                                      int* ksplt_out) {                                                                        \
    NEED(ctx, P);                                                                                                              \
    if (trdm > REAL(1.e-4) && !I->del6_u)                                                                                      \
      return fail("fv3tracer: tracer damping (trdm2 > 1e-4, deln_flux) needs the damping metrics: call fv3t_*_set_damping first"); \
    const REAL trdm_keep = I->damp_trdm;                                                                                       \
    const int nord_keep = I->damp_nord;                                                                                        \
    I->damp_trdm = trdm;                                                                                                       \
    I->damp_nord = nord_tr;                                                                                                    \
    struct Restore {                                                                                                           \
      Impl<REAL>* i; REAL t; int n;                                                                                            \
      ~Restore() { i->damp_trdm = t; i->damp_nord = n; }                                                                       \
    } restore{I, trdm_keep, nord_keep};                                                                                        \
    int rc;                                                                                                                    \
    if ((rc = I->upload(FV3T_Q, q, nq))) return rc;                                                                            \
    if ((rc = I->upload(FV3T_DP1, dp1, nq))) return rc;                                                                        \
    if ((rc = I->upload(FV3T_MFX, mfx, nq))) return rc;                                                                        \
    if ((rc = I->upload(FV3T_MFY, mfy, nq))) return rc;                                                                        \
    if ((rc = I->upload(FV3T_CX, cx, nq))) return rc;                                                                          \
    if ((rc = I->upload(FV3T_CY, cy, nq))) return rc;                                                                          \
    if ((rc = I->tracer_2d_resident(nq, hord, q_split, lim_fac, nsplt_out))) return rc;                                        \
    if (ksplt_out) std::memcpy(ksplt_out, I->ksplt.data(), sizeof(int) * I->npz);                                              \
    if ((rc = I->download(FV3T_Q, q, nq))) return rc;                                                                          \
    if ((rc = I->download(FV3T_DP1, dp1, nq))) return rc;                                                                      \
    if (I->nsplt != 1) {                                                                                                       \
      if ((rc = I->download(FV3T_MFX, mfx, nq))) return rc;                                                                    \
      if ((rc = I->download(FV3T_MFY, mfy, nq))) return rc;                                                                    \
      if ((rc = I->download(FV3T_CX, cx, nq))) return rc;                                                                      \
      if ((rc = I->download(FV3T_CY, cy, nq))) return rc;                                                                      \
    }                                                                                                                          \
    return 0;                                                                                                                  \
  }                                                                                                                            \
  extern "C" int fv3t_##P##_set_damping(fv3t_ctx* ctx, const REAL* del6_u, const REAL* del6_v, REAL da_min, int nord_tr,      \
                                        REAL trdm) {                                                                           \
    NEED(ctx, P);                                                                                                              \
    CK(cudaSetDevice(I->device));                                                                                              \
    if (del6_u && del6_v) {                                                                                                    \
      const size_t ne = (size_t)I->nt * (I->n + 6) * (I->n + 7);                                                               \
      if (!I->del6_u) CK(cudaMalloc((void**)&I->del6_u, ne * sizeof(REAL)));                                                   \
      if (!I->del6_v) CK(cudaMalloc((void**)&I->del6_v, ne * sizeof(REAL)));                                                   \
      CK(cudaMemcpyAsync(I->del6_u, del6_u, ne * sizeof(REAL), cudaMemcpyHostToDevice, I->stream));                            \
      CK(cudaMemcpyAsync(I->del6_v, del6_v, ne * sizeof(REAL), cudaMemcpyHostToDevice, I->stream));                            \
      CK(cudaStreamSynchronize(I->stream));                                                                                    \
      I->da_min = da_min;                                                                                                      \
    } else if (trdm > REAL(1.e-4) && !I->del6_u) {                                                                             \
      return fail("fv3tracer: set_damping: del6_u / del6_v have never been provided");                                        \
    }                                                                                                                          \
    if (nord_tr < 0 || nord_tr > 2) return fail("fv3tracer: nord_tr = %d outside 0..2", nord_tr);                             \
    I->damp_nord = nord_tr;                                                                                                    \
    I->damp_trdm = trdm;                                                                                                       \
    return 0;                                                                                                                  \
  }                                                                                                                            \
  extern "C" int fv3t_##P##_tracer_2d_1L(fv3t_ctx* ctx, REAL* q, REAL* dp1, REAL* mfx, REAL* mfy, REAL* cx, REAL* cy, int nq,  \
                                         int hord, int q_split, int nord_tr, REAL trdm, REAL lim_fac, int* nsplt_out,         \
                                         int* ksplt_out) {                                                                     \
    NEED(ctx, P);                                                                                                              \
    (void)q_split; /* not used by tracer_2d_1L (fv_tracer2d.F90:92-321) */                                                     \
    I->mode_1l = true;                                                                                                         \
    const int rc = fv3t_##P##_tracer_2d(ctx, q, dp1, mfx, mfy, cx, cy, nq, hord, 0, nord_tr, trdm, lim_fac, nsplt_out, ksplt_out); \
    I->mode_1l = false;                                                                                                        \
    return rc;                                                                                                                 \
  }                                                                                                                            \
  extern "C" int fv3t_##P##_remap_tracers(fv3t_ctx* ctx, const REAL* pe, const REAL* ak, const REAL* bk, REAL ptop, REAL* q,   \
                                          REAL* delp, int nq, const int* kord_tr, int fill) {                                  \
    NEED(ctx, P);                                                                                                              \
    int rc;                                                                                                                    \
    if ((rc = fv3t_##P##_set_vertical(ctx, ak, bk, ptop))) return rc;                                                          \
    if ((rc = I->upload(FV3T_PE, pe, nq))) return rc;                                                                          \
    if ((rc = I->upload(FV3T_Q, q, nq))) return rc;                                                                            \
    if ((rc = I->upload(FV3T_DELP, delp, nq))) return rc;                                                                      \
    if ((rc = I->remap_resident(nq, kord_tr, fill, 0, I->n))) return rc;                                                       \
    if ((rc = I->download(FV3T_Q, q, nq))) return rc;                                                                          \
    return I->download(FV3T_DELP, delp, nq);                                                                                   \
  }                                                                                                                            \
  extern "C" int fv3t_##P##_map_field(fv3t_ctx* ctx, REAL* q, const REAL* qs, int iv, int kord, REAL q_min, int use_cs) {        \
    NEED(ctx, P);                                                                                                              \
    if (!q) return fail("fv3tracer: null argument");                                                                           \
    if (!I->have_vertical) return fail("fv3tracer: set_vertical(ak, bk, ptop) has not been called");                          \
    if (iv == -2 && !qs) return fail("fv3tracer: iv = -2 needs the bottom boundary values qs");                               \
    if (iv < -2 || iv > 2) return fail("fv3tracer: iv = %d is not a mode of scalar_profile / cs_profile", iv);                \
    CK(cudaSetDevice(I->device));                                                                                              \
    const size_t ne = I->sz_c() * I->nt, n2 = I->plane() * I->nt;                                                              \
    if (!I->fld) CK(cudaMalloc((void**)&I->fld, ne * sizeof(REAL)));                                                           \
    CK(cudaMemcpyAsync(I->fld, q, ne * sizeof(REAL), cudaMemcpyHostToDevice, I->stream));                                      \
    if (qs) {                                                                                                                  \
      if (!I->fld_qs) CK(cudaMalloc((void**)&I->fld_qs, n2 * sizeof(REAL)));                                                   \
      CK(cudaMemcpyAsync(I->fld_qs, qs, n2 * sizeof(REAL), cudaMemcpyHostToDevice, I->stream));                                \
    }                                                                                                                          \
    fv3t::MapFieldParams<REAL> p{I->fld, qs ? I->fld_qs : nullptr, I->pe, I->ak, I->bk, I->ptop, q_min, I->n, I->npz, I->nt, iv, kord, \
                                 use_cs ? 1 : 0};                                                                              \
    dim3 grid((I->n * I->n + 63) / 64, I->nt);                                                                                 \
    I->kbegin();                                                                                                               \
    if (I->npz <= 64)                                                                                                          \
      fv3t::k_map_field<REAL, 64><<<grid, 64, 0, I->stream>>>(p);                                                              \
    else                                                                                                                       \
      fv3t::k_map_field<REAL, 128><<<grid, 64, 0, I->stream>>>(p);                                                             \
    I->kend(KC_REMAP);                                                                                                         \
    CK(cudaGetLastError());                                                                                                    \
    CK(cudaMemcpyAsync(q, I->fld, ne * sizeof(REAL), cudaMemcpyDeviceToHost, I->stream));                                      \
    CK(cudaStreamSynchronize(I->stream));                                                                                      \
    return 0;                                                                                                                  \
  }                                                                                                                            \
  extern "C" int fv3t_##P##_map_scalar(fv3t_ctx* ctx, REAL* q, const REAL* qs, int iv, int kord, REAL q_min) {                 \
    return fv3t_##P##_map_field(ctx, q, qs, iv, kord, q_min, 0);                                                               \
  }                                                                                                                            \
  extern "C" int fv3t_##P##_map1_ppm(fv3t_ctx* ctx, REAL* q, const REAL* qs, int iv, int kord) {                               \
    return fv3t_##P##_map_field(ctx, q, qs, iv, kord, REAL(0), 1);                                                             \
  }                                                                                                                            \
  extern "C" int fv3t_##P##_mapn_tracer(fv3t_ctx* ctx, int nq, int km, const REAL* pe1, const REAL* pe2, REAL* q1,             \
                                        const REAL* dp2, const int* kord, int j, int i1, int i2, int isd, int ied, int jsd,   \
                                        int jed, REAL q_min, int fill) {                                                       \
    NEED(ctx, P);                                                                                                              \
    if (!pe1 || !pe2 || !q1 || !dp2 || !kord) return fail("fv3tracer: null argument");                                         \
    if (km != I->npz || i1 != 1 || i2 != I->n || isd != -2 || ied != I->n + 3 || jsd != -2 || jed != I->n + 3)                 \
      return fail("fv3tracer: mapn_tracer bounds do not match the context (whole-tile rows only)");                           \
    if (q_min != REAL(0)) return fail("fv3tracer: mapn_tracer is only called with q_min = 0 on this path");                   \
    if (I->nt != 1) return fail("fv3tracer: the row-granular mapn_tracer entry needs a one-tile context");                    \
    if (nq < 1 || nq > I->nqmax) return fail("fv3tracer: nq = %d outside 1..nq_max = %d", nq, I->nqmax);                       \
    std::lock_guard<std::mutex> lk(I->row_mutex);                                                                              \
    CK(cudaSetDevice(I->device));                                                                                              \
    const size_t nd = I->n + 6, pl = nd * nd;                                                                                  \
    /* pe row: pe(i1:i2, k) for this j into the (is-1:ie+1, km+1, js-1:je+1) mirror */                                         \
    CK(cudaMemcpy2DAsync(I->pe + (size_t)j * (I->n + 2) * (km + 1) + 1, (I->n + 2) * sizeof(REAL), pe1, I->n * sizeof(REAL),   \
                         I->n * sizeof(REAL), km + 1, cudaMemcpyHostToDevice, I->stream));                                     \
    /* q rows: (isd:ied) of row j for every (k, iq) */                                                                         \
    CK(cudaMemcpy2DAsync(I->q[I->cur] + (size_t)(j + 2) * nd, pl * sizeof(REAL), q1 + (size_t)(j + 2) * nd, pl * sizeof(REAL), \
                         nd * sizeof(REAL), (size_t)km * nq, cudaMemcpyHostToDevice, I->stream));                              \
    /* the caller's target grid: pe2 (i1:i2, km+1) and dp2 (i1:i2, km) of this row, consumed as given */                      \
    const size_t npe2 = (size_t)I->n * (km + 1), ndp2 = (size_t)I->n * km;                                                     \
    if (!I->row_buf) CK(cudaMalloc((void**)&I->row_buf, (npe2 + ndp2) * sizeof(REAL)));                                        \
    CK(cudaMemcpyAsync(I->row_buf, pe2, npe2 * sizeof(REAL), cudaMemcpyHostToDevice, I->stream));                              \
    CK(cudaMemcpyAsync(I->row_buf + npe2, dp2, ndp2 * sizeof(REAL), cudaMemcpyHostToDevice, I->stream));                       \
    I->nq_layout = nq; /* the rows were just laid out for nq tracers */                                                        \
    int rc = I->remap_resident(nq, kord, fill, j - 1, 1, I->row_buf, I->row_buf + npe2);                                       \
    if (rc) return rc;                                                                                                         \
    CK(cudaMemcpy2DAsync(q1 + (size_t)(j + 2) * nd, pl * sizeof(REAL), I->q[I->cur ^ 1] + (size_t)(j + 2) * nd,                \
                         pl * sizeof(REAL),                                                                                    \
                         nd * sizeof(REAL), (size_t)km * nq, cudaMemcpyDeviceToHost, I->stream));                              \
    CK(cudaStreamSynchronize(I->stream));                                                                                      \
    return 0;                                                                                                                  \
  }                                                                                                                            \
  extern "C" int fv3t_##P##_fv_tp_2d(fv3t_ctx* ctx, int nlev, REAL* q, const REAL* crx, const REAL* cry, int hord, REAL* fx,   \
                                     REAL* fy, const REAL* xfx, const REAL* yfx, const REAL* ra_x, const REAL* ra_y,           \
                                     REAL lim_fac, const REAL* mfx, const REAL* mfy, const REAL* mass, int nord, REAL damp_c) { \
    NEED(ctx, P);                                                                                                              \
    return I->fv_tp_2d_host(nlev, q, crx, cry, hord, fx, fy, xfx, yfx, ra_x, ra_y, lim_fac, mfx, mfy, mass, nord, damp_c);      \
  }                                                                                                                            \
  extern "C" int fv3t_##P##_tracer_2d_begin(fv3t_ctx* ctx, int nq, int q_split, REAL* cmax_local) {                            \
    NEED(ctx, P);                                                                                                              \
    return I->begin(nq, q_split, cmax_local);                                                                                  \
  }                                                                                                                            \
  extern "C" int fv3t_##P##_tracer_2d_set_cmax(fv3t_ctx* ctx, const REAL* cmax_global, int q_split, int* nsplt_out) {          \
    NEED(ctx, P);                                                                                                              \
    return I->set_cmax(cmax_global, q_split, nsplt_out);                                                                       \
  }                                                                                                                            \
  extern "C" int fv3t_##P##_halo_local(fv3t_ctx* ctx, int it) {                                                                \
    NEED(ctx, P);                                                                                                              \
    return I->halo_local(it);                                                                                                  \
  }                                                                                                                            \
  extern "C" int fv3t_##P##_halo_pack(fv3t_ctx* ctx, int it, int local_tile, int edge, REAL* dev_buf) {                        \
    NEED(ctx, P);                                                                                                              \
    return I->halo_pack(it, local_tile, edge, dev_buf, false);                                                                 \
  }                                                                                                                            \
  extern "C" int fv3t_##P##_halo_unpack(fv3t_ctx* ctx, int it, int local_tile, int edge, const REAL* dev_buf) {                \
    NEED(ctx, P);                                                                                                              \
    return I->halo_pack(it, local_tile, edge, const_cast<REAL*>(dev_buf), true);                                               \
  }                                                                                                                            \
  extern "C" int fv3t_##P##_halo_pack_host(fv3t_ctx* ctx, int it, int local_tile, int edge, REAL* host_buf) {                  \
    NEED(ctx, P);                                                                                                              \
    const size_t ne = (size_t)3 * I->n * I->npz * I->nq_cur;                                                                   \
    if (!host_buf || ne == 0) return fail("fv3tracer: halo_pack_host: null buffer or no resident tracers");                   \
    CK(cudaSetDevice(I->device));                                                                                              \
    if (I->strip_cap < ne) {                                                                                                   \
      if (I->strip_buf) cudaFree(I->strip_buf);                                                                                \
      I->strip_buf = nullptr;                                                                                                  \
      CK(cudaMalloc((void**)&I->strip_buf, ne * sizeof(REAL)));                                                                \
      I->strip_cap = ne;                                                                                                       \
    }                                                                                                                          \
    const int rc = I->halo_pack(it, local_tile, edge, I->strip_buf, false);                                                    \
    if (rc) return rc;                                                                                                         \
    CK(cudaMemcpyAsync(host_buf, I->strip_buf, ne * sizeof(REAL), cudaMemcpyDeviceToHost, I->stream));                         \
    CK(cudaStreamSynchronize(I->stream));                                                                                      \
    return 0;                                                                                                                  \
  }                                                                                                                            \
  extern "C" int fv3t_##P##_halo_unpack_host(fv3t_ctx* ctx, int it, int local_tile, int edge, const REAL* host_buf) {          \
    NEED(ctx, P);                                                                                                              \
    const size_t ne = (size_t)3 * I->n * I->npz * I->nq_cur;                                                                   \
    if (!host_buf || ne == 0) return fail("fv3tracer: halo_unpack_host: null buffer or no resident tracers");                 \
    CK(cudaSetDevice(I->device));                                                                                              \
    if (I->strip_cap < ne) {                                                                                                   \
      if (I->strip_buf) cudaFree(I->strip_buf);                                                                                \
      I->strip_buf = nullptr;                                                                                                  \
      CK(cudaMalloc((void**)&I->strip_buf, ne * sizeof(REAL)));                                                                \
      I->strip_cap = ne;                                                                                                       \
    }                                                                                                                          \
    CK(cudaMemcpyAsync(I->strip_buf, host_buf, ne * sizeof(REAL), cudaMemcpyHostToDevice, I->stream));                         \
    const int rc = I->halo_pack(it, local_tile, edge, I->strip_buf, true);                                                     \
    if (rc) return rc;                                                                                                         \
    CK(cudaStreamSynchronize(I->stream));                                                                                      \
    return 0;                                                                                                                  \
  }                                                                                                                            \
  extern "C" int fv3t_##P##_halo_gather(fv3t_ctx* ctx, int it, int local_tile, int list, REAL* dev_buf, int buf_stride) {     \
    NEED(ctx, P);                                                                                                              \
    return I->halo_list_move(it, local_tile, list, dev_buf, false, buf_stride);                                                \
  }                                                                                                                            \
  extern "C" int fv3t_##P##_halo_scatter(fv3t_ctx* ctx, int it, int local_tile, int list, const REAL* dev_buf,                 \
                                         int buf_stride) {                                                                     \
    NEED(ctx, P);                                                                                                              \
    return I->halo_list_move(it, local_tile, list, const_cast<REAL*>(dev_buf), true, buf_stride);                              \
  }                                                                                                                            \
  extern "C" int fv3t_##P##_tracer_2d_substep(fv3t_ctx* ctx, int it, int hord, REAL lim_fac) {                                 \
    NEED(ctx, P);                                                                                                              \
    return I->substep(it, hord, lim_fac);                                                                                      \
  }                                                                                                                            \
  extern "C" int fv3t_##P##_tracer_2d_finish(fv3t_ctx* ctx) {                                                                  \
    NEED(ctx, P);                                                                                                              \
    return I->finish();                                                                                                        \
  }

FV3T_DEFINE(f64, double)
FV3T_DEFINE(f32, float)

#define DISPATCH(ctx, expr64, expr32) ((ctx)->prec == 8 ? (expr64) : (expr32))

extern "C" int fv3t_destroy(fv3t_ctx* ctx) {
  if (!ctx) return 0;
  if (ctx->f64) {
    ctx->f64->destroy();
    delete ctx->f64;
  }
  if (ctx->f32) {
    ctx->f32->destroy();
    delete ctx->f32;
  }
  delete ctx;
  return 0;
}
extern "C" int fv3t_sync(fv3t_ctx* ctx) {
  if (!ctx) return fail("fv3tracer: null context");
  cudaStream_t s = DISPATCH(ctx, ctx->f64->stream, ctx->f32->stream);
  CK(cudaStreamSynchronize(s));
  return 0;
}
extern "C" void* fv3t_device_ptr(fv3t_ctx* ctx, int field) {
  if (!ctx) return nullptr;
  return DISPATCH(ctx, (void*)ctx->f64->field_ptr(field), (void*)ctx->f32->field_ptr(field));
}
extern "C" size_t fv3t_halo_strip_elems(fv3t_ctx* ctx) {
  if (!ctx) return 0;
  return DISPATCH(ctx, (size_t)3 * ctx->f64->n * ctx->f64->npz * ctx->f64->nq_cur, (size_t)3 * ctx->f32->n * ctx->f32->npz * ctx->f32->nq_cur);
}
extern "C" int fv3t_halo_list_create(fv3t_ctx* ctx, const int* offsets, int count, int* list) {
  if (!ctx) return fail("fv3tracer: null context");
  return DISPATCH(ctx, ctx->f64->halo_list_create(offsets, count, list), ctx->f32->halo_list_create(offsets, count, list));
}
extern "C" int fv3t_halo_list_count(fv3t_ctx* ctx, int list) {
  if (!ctx) return -1;
  const std::vector<int>& ll = DISPATCH(ctx, ctx->f64->list_len, ctx->f32->list_len);
  return (list >= 0 && list < (int)ll.size()) ? ll[list] : -1;
}
extern "C" int fv3t_halo_local_table(fv3t_ctx* ctx, const int* dst, const int* src, int len) {
  if (!ctx) return fail("fv3tracer: null context");
  return DISPATCH(ctx, ctx->f64->halo_local_table(dst, src, len), ctx->f32->halo_local_table(dst, src, len));
}
extern "C" int fv3t_neighbor(fv3t_ctx* ctx, int global_tile, int edge, int* nbr_tile, int* nbr_edge, int* rotated) {
  if (!ctx || global_tile < 1 || global_tile > 6 || edge < 0 || edge > 3) return fail("fv3tracer: bad tile/edge");
  const EdgeMap& m = DISPATCH(ctx, ctx->f64->maps[global_tile - 1][edge], ctx->f32->maps[global_tile - 1][edge]);
  if (nbr_tile) *nbr_tile = m.nbr_tile + 1;
  if (nbr_edge) *nbr_edge = m.nbr_edge;
  if (rotated) *rotated = (m.A[0][0] == 0);
  return 0;
}
extern "C" uint64_t fv3t_kernel_launches(fv3t_ctx* ctx) {
  if (!ctx) return 0;
  return DISPATCH(ctx, ctx->f64->launches, ctx->f32->launches);
}
extern "C" int fv3t_timer_start(fv3t_ctx* ctx) {
  if (!ctx) return fail("fv3tracer: null context");
  if (ctx->prec == 8)
    CK(cudaEventRecord(ctx->f64->ev0, ctx->f64->stream));
  else
    CK(cudaEventRecord(ctx->f32->ev0, ctx->f32->stream));
  return 0;
}
extern "C" int fv3t_timer_stop_ms(fv3t_ctx* ctx, float* ms) {
  if (!ctx || !ms) return fail("fv3tracer: null argument");
  cudaEvent_t e0 = DISPATCH(ctx, ctx->f64->ev0, ctx->f32->ev0), e1 = DISPATCH(ctx, ctx->f64->ev1, ctx->f32->ev1);
  cudaStream_t s = DISPATCH(ctx, ctx->f64->stream, ctx->f32->stream);
  CK(cudaEventRecord(e1, s));
  CK(cudaEventSynchronize(e1));
  CK(cudaEventElapsedTime(ms, e0, e1));
  return 0;
}
extern "C" int fv3t_profile_enable(fv3t_ctx* ctx, int on) {
  if (!ctx) return fail("fv3tracer: null context");
  if (ctx->prec == 8) {
    ctx->f64->prof = on != 0;
    for (int k = 0; k < KC_N; ++k) ctx->f64->prof_ms[k] = 0, ctx->f64->prof_n[k] = 0;
  } else {
    ctx->f32->prof = on != 0;
    for (int k = 0; k < KC_N; ++k) ctx->f32->prof_ms[k] = 0, ctx->f32->prof_n[k] = 0;
  }
  return 0;
}
extern "C" int fv3t_profile_get_ms(fv3t_ctx* ctx, int kc, float* total_ms, int* launches) {
  if (!ctx || kc < 0 || kc >= KC_N) return fail("fv3tracer: bad argument");
  if (total_ms) *total_ms = DISPATCH(ctx, ctx->f64->prof_ms[kc], ctx->f32->prof_ms[kc]);
  if (launches) *launches = DISPATCH(ctx, ctx->f64->prof_n[kc], ctx->f32->prof_n[kc]);
  return 0;
}
