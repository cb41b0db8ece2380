// fv3atm_b200: horizontal tracer advection kernels (tracer_2d -> fv_tp_2d) for sm_100a.
//
// Replaces, per sub-step, the k / iq loops of tracer_2d (atmos_cubed_sphere/model/fv_tracer2d.F90:503-556)
// and everything fv_tp_2d calls (model/tp_core.F90:110-249): one CTA owns a TX x TY block of cells of one
// (tile, level), stages the tracer-independent fields of that block once (Courant numbers, area fluxes,
// areas, mass fluxes, dp1/dp2) and then loops over all tracers, running the four PPM sweeps, the flux
// average and the flux-divergence update entirely in shared memory / registers.  xfx, yfx, ra_x, ra_y,
// dp2, fx, fy and the q_i / q_j intermediates of the reference never touch HBM.
//
// Data layout in HBM = the Fortran layout of the host arrays, tile-major:
//   q   (isd:ied, jsd:jed, npz, nq)  two buffers (ping-pong: a sub-step reads one and writes the other,
//                                    because neighbouring CTAs read each other's cells through the halo)
//   dp1 (isd:ied, jsd:jed, npz)   cx (is:ie+1, jsd:jed, npz)   cy (isd:ied, js:je+1, npz)
//   mfx (is:ie+1, js:je, npz)     mfy (is:ie, js:je+1, npz)
// cx, cy, mfx, mfy stay UNSCALED in memory during the sub-steps; the per-level factor frac = 1/ksplt(k)
// (fv_tracer2d.F90:457-481) is applied on the fly in the reference's operation order
// (cx*frac ; ((cx*dxa)*dy*sin_sg)*frac ; mfx*frac), and k_scale_frac writes the scaled post-state once.
#pragma once
#include "fv3t_ppm.cuh"

namespace fv3t {

template <class T> struct GridDev {
  const T *area, *rarea, *dx, *dy, *dxa, *dya, *sin_sg;  // tile-major, Fortran extents
};

template <class T> struct AdvParams {
  const T* qin;
  T* qout;
  T* dp1;
  const T *mfx, *mfy, *cx, *cy;
  GridDev<T> g;
  const int* ksplt;  // [npz] device
  int n, npz, nq, ntiles;
  int it, nsplt;
  T lim_fac;
};

// ---------------------------------------------------------------------------------------------------
// cmax(k) per tile (fv_tracer2d.F90:409-424).  One CTA per (tile, level); max is order independent.
// ---------------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(512) k_cmax(const T* __restrict__ cx, const T* __restrict__ cy, const T* __restrict__ sin_sg,
                                             T* __restrict__ cmax_t, int n, int npz) {
  const int t = blockIdx.x / npz, kz = blockIdx.x % npz;
  const long nd = n + 6;
  const T* cxp = cx + ((long)t * npz + kz) * (long)(n + 1) * nd;
  const T* cyp = cy + ((long)t * npz + kz) * nd * (long)(n + 1);
  const T* s5 = sin_sg + (long)t * nd * nd * 5 + 4 * nd * nd;
  const bool upper = (kz + 1) < npz / 6;
  T cm = T(0);
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    const int j = idx / n + 1, i = idx % n + 1;
    const T ax = f_abs(cxp[(long)(j + 2) * (n + 1) + (i - 1)]);
    const T ay = f_abs(cyp[(long)(j - 1) * nd + (i + 2)]);
    if (upper)
      cm = f_max(cm, ax, ay);
    else
      cm = f_max(cm, f_max(ax, ay) + T(1) - s5[(long)(j + 2) * nd + (i + 2)]);
  }
  __shared__ T red[512];
  red[threadIdx.x] = cm;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] = f_max(red[threadIdx.x], red[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) cmax_t[blockIdx.x] = red[0];
}

// In-place frac scaling of cx, cy, mfx, mfy (fv_tracer2d.F90:449-486): the post-state the caller sees.
template <class T>
__global__ void k_scale_frac(T* __restrict__ a, const int* __restrict__ ksplt, long plane, int npz, long total) {
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int kz = (int)((e / plane) % npz);
    const T frac = T(1) / (T)ksplt[kz];
    a[e] = a[e] * frac;
  }
}

// Scalar edge-halo fill of q for all tracers and the levels still active at sub-step `it`
// (complete_group_halo_update, fv_tracer2d.F90:499).  dst/src: flat offsets into the tile-major stack of
// (n+6)^2 planes, built from the 12-contact mosaic (fv_mp_mod.F90:581-629).
template <class T>
__global__ void k_halo_fill(T* __restrict__ q, const int* __restrict__ dst, const int* __restrict__ src, int len, int n, int npz,
                            int nq, const int* __restrict__ ksplt, int it) {
  const long plane = (long)(n + 6) * (n + 6);
  const long tile_stride = plane * npz * nq;
  const int pl = blockIdx.y;  // iq*npz + kz
  const int kz = pl % npz;
  if (it > ksplt[kz]) return;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < len; e += gridDim.x * blockDim.x) {
    const int d = dst[e], s = src[e];
    const long dt = d / plane, dr = d % plane, st = s / plane, sr = s % plane;
    q[dt * tile_stride + (long)pl * plane + dr] = q[st * tile_stride + (long)pl * plane + sr];
  }
}

// Move levels that an earlier call left in buffer 1 back to buffer 0 (only when tracer_2d is called twice
// without a remap in between).
template <class T>
__global__ void k_copy_levels(T* __restrict__ q0, const T* __restrict__ q1, const int* __restrict__ par, long plane, int npz,
                              long total) {
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int kz = (int)((e / plane) % npz);
    if (par[kz]) q0[e] = q1[e];
  }
}

// ---------------------------------------------------------------------------------------------------
// The fused sub-step kernel.
// Shared-memory frame: every array is an (TY+6) x P slab addressed [lj*P + li]; local cell (li, lj) is global
// cell (i0-3+li, j0-3+lj).  x-faces / y-faces carry the index of the cell to their east / north.
// ---------------------------------------------------------------------------------------------------
template <int TX, int TY> struct AdvTile {
  static constexpr int P = TX + 7;  // odd pitch (TX even): rows of 8-byte words fall on distinct bank pairs
  static constexpr int H = TY + 6;
  static constexpr int SZ = P * H;
  static constexpr int NARR = 13;
  static constexpr int NTHREADS = 256;
  static constexpr int CPT = (TX * TY + NTHREADS - 1) / NTHREADS;  // cells per thread in the update
  template <class T> static constexpr size_t smem_bytes() { return (size_t)NARR * SZ * sizeof(T) + SZ; }
};

// One PPM sweep over a shared-memory slab.  DIR = 0: along li (x), DIR = 1: along lj (y).
//   src     : field to reconstruct (q, q_i or q_j)
//   cour    : Courant numbers at the faces of this direction (already scaled by frac)
//   flux    : output, upwind flux at faces u = 3 .. 3+nu  for cross index v in [v0, v1)
//   g0      : global index of local u = 0;  nu = number of compute cells of the block along the sweep
//   met     : global edge metric accessor met(global_u, local_v)  (dxa(:,j) or dya(i,:))
template <class T, int ORD, int DIR, int TX, int TY, class MF>
__device__ __forceinline__ void ppm_sweep(const T* __restrict__ src, const T* __restrict__ cour, T* __restrict__ flux,
                                          T* __restrict__ sa, T* __restrict__ sbl, T* __restrict__ sbr,
                                          unsigned char* __restrict__ sfl, int g0, int nu, int v0, int v1, int npx, T lim_fac,
                                          MF met) {
  using TL = AdvTile<TX, TY>;
  constexpr int P = TL::P;
  const int tid = threadIdx.x;
  const int nv = v1 - v0;
  // address of (u, v)
  auto at = [&](int u, int v) -> int { return DIR == 0 ? v * P + u : u * P + v; };
  // iterate a (u-range) x (v-range) box with the contiguous smem coordinate (li) fastest
  auto for_box = [&](int u0, int u1, auto&& body) {
    const int nuu = u1 - u0;
    const int tot = nuu * nv;
    for (int idx = tid; idx < tot; idx += TL::NTHREADS) {
      int u, v;
      if (DIR == 0) {
        v = v0 + idx / nuu;
        u = u0 + idx % nuu;
      } else {
        u = u0 + idx / nv;
        v = v0 + idx % nv;
      }
      body(u, v);
    }
  };
  // phase pre: ORD >= 7: dm on cells u = 1 .. nu+4 ; ORD < 7: al on faces u = 2 .. nu+4
  for_box(ORD >= 7 ? 1 : 2, nu + 5, [&](int u, int v) {
    auto q = [&](int gi) -> T { return src[at(gi - g0, v)]; };
    auto dxa = [&](int gi) -> T { return met(gi, v); };
    sa[at(u, v)] = ppm_pre<T, ORD>(g0 + u, npx, q, dxa);
  });
  __syncthreads();
  // phase blbr: cells u = 2 .. nu+3 (global g0+2 = first compute cell - 1)
  if (!(ORD < 7 && (ORD == 2 || ORD == -2))) {
    for_box(2, nu + 4, [&](int u, int v) {
      auto q = [&](int gi) -> T { return src[at(gi - g0, v)]; };
      auto a = [&](int gi) -> T { return sa[at(gi - g0, v)]; };
      auto dxa = [&](int gi) -> T { return met(gi, v); };
      T bl, br;
      int flg;
      ppm_blbr<T, ORD>(g0 + u, npx, q, a, dxa, lim_fac, bl, br, flg);
      sbl[at(u, v)] = bl;
      sbr[at(u, v)] = br;
      if (ORD < 7) sfl[at(u, v)] = (unsigned char)flg;
    });
    __syncthreads();
  }
  // phase flux: faces u = 3 .. nu+3
  for_box(3, nu + 4, [&](int u, int v) {
    auto q = [&](int gi) -> T { return src[at(gi - g0, v)]; };
    auto bl = [&](int gi) -> T { return sbl[at(gi - g0, v)]; };
    auto br = [&](int gi) -> T { return sbr[at(gi - g0, v)]; };
    auto fl = [&](int gi) -> int { return (int)sfl[at(gi - g0, v)]; };
    auto al = [&](int gi) -> T { return sa[at(gi - g0, v)]; };
    flux[at(u, v)] = ppm_flux<T, ORD>(g0 + u, cour[at(u, v)], q, bl, br, fl, al);
  });
  __syncthreads();
}

template <class T, int ORD_IN, int ORD_OU, int TX, int TY>
__global__ void __launch_bounds__(256) k_advect(const AdvParams<T> p) {
  using TL = AdvTile<TX, TY>;
  constexpr int P = TL::P, SZ = TL::SZ, NT = TL::NTHREADS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  T* qs = sm;            // q tile (later overwritten in place by q_j)
  T* sa = sm + 1 * SZ;   // dm / al
  T* sbl = sm + 2 * SZ;
  T* sbr = sm + 3 * SZ;
  T* fx2 = sm + 4 * SZ;  // inner x flux (ord_in on q)
  T* fy2 = sm + 5 * SZ;  // inner y flux
  T* qi = sm + 6 * SZ;   // q_i, later the outer y flux
  T* fxo = sm + 7 * SZ;  // outer x flux
  T* cxs = sm + 8 * SZ;  // cx*frac at x-faces
  T* cys = sm + 9 * SZ;
  T* xfs = sm + 10 * SZ;  // xfx at x-faces
  T* yfs = sm + 11 * SZ;
  T* ars = sm + 12 * SZ;  // area
  unsigned char* sfl = smem_raw + (size_t)TL::NARR * SZ * sizeof(T);
  T* fyo = qi;

  const int n = p.n, npz = p.npz, npx = n + 1;
  const int t = blockIdx.z / npz, kz = blockIdx.z % npz;
  const int ks = p.ksplt[kz];
  if (p.it > ks) return;
  const T frac = T(1) / (T)ks;
  const int i0 = 1 + blockIdx.x * TX, j0 = 1 + blockIdx.y * TY;
  const int ni = min(TX, n - i0 + 1), nj = min(TY, n - j0 + 1);
  const int tid = threadIdx.x;
  const long nd = n + 6;
  const long plane = nd * nd;
  // global (i,j) -> offsets
  auto o_c = [&](int i, int j) -> long { return (long)(j + 2) * nd + (i + 2); };         // (isd:ied, jsd:jed)
  auto o_cx = [&](int i, int j) -> long { return (long)(j + 2) * (n + 1) + (i - 1); };  // (is:ie+1, jsd:jed)
  auto o_cy = [&](int i, int j) -> long { return (long)(j - 1) * nd + (i + 2); };       // (isd:ied, js:je+1)
  auto o_dy = [&](int i, int j) -> long { return (long)(j + 2) * (nd + 1) + (i + 2); }; // (isd:ied+1, jsd:jed)
  auto o_mfx = [&](int i, int j) -> long { return (long)(j - 1) * (n + 1) + (i - 1); };
  auto o_mfy = [&](int i, int j) -> long { return (long)(j - 1) * n + (i - 1); };
  const T* area = p.g.area + (long)t * plane;
  const T* rarea = p.g.rarea + (long)t * plane;
  const T* dxg = p.g.dx + (long)t * nd * (nd + 1);
  const T* dyg = p.g.dy + (long)t * (nd + 1) * nd;
  const T* dxa = p.g.dxa + (long)t * plane;
  const T* dya = p.g.dya + (long)t * plane;
  const T* ssg = p.g.sin_sg + (long)t * plane * 5;
  const T* cxp = p.cx + ((long)t * npz + kz) * (long)(n + 1) * nd;
  const T* cyp = p.cy + ((long)t * npz + kz) * nd * (long)(n + 1);
  const T* mfxp = p.mfx + ((long)t * npz + kz) * (long)(n + 1) * n;
  const T* mfyp = p.mfy + ((long)t * npz + kz) * (long)n * (n + 1);
  T* dp1p = p.dp1 + ((long)t * npz + kz) * plane;
  const int gi0 = i0 - 3, gj0 = j0 - 3;  // global index of local 0

  // ---- tracer-independent staging -------------------------------------------------------------
  // x faces li = 3..3+ni (global i0..i0+ni), all rows lj = 0..nj+5
  for (int idx = tid; idx < (ni + 1) * (nj + 6); idx += NT) {
    const int lj = idx / (ni + 1), li = 3 + idx % (ni + 1);
    const int i = gi0 + li, j = gj0 + lj;
    const T c = cxp[o_cx(i, j)];
    T xf;
    if (c > T(0))
      xf = c * dxa[o_c(i - 1, j)] * dyg[o_dy(i, j)] * ssg[2 * plane + o_c(i - 1, j)];
    else
      xf = c * dxa[o_c(i, j)] * dyg[o_dy(i, j)] * ssg[0 * plane + o_c(i, j)];
    cxs[lj * P + li] = c * frac;
    xfs[lj * P + li] = xf * frac;
  }
  // y faces lj = 3..3+nj, all columns li = 0..ni+5
  for (int idx = tid; idx < (ni + 6) * (nj + 1); idx += NT) {
    const int lj = 3 + idx / (ni + 6), li = idx % (ni + 6);
    const int i = gi0 + li, j = gj0 + lj;
    const T c = cyp[o_cy(i, j)];
    T yf;
    if (c > T(0))
      yf = c * dya[o_c(i, j - 1)] * dxg[o_c(i, j)] * ssg[3 * plane + o_c(i, j - 1)];
    else
      yf = c * dya[o_c(i, j)] * dxg[o_c(i, j)] * ssg[1 * plane + o_c(i, j)];
    cys[lj * P + li] = c * frac;
    yfs[lj * P + li] = yf * frac;
  }
  for (int idx = tid; idx < (ni + 6) * (nj + 6); idx += NT) {
    const int lj = idx / (ni + 6), li = idx % (ni + 6);
    ars[lj * P + li] = area[o_c(gi0 + li, gj0 + lj)];
  }
  // per-thread compute cells: mass fluxes, dp1, dp2, rarea
  T r_mfx0[TL::CPT], r_mfx1[TL::CPT], r_mfy0[TL::CPT], r_mfy1[TL::CPT], r_dp1[TL::CPT], r_dp2[TL::CPT], r_ra[TL::CPT];
#pragma unroll
  for (int c = 0; c < TL::CPT; ++c) {
    const int idx = tid + c * NT;
    const int lj = 3 + idx / TX, li = 3 + idx % TX;
    const bool ok = (idx < TX * TY) && (li < 3 + ni) && (lj < 3 + nj);
    if (ok) {
      const int i = gi0 + li, j = gj0 + lj;
      r_mfx0[c] = mfxp[o_mfx(i, j)] * frac;
      r_mfx1[c] = mfxp[o_mfx(i + 1, j)] * frac;
      r_mfy0[c] = mfyp[o_mfy(i, j)] * frac;
      r_mfy1[c] = mfyp[o_mfy(i, j + 1)] * frac;
      r_ra[c] = rarea[o_c(i, j)];
      r_dp1[c] = dp1p[o_c(i, j)];
      r_dp2[c] = r_dp1[c] + (r_mfx0[c] - r_mfx1[c] + r_mfy0[c] - r_mfy1[c]) * r_ra[c];
    }
  }
  __syncthreads();

  const bool at_w = (i0 == 1), at_e = (i0 + ni - 1 == n), at_s = (j0 == 1), at_n = (j0 + nj - 1 == n);
  auto met_x = [&](int gi, int lj) -> T { return dxa[o_c(gi, gj0 + lj)]; };
  auto met_y = [&](int gj, int li) -> T { return dya[o_c(gi0 + li, gj)]; };

  for (int iq = 0; iq < p.nq; ++iq) {
    const T* qg = p.qin + (((long)t * p.nq + iq) * npz + kz) * plane;
    T* qo = p.qout + (((long)t * p.nq + iq) * npz + kz) * plane;
    // ---- load the q tile with its 3-cell halo
    for (int idx = tid; idx < (ni + 6) * (nj + 6); idx += NT) {
      const int lj = idx / (ni + 6), li = idx % (ni + 6);
      qs[lj * P + li] = qg[o_c(gi0 + li, gj0 + lj)];
    }
    __syncthreads();
    // own cells (needed for the final update after qs is overwritten by q_j)
    T r_q[TL::CPT];
#pragma unroll
    for (int c = 0; c < TL::CPT; ++c) {
      const int idx = tid + c * NT;
      const int lj = 3 + idx / TX, li = 3 + idx % TX;
      if ((idx < TX * TY) && (li < 3 + ni) && (lj < 3 + nj)) r_q[c] = qs[lj * P + li];
    }
    // ---- copy_corners(dir = 2) for CTAs that own a tile corner (tp_core.F90:164-166, 298-326)
    if ((at_w || at_e) && (at_s || at_n) && tid < 9) {
      const int a = tid / 3, b = tid % 3;
      const int npy = npx;
      if (at_w && at_s) {
        const int i = -2 + b, j = -2 + a;
        qs[(j - gj0) * P + (i - gi0)] = qg[o_c(1 - j, i)];
      }
      if (at_e && at_s) {
        const int i = npx + b, j = -2 + a;
        qs[(j - gj0) * P + (i - gi0)] = qg[o_c(npy + j - 1, npx - i)];
      }
      if (at_e && at_n) {
        const int i = npx + b, j = npy + a;
        qs[(j - gj0) * P + (i - gi0)] = qg[o_c(2 * npy - 1 - j, i)];
      }
      if (at_w && at_n) {
        const int i = -2 + b, j = npy + a;
        qs[(j - gj0) * P + (i - gi0)] = qg[o_c(j + 1 - npx, npy - i)];
      }
    }
    __syncthreads();
    // ---- inner y sweep on q over all columns li = 0..ni+5  -> fy2 ; q_i  (tp_core.F90:168-180)
    ppm_sweep<T, ORD_IN, 1, TX, TY>(qs, cys, fy2, sa, sbl, sbr, sfl, gj0, nj, 0, ni + 6, npx, p.lim_fac, met_y);
    for (int idx = tid; idx < (ni + 6) * nj; idx += NT) {
      const int lj = 3 + idx / (ni + 6), li = idx % (ni + 6);
      const int o = lj * P + li;
      const T fyy0 = yfs[o] * fy2[o];
      const T fyy1 = yfs[o + P] * fy2[o + P];
      const T ra_y = ars[o] + yfs[o] - yfs[o + P];
      qi[o] = (qs[o] * ars[o] + fyy0 - fyy1) / ra_y;
    }
    __syncthreads();
    // ---- outer x sweep on q_i over rows lj = 3..3+nj-1 -> fxo  (tp_core.F90:182)
    ppm_sweep<T, ORD_OU, 0, TX, TY>(qi, cxs, fxo, sa, sbl, sbr, sfl, gi0, ni, 3, 3 + nj, npx, p.lim_fac, met_x);
    // ---- copy_corners(dir = 1)  (tp_core.F90:185-187, 265-296)
    if ((at_w || at_e) && (at_s || at_n) && tid < 9) {
      const int a = tid / 3, b = tid % 3;
      const int npy = npx;
      if (at_w && at_s) {
        const int i = -2 + b, j = -2 + a;
        qs[(j - gj0) * P + (i - gi0)] = qg[o_c(j, 1 - i)];
      }
      if (at_e && at_s) {
        const int i = npx + b, j = -2 + a;
        qs[(j - gj0) * P + (i - gi0)] = qg[o_c(npy - j, i - npx + 1)];
      }
      if (at_e && at_n) {
        const int i = npx + b, j = npy + a;
        qs[(j - gj0) * P + (i - gi0)] = qg[o_c(j, 2 * npx - 1 - i)];
      }
      if (at_w && at_n) {
        const int i = -2 + b, j = npy + a;
        qs[(j - gj0) * P + (i - gi0)] = qg[o_c(npy - j, i - 1 + npx)];
      }
    }
    __syncthreads();
    // ---- inner x sweep on q over all rows lj = 0..nj+5 -> fx2 ; q_j in place of q  (tp_core.F90:189-199)
    ppm_sweep<T, ORD_IN, 0, TX, TY>(qs, cxs, fx2, sa, sbl, sbr, sfl, gi0, ni, 0, nj + 6, npx, p.lim_fac, met_x);
    for (int idx = tid; idx < ni * (nj + 6); idx += NT) {
      const int lj = idx / ni, li = 3 + idx % ni;
      const int o = lj * P + li;
      const T fx10 = xfs[o] * fx2[o];
      const T fx11 = xfs[o + 1] * fx2[o + 1];
      const T ra_x = ars[o] + xfs[o] - xfs[o + 1];
      qs[o] = (qs[o] * ars[o] + fx10 - fx11) / ra_x;
    }
    __syncthreads();
    // ---- outer y sweep on q_j over columns li = 3..3+ni-1 -> fyo  (tp_core.F90:201)
    ppm_sweep<T, ORD_OU, 1, TX, TY>(qs, cys, fyo, sa, sbl, sbr, sfl, gj0, nj, 3, 3 + ni, npx, p.lim_fac, met_y);
    // ---- flux average with the mass fluxes and flux-divergence update
    //      (tp_core.F90:208-221, fv_tracer2d.F90:539-544)
#pragma unroll
    for (int c = 0; c < TL::CPT; ++c) {
      const int idx = tid + c * NT;
      const int lj = 3 + idx / TX, li = 3 + idx % TX;
      if ((idx < TX * TY) && (li < 3 + ni) && (lj < 3 + nj)) {
        const int o = lj * P + li;
        const T fxa = T(0.5) * (fxo[o] + fx2[o]) * r_mfx0[c];
        const T fxb = T(0.5) * (fxo[o + 1] + fx2[o + 1]) * r_mfx1[c];
        const T fya = T(0.5) * (fyo[o] + fy2[o]) * r_mfy0[c];
        const T fyb = T(0.5) * (fyo[o + P] + fy2[o + P]) * r_mfy1[c];
        qo[o_c(gi0 + li, gj0 + lj)] = (r_q[c] * r_dp1[c] + (fxa - fxb + fya - fyb) * r_ra[c]) / r_dp2[c];
      }
    }
    __syncthreads();
  }
  // dp1 <- dp2 between sub-steps (fv_tracer2d.F90:547-553; note: tests the GLOBAL nsplt)
  if (p.it != p.nsplt) {
#pragma unroll
    for (int c = 0; c < TL::CPT; ++c) {
      const int idx = tid + c * NT;
      const int lj = 3 + idx / TX, li = 3 + idx % TX;
      if ((idx < TX * TY) && (li < 3 + ni) && (lj < 3 + nj)) dp1p[o_c(gi0 + li, gj0 + lj)] = r_dp2[c];
    }
  }
}

}  // namespace fv3t
