// fv3atm_b200: fv_tp_2d as an operator of its own (atmos_cubed_sphere/model/tp_core.F90:110-249) -- the entry the dynamical core
// calls for delp, pt, vorticity and w (the branch WITHOUT mass fluxes, :236-248) as well as for tracers (:209-234).  SURVEY.md
// section 8, row f1.
//
// tracer_2d never needs the fluxes themselves (k_advect5 keeps them in registers), so this entry is not on the bench path: it is a
// sequence of plane kernels, one thread per face / cell, every (tile, level) plane independent, that materialise exactly the
// reference's intermediates in the reference's operation order (this translation unit is built with -fmad=false: bit-identical to
// the FMA-free oracle for every hord):
//     k_tp2d_flux<ORD>  fy2 = yppm(q [dir-2 corner view], cry, ord_in)              i = isd..ied
//     k_tp2d_mid        q_i = (q area + yfx fy2(j) - yfx fy2(j+1)) / ra_y
//     k_tp2d_flux<ORD>  fx  = xppm(q_i, crx, ord_ou)
//     k_tp2d_flux<ORD>  fx2 = xppm(q [dir-1 corner view], crx, ord_in)             j = jsd..jed
//     k_tp2d_mid        q_j = (q area + xfx fx2(i) - xfx fx2(i+1)) / ra_x
//     k_tp2d_flux<ORD>  fy  = yppm(q_j, cry, ord_ou)
//     k_tp2d_avg        fx = 0.5 (fx + fx2) mfx|xfx,  fy = 0.5 (fy + fy2) mfy|yfx
//     [k_tp2d_scale, k_deln_flux, nord x (k_deln_div, k_deln_flux), k_tp2d_damp_add]   deln_flux (:1239-1387), both forms
//     k_tp2d_corners    q's corner blocks <- the dir = 1 view (the post-state of the INTENT(INOUT) dummy, :189)
// copy_corners itself is never executed before the sweeps: they read through corner_src (fv3t_deln.cuh).
#pragma once
#include "fv3t_advect2.cuh"
#include "fv3t_deln.cuh"

namespace fv3t {

template <class T> struct Tp2dFlux {
  const T* src;    // operand planes [plane][nd][nd]
  const T* cour;   // Courant numbers; index of (face g, line l): dir 0: (l - c_l0) * c_pitch + (g - 1), dir 1: (g - 1) * c_pitch + (l - c_l0)
  T* flux;         // same rule with f_pitch, f_l0
  const T* dxa;    // per tile [nd][nd]: dxa (dir 0) or dya (dir 1)
  long c_plane, f_plane;  // elements per plane of cour / flux
  int c_pitch, c_l0, f_pitch, f_l0;
  int n, nlev, dir, view;  // dir 0: x faces along rows (line = j); dir 1: y faces along columns (line = i).  view: 0 | copy_corners dir
  int l_lo, l_hi;
  T lim_fac;
};

template <class T> struct Tp2dMid {
  const T *q, *flux, *xf, *area, *ra;  // flux, xf: (1:n+1, jsd:jed) (dir 0) | (isd:ied, 1:n+1) (dir 1); ra: ra_x (1:n, jsd:jed) | ra_y (isd:ied, 1:n)
  T* out;                              // [plane][nd][nd]: q_j (dir 0) | q_i (dir 1)
  int n, nlev, dir;
};

template <class T> struct Tp2dAvg {
  T *fx, *fy;               // (1:n+1, 1:n), (1:n, 1:n+1): outer fluxes in, averaged fluxes out
  const T *fx2, *fy2;       // (1:n+1, jsd:jed), (isd:ied, 1:n+1)
  const T *mx, *my;         // mfx (1:n+1, 1:n), mfy (1:n, 1:n+1)  |  xfx, yfx (shapes of fx2, fy2)
  int n, tracer;
};

template <class T> struct Tp2dDampAdd {
  T *fx, *fy;
  const T *dfx2, *dfy2;  // deln_flux's fx2 [plane][nd][nd+1], fy2 [plane][nd+1][nd]
  const T* mass;         // [plane][nd][nd] or null
  T damp;
  int n;
};

#ifdef __CUDACC__
template <class T, int ORD> __global__ void __launch_bounds__(128) k_tp2d_flux(const Tp2dFlux<T> p) {
  const int n = p.n, nd = n + 6, npx = n + 1;
  const int pl = blockIdx.y, t = pl / p.nlev;
  const T* s = p.src + (long)pl * nd * nd;
  const T* d = p.dxa + (long)t * nd * nd;
  const T* cr = p.cour + (long)pl * p.c_plane;
  T* fl = p.flux + (long)pl * p.f_plane;
  const int nl = p.l_hi - p.l_lo + 1;
  const int dir = p.dir, view = p.view;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nl * (n + 1); e += gridDim.x * blockDim.x) {
    // consecutive threads walk along i (the contiguous index) in both directions
    const int g = dir == 0 ? e % (n + 1) + 1 : e / nl + 1;
    const int l = p.l_lo + (dir == 0 ? e / (n + 1) : e % nl);
    auto off = [&](int gi) -> int { return dir == 0 ? (l + 2) * nd + (gi + 2) : (gi + 2) * nd + (l + 2); };
    auto q = [&](int gi) -> T { return s[view ? (dir == 0 ? corner_src(view, gi, l, n) : corner_src(view, l, gi, n)) : off(gi)]; };
    auto dx = [&](int gi) -> T { return d[off(gi)]; };
    auto a = [&](int gi) -> T { return ppm_pre<T, ORD>(gi, npx, q, dx); };
    const long ci = dir == 0 ? (long)(l - p.c_l0) * p.c_pitch + (g - 1) : (long)(g - 1) * p.c_pitch + (l - p.c_l0);
    const long fi = dir == 0 ? (long)(l - p.f_l0) * p.f_pitch + (g - 1) : (long)(g - 1) * p.f_pitch + (l - p.f_l0);
    fl[fi] = xface_flux<T, ORD>(g, cr[ci], npx, p.lim_fac, q, a, dx);
  }
}

template <class T> __global__ void __launch_bounds__(256) k_tp2d_mid(const Tp2dMid<T> p) {
  const int n = p.n, nd = n + 6;
  const int pl = blockIdx.y, t = pl / p.nlev;
  const T* q = p.q + (long)pl * nd * nd;
  const T* ar = p.area + (long)t * nd * nd;
  T* out = p.out + (long)pl * nd * nd;
  if (p.dir == 1) {  // q_i on (isd:ied, 1:n)
    const T* f = p.flux + (long)pl * nd * (n + 1);
    const T* xf = p.xf + (long)pl * nd * (n + 1);
    const T* ra = p.ra + (long)pl * nd * n;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nd * n; e += gridDim.x * blockDim.x) {
      const int j = e / nd + 1, c = e % nd;  // c = i + 2
      const long o = (long)(j + 2) * nd + c, of = (long)(j - 1) * nd + c;
      const T f1 = xf[of] * f[of], f2 = xf[of + nd] * f[of + nd];
      out[o] = (q[o] * ar[o] + f1 - f2) / ra[of];
    }
  } else {  // q_j on (1:n, jsd:jed)
    const T* f = p.flux + (long)pl * (n + 1) * nd;
    const T* xf = p.xf + (long)pl * (n + 1) * nd;
    const T* ra = p.ra + (long)pl * n * nd;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nd * n; e += gridDim.x * blockDim.x) {
      const int r = e / n, i = e % n + 1;  // r = j + 2
      const long o = (long)r * nd + (i + 2), of = (long)r * (n + 1) + (i - 1);
      const T f1 = xf[of] * f[of], f2 = xf[of + 1] * f[of + 1];
      out[o] = (q[o] * ar[o] + f1 - f2) / ra[(long)r * n + (i - 1)];
    }
  }
}

template <class T> __global__ void __launch_bounds__(256) k_tp2d_avg(const Tp2dAvg<T> p) {
  const int n = p.n, nd = n + 6;
  const int pl = blockIdx.y;
  T* fx = p.fx + (long)pl * (n + 1) * n;
  T* fy = p.fy + (long)pl * n * (n + 1);
  const T* fx2 = p.fx2 + (long)pl * (n + 1) * nd;
  const T* fy2 = p.fy2 + (long)pl * nd * (n + 1);
  const T* mx = p.mx + (long)pl * (n + 1) * (p.tracer ? n : nd);
  const T* my = p.my + (long)pl * (p.tracer ? n : nd) * (n + 1);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n * (n + 1); e += gridDim.x * blockDim.x) {
    {  // fx(i, j), i = 1..n+1, j = 1..n
      const int j = e / (n + 1) + 1, i = e % (n + 1) + 1;
      const long o = (long)(j - 1) * (n + 1) + (i - 1), o2 = (long)(j + 2) * (n + 1) + (i - 1);
      fx[o] = T(0.5) * (fx[o] + fx2[o2]) * mx[p.tracer ? o : o2];
    }
    {  // fy(i, j), i = 1..n, j = 1..n+1
      const int j = e / n + 1, i = e % n + 1;
      const long o = (long)(j - 1) * n + (i - 1), o2 = (long)(j - 1) * nd + (i + 2);
      fy[o] = T(0.5) * (fy[o] + fy2[o2]) * my[p.tracer ? o : o2];
    }
  }
}

// d2 = damp * q, every cell of the plane (deln_flux without `mass`, tp_core.F90:1275-1281)
template <class T> __global__ void __launch_bounds__(256) k_tp2d_scale(T* d2, const T* q, T damp, long count) {
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (long)gridDim.x * blockDim.x) d2[e] = damp * q[e];
}

template <class T> __global__ void __launch_bounds__(256) k_tp2d_damp_add(const Tp2dDampAdd<T> p) {
  const int n = p.n, nd = n + 6;
  const int pl = blockIdx.y;
  T* fx = p.fx + (long)pl * (n + 1) * n;
  T* fy = p.fy + (long)pl * n * (n + 1);
  const T* dfx = p.dfx2 + (long)pl * nd * (nd + 1);
  const T* dfy = p.dfy2 + (long)pl * (nd + 1) * nd;
  const T* m = p.mass ? p.mass + (long)pl * nd * nd : nullptr;
  const T damp2 = T(0.5) * p.damp;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n * (n + 1); e += gridDim.x * blockDim.x) {
    {
      const int j = e / (n + 1) + 1, i = e % (n + 1) + 1;
      const long o = (long)(j - 1) * (n + 1) + (i - 1), od = (long)(j + 2) * (nd + 1) + (i + 2), oc = (long)(j + 2) * nd + (i + 2);
      fx[o] = m ? fx[o] + damp2 * (m[oc - 1] + m[oc]) * dfx[od] : fx[o] + dfx[od];
    }
    {
      const int j = e / n + 1, i = e % n + 1;
      const long o = (long)(j - 1) * n + (i - 1), oc = (long)(j + 2) * nd + (i + 2);
      fy[o] = m ? fy[o] + damp2 * (m[oc - nd] + m[oc]) * dfy[oc] : fy[o] + dfy[oc];
    }
  }
}

// the four 3 x 3 corner blocks of q <- the dir = 1 view (sources are edge-halo cells, never corner cells: no race)
template <class T> __global__ void __launch_bounds__(64) k_tp2d_corners(T* q, int n) {
  const int nd = n + 6;
  T* s = q + (long)blockIdx.x * nd * nd;
  const int e = threadIdx.x;
  if (e >= 36) return;
  const int c = e / 9, a = e % 9 / 3, b = e % 3;
  const int i = (c & 1) ? n + 1 + a : a - 2, j = (c & 2) ? n + 1 + b : b - 2;
  s[(j + 2) * nd + (i + 2)] = s[corner_src(1, i, j, n)];
}
#endif

}  // namespace fv3t
