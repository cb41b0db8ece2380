// fv3atm_b200: tracer damping -- deln_flux (atmos_cubed_sphere/model/tp_core.F90:1239-1387, the form fv_tp_2d calls with `mass`
// present, :229-234) as used by tracer_2d on its FIRST sub-step when trdm2 > 1e-4 (model/fv_tracer2d.F90:487-494, 527-532).
//
// The reference adds the del-(2 nord + 2) diffusive fluxes to fx, fy before the flux-form update.  The marching advection kernels
// never materialise fx, fy, and the update is linear in them, so the damping is applied as a correction of the advected field:
//     q_new += (dfx(i,j) - dfx(i+1,j) + dfy(i,j) - dfy(i,j+1)) * rarea / dp2,   dfx = 0.5 damp (mass(i-1,j) + mass(i,j)) fx2(i,j)
// (same terms, summed in a different order: agrees with the reference to rounding, not bit for bit).  Off by default in every
// BASELINE configuration, so this is a plain sequence of bandwidth-bound plane kernels, one tracer at a time:
//     k_deln_flux0   fx2, fy2 from q (the sub-step's input field with its edge halos filled)
//     nord x { k_deln_div: d2 = div(fx2, fy2) rarea ;  k_deln_flux: fx2, fy2 from d2 }
//     k_deln_apply   the correction above
// copy_corners (tp_core.F90:253-330) is never executed: fx2 reads its operand through the dir = 1 corner view and fy2 through the
// dir = 2 view (corner_src), exactly the cells copy_corners would have copied.
#pragma once
#include "fv3t_advect3.cuh"

namespace fv3t {

// plane offset (row-major (j+2)*nd + (i+2)) of the cell that copy_corners(dir) puts at (i, j); identity outside the corner blocks
FV3T_HD int corner_src(int dir, int i, int j, int n) {
  const int nd = n + 6, npx = n + 1;
  if ((i >= 1 && i <= n) || (j >= 1 && j <= n)) return (j + 2) * nd + (i + 2);
  int si, sj;
  if (dir == 1) {
    if (i < 1 && j < 1) {
      si = j, sj = 1 - i;                    // SW
    } else if (i > n && j < 1) {
      si = npx - j, sj = i - npx + 1;        // SE
    } else if (i > n) {
      si = j, sj = 2 * npx - 1 - i;          // NE
    } else {
      si = npx - j, sj = i - 1 + npx;        // NW
    }
  } else {
    if (i < 1 && j < 1) {
      si = 1 - j, sj = i;
    } else if (i > n && j < 1) {
      si = npx + j - 1, sj = npx - i;
    } else if (i > n) {
      si = 2 * npx - 1 - j, sj = i;
    } else {
      si = j + 1 - npx, sj = npx - i;
    }
  }
  return (sj + 2) * nd + (si + 2);
}

template <class T> struct DelnParams {
  const T* src;      // operand planes: q of one tracer (stage 0) or d2, [tile*npz][nd][nd]
  T *fx2, *fy2;      // [tile*npz][nd][nd+1] / [tile*npz][nd+1][nd]
  T* d2;             // [tile*npz][nd][nd]
  const T *del6_u, *del6_v, *rarea;  // per tile
  int n, npz, nord, nt;  // nt = nord - stage
  int first;             // stage 0: differences taken as (d2(i-1) - d2(i)); later stages (d2(i) - d2(i-1))
  long src_tile_stride;  // elements between the tiles of `src` (q: plane*npz*nq; d2: plane*npz)
};

#ifdef __CUDACC__
// fx2 on (is-nt : ie+nt+1, js-nt : je+nt), fy2 on (is-nt : ie+nt, js-nt : je+nt+1)
template <class T> __global__ void __launch_bounds__(256) k_deln_flux(const DelnParams<T> p) {
  const int n = p.n, nd = n + 6;
  const long plane = (long)nd * nd;
  const int lev = blockIdx.y, t = lev / p.npz, kz = lev % p.npz;
  const T* s = p.src + (long)t * p.src_tile_stride + (long)kz * plane;
  T* fx2 = p.fx2 + (long)lev * nd * (nd + 1);
  T* fy2 = p.fy2 + (long)lev * (nd + 1) * nd;
  const T* d6v = p.del6_v + (long)t * nd * (nd + 1);
  const T* d6u = p.del6_u + (long)t * (nd + 1) * nd;
  const int lo = 1 - p.nt, hi = n + p.nt;
  const int w = hi - lo + 2;  // the larger of the two extents
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < w * w; e += gridDim.x * blockDim.x) {
    const int j = lo + e / w, i = lo + e % w;
    if (j <= hi) {  // fx2(i, j), i = lo .. hi+1
      const T a = s[corner_src(1, i - 1, j, n)], b = s[corner_src(1, i, j, n)];
      fx2[(long)(j + 2) * (nd + 1) + (i + 2)] = d6v[(long)(j + 2) * (nd + 1) + (i + 2)] * (p.first ? a - b : b - a);
    }
    if (i <= hi) {  // fy2(i, j), j = lo .. hi+1
      const T a = s[corner_src(2, i, j - 1, n)], b = s[corner_src(2, i, j, n)];
      fy2[(long)(j + 2) * nd + (i + 2)] = d6u[(long)(j + 2) * nd + (i + 2)] * (p.first ? a - b : b - a);
    }
  }
}
// d2 on (is-nt-1 : ie+nt+1)^2 (its corner blocks are never read: see corner_src)
template <class T> __global__ void __launch_bounds__(256) k_deln_div(const DelnParams<T> p) {
  const int n = p.n, nd = n + 6;
  const long plane = (long)nd * nd;
  const int lev = blockIdx.y, t = lev / p.npz;
  const T* fx2 = p.fx2 + (long)lev * nd * (nd + 1);
  const T* fy2 = p.fy2 + (long)lev * (nd + 1) * nd;
  T* d2 = p.d2 + (long)lev * plane;
  const T* ra = p.rarea + (long)t * plane;
  const int lo = 1 - p.nt - 1, hi = n + p.nt + 1, w = hi - lo + 1;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < w * w; e += gridDim.x * blockDim.x) {
    const int j = lo + e / w, i = lo + e % w;
    if ((i < 1 || i > n) && (j < 1 || j > n)) continue;
    const long ox = (long)(j + 2) * (nd + 1) + (i + 2), oy = (long)(j + 2) * nd + (i + 2);
    d2[oy] = (fx2[ox] - fx2[ox + 1] + fy2[oy] - fy2[oy + nd]) * ra[oy];
  }
}

template <class T> struct DelnApplyParams {
  T* qout;           // the advected field of this tracer, [tile][npz] planes with tile stride qstride
  const T *fx2, *fy2, *dp1, *mfx, *mfy, *rarea;
  const int* ksplt;
  long qstride;
  int n, npz, scaled;  // scaled: mfx, mfy already carry the 1/ksplt factor
  T damp;
};
template <class T> __global__ void __launch_bounds__(256) k_deln_apply(const DelnApplyParams<T> p) {
  const int n = p.n, nd = n + 6;
  const long plane = (long)nd * nd;
  const int lev = blockIdx.y, t = lev / p.npz, kz = lev % p.npz;
  const T* fx2 = p.fx2 + (long)lev * nd * (nd + 1);
  const T* fy2 = p.fy2 + (long)lev * (nd + 1) * nd;
  const T* m = p.dp1 + (long)lev * plane;
  const T* mx = p.mfx + (long)lev * (n + 1) * n;
  const T* my = p.mfy + (long)lev * n * (n + 1);
  const T* ra = p.rarea + (long)t * plane;
  T* q = p.qout + (long)t * p.qstride + (long)kz * plane;
  const T frac = p.scaled ? T(1) : T(1) / (T)p.ksplt[kz];
  const T damp2 = T(0.5) * p.damp;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n * n; e += gridDim.x * blockDim.x) {
    const int j = e / n + 1, i = e % n + 1;
    const long oc = (long)(j + 2) * nd + (i + 2), ox = (long)(j + 2) * (nd + 1) + (i + 2);
    const long om = (long)(j - 1) * (n + 1) + (i - 1), on = (long)(j - 1) * n + (i - 1);
    const T dfx0 = damp2 * (m[oc - 1] + m[oc]) * fx2[ox], dfx1 = damp2 * (m[oc] + m[oc + 1]) * fx2[ox + 1];
    const T dfy0 = damp2 * (m[oc - nd] + m[oc]) * fy2[oc], dfy1 = damp2 * (m[oc] + m[oc + nd]) * fy2[oc + nd];
    const T dp2 = dp2_of<T>(m[oc], mul_rn(mx[om], frac), mul_rn(mx[om + 1], frac), mul_rn(my[on], frac), mul_rn(my[on + n], frac), ra[oc]);
    q[oc] = q[oc] + (dfx0 - dfx1 + dfy0 - dfy1) * ra[oc] / dp2;
  }
}
#endif

}  // namespace fv3t
