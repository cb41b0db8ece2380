// fv3atm_b200: marching horizontal tracer advection kernel (tracer_2d sub-step -> fv_tp_2d) for sm_100a.
//
// Replaces the k / iq loops of one sub-step of tracer_2d (atmos_cubed_sphere/model/fv_tracer2d.F90:503-556) and
// everything fv_tp_2d calls (model/tp_core.F90:110-249, xppm :332-704, yppm :707-1124, copy_corners :253-330,
// pert_ppm :1178-1236).  Work decomposition:
//   one CTA  = one strip of W columns (plus a 3-column halo on either side) of one (tile, level, tracer);
//   one thread = one column i of the strip; the CTA marches over the rows r = jsd..jed of the tile.
// y direction (both yppm sweeps): each thread keeps a rolling window of its column in registers -- dm, al, bl, br of
//   every cell are computed exactly once, with the tile-edge formulas selected by the row index.
// x direction (both xppm sweeps): the row is exchanged through shared memory (q, then dm/al), each x-face evaluates
//   bl/br of its upwind cell only; three barriers per row step serve the inner sweep of row r and the outer sweep of
//   row r-3 together.
// Pipeline at step r:  inner-y flux at face r-2 -> q_i(r-3) -> outer-x flux of row r-3 ;  inner-x flux of row r -> q_j(r)
//   -> outer-y flux at face r-2 ;  flux-form update of row r-3.  Nothing but q (read once, written once) and the
//   tracer-independent level fields moves through HBM; the 9 tracer CTAs of a (tile, level, strip) are adjacent in the
//   grid so that they march together and share cx, xfx, cy, yfx, mfx, mfy, dp1 through L2.
// Arithmetic: operation order of the reference, FMA-free (see fv3t_common.cuh).
#pragma once
#include "fv3t_advect.cuh"

namespace fv3t {

template <class T> struct Adv2Params {
  const T* qin;
  T* qout;
  const T* dp1;
  const T *mfx, *mfy, *cx, *cy;  // already scaled by frac = 1/ksplt(k) (k_prep), like the reference's in-place scaling
  const T *xfs, *yfs;            // xfx, yfx of fv_tracer2d.F90:392-405, scaled (scratch, same extents as cx, cy)
  GridDev<T> g;
  const int* ksplt;
  int n, npz, nq, ntiles, it, W;
  T lim_fac;
};

// ---------------------------------------------------------------------------------------------------------------------
// steps A and C of tracer_2d for one level field set (fv_tracer2d.F90:387-405, 449-486): xfx, yfx from the unscaled
// Courant numbers, then everything times frac = 1/ksplt(k).  cx, cy, mfx, mfy are scaled IN PLACE (the post-state
// the caller sees); with nsplt == 1 the reference skips the scaling (frac would be 1).
// ---------------------------------------------------------------------------------------------------------------------
template <class T>
__global__ void k_prep(T* __restrict__ cx, T* __restrict__ cy, T* __restrict__ mfx, T* __restrict__ mfy, T* __restrict__ xfs,
                       T* __restrict__ yfs, GridDev<T> g, const int* __restrict__ ksplt, int n, int npz, int ntiles, int scale) {
  const long nd = n + 6, plane = nd * nd;
  const long ncx = (long)(n + 1) * nd, nmf = (long)(n + 1) * n;
  const int lev = blockIdx.y;  // t*npz + kz
  const int t = lev / npz, kz = lev % npz;
  const T frac = T(1) / (T)ksplt[kz];
  const T* dxa = g.dxa + (long)t * plane;
  const T* dya = g.dya + (long)t * plane;
  const T* dxg = g.dx + (long)t * nd * (nd + 1);
  const T* dyg = g.dy + (long)t * (nd + 1) * nd;
  const T* ssg = g.sin_sg + (long)t * plane * 5;
  T* cxp = cx + (long)lev * ncx;
  T* cyp = cy + (long)lev * ncx;
  T* xfp = xfs + (long)lev * ncx;
  T* yfp = yfs + (long)lev * ncx;
  T* mxp = mfx + (long)lev * nmf;
  T* myp = mfy + (long)lev * nmf;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < ncx; e += stride) {
    {  // x face (i, j): i = 1..n+1, j = -2..n+3
      const int j = (int)(e / (n + 1)) - 2, i = (int)(e % (n + 1)) + 1;
      const T c = cxp[e];
      const long oc = (long)(j + 2) * nd + (i + 2), ody = (long)(j + 2) * (nd + 1) + (i + 2);
      T xf;
      if (c > T(0))
        xf = c * dxa[oc - 1] * dyg[ody] * ssg[2 * plane + oc - 1];
      else
        xf = c * dxa[oc] * dyg[ody] * ssg[0 * plane + oc];
      xfp[e] = xf * frac;
      if (scale) cxp[e] = c * frac;
    }
    {  // y face (i, j): i = -2..n+3, j = 1..n+1
      const int j = (int)(e / nd) + 1, i = (int)(e % nd) - 2;
      const T c = cyp[e];
      const long oc = (long)(j + 2) * nd + (i + 2);
      T yf;
      if (c > T(0))
        yf = c * dya[oc - nd] * dxg[oc] * ssg[3 * plane + oc - nd];
      else
        yf = c * dya[oc] * dxg[oc] * ssg[1 * plane + oc];
      yfp[e] = yf * frac;
      if (scale) cyp[e] = c * frac;
    }
    if (scale && e < nmf) {
      mxp[e] = mxp[e] * frac;
      myp[e] = myp[e] * frac;
    }
  }
}

// dp1 <- dp2 = dp1 + (mfx(i)-mfx(i+1)+mfy(j)-mfy(j+1))*rarea for the levels active in sub-step `it`
// (fv_tracer2d.F90:510-515, 547-553); launched only when it /= nsplt.
template <class T>
__global__ void k_dp1_update(T* __restrict__ dp1, const T* __restrict__ mfx, const T* __restrict__ mfy, const T* __restrict__ rarea,
                             const int* __restrict__ ksplt, int n, int npz, int it, int mode_1l = 0) {
  const long nd = n + 6, plane = nd * nd;
  const int lev = blockIdx.y;
  const int t = lev / npz, kz = lev % npz;
  if (it > ksplt[kz]) return;
  if (mode_1l && it >= ksplt[kz]) return;  // tracer_2d_1L: dp1 <- dp2 only between the level's OWN sub-steps (fv_tracer2d.F90:305)
  T* dp = dp1 + (long)lev * plane;
  const T* mx = mfx + (long)lev * (n + 1) * n;
  const T* my = mfy + (long)lev * (n + 1) * n;
  const T* ra = rarea + (long)t * plane;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < (long)n * n; e += stride) {
    const int j = (int)(e / n) + 1, i = (int)(e % n) + 1;
    const long oc = (long)(j + 2) * nd + (i + 2);
    const long ox = (long)(j - 1) * (n + 1) + (i - 1), oy = (long)(j - 1) * n + (i - 1);
    dp[oc] = dp[oc] + (mx[ox] - mx[ox + 1] + my[oy] - my[oy + n]) * ra[oc];
  }
}

// monotone slope dm of cell i from q(i-1), q(i), q(i+1)  (tp_core.F90:563-567)
// Evaluated as min(|xt|, max(min(hi - q0, q0 - lo), 0)) with (lo, hi) = minmax(qm, qp): bit-identical to the reference
// expression min(|xt|, max(qm,q0,qp) - q0, q0 - min(qm,q0,qp)) (when q0 lies outside [lo, hi] the reference yields +0
// through q0 - q0) with three comparisons instead of six.
template <class T> FV3T_HD T dm_of(T qm, T q0, T qp) {
  const T xt = T(0.25) * (qp - qm);
  const bool up = qm < qp;
  const T lo = up ? qm : qp, hi = up ? qp : qm;
  return dm_limit<T>(xt, f_min(hi - q0, q0 - lo));
}

// two-sided edge value between cells (a, b | c, d) with metric (ma, mb | mc, md)  (tp_core.F90:384-385, 640-641)
template <class T> FV3T_HD T edge_value4(T qa, T qb, T qc, T qd, T ma, T mb, T mc, T md) {
  return T(0.5) * (((T(2) * mb + ma) * qb - mb * qa) / (ma + mb) + ((T(2) * mc + md) * qc - mc * qd) / (mc + md));
}

// ---------------------------------------------------------------------------------------------------------------------
// Rolling 1-D PPM state of one column (yppm restated cell by cell).  When row r arrives the cell c = r-2 is
// reconstructed (its stencil q(c-2..c+2) is complete) and the flux at face c (between cells c-1 and c) is returned.
// ---------------------------------------------------------------------------------------------------------------------
// One step of the rolling 1-D PPM reconstruction of a column (yppm restated cell by cell): with q(c-2..c+2) and the carried
// quantities of cells c-1, c in hand, reconstruct cell c and return the flux at face c (between cells c-1 and c).
// EDGE = false compiles the tile-edge formulas (and the tests that select them) out: the caller guarantees 3 <= c <= npx-3
// (ORD >= 7) resp. 3 <= c+1 <= npx-2 (ORD < 7).  Outputs: a_p1 = dm(c+1) / al(c+1), al_p1 = al(c+1) (ORD >= 7), bl/br/fl of
// cell c; xt, xt2 carry the tile-edge values across the three edge cells.
template <class T, int ORD, bool EDGE, class MF>
FV3T_HD T ystream_core(int c, T qm2, T qm1, T q0, T qp1, T qp2, T a_m1, T a_0, T al_0, T bl_m1, T br_m1, int fl_m1, T& xt, T& xt2,
                       T cour, int npx, T lim_fac, MF met, T& a_p1, T& al_p1, T& bl, T& br, int& fl) {
  constexpr int mord = ORD < 0 ? -ORD : ORD;
  bl = T(0);
  br = T(0);
  fl = 0;
  al_p1 = T(0);
  if (ORD >= 7) {
    a_p1 = dm_of<T>(q0, qp1, qp2);
    al_p1 = T(0.5) * (q0 + qp1) + K<T>::r3() * (a_0 - a_p1);
    if (!EDGE || (c >= 3 && c <= npx - 3)) {
      const T qm = qm1, qp = qp1, dm0 = a_0;
      const T al0 = al_0, al1 = al_p1;
      if (ORD == 8 || ORD == 11) {
        const T x = (ORD == 8 ? T(2) : K<T>::ppm_fac()) * dm0;
        bl = -sign_min_abs<T>(x, al0 - q0);
        br = sign_min_abs<T>(x, al1 - q0);
      } else if (ORD == 10) {
        bl = al0 - q0;
        br = al1 - q0;
        if (f_abs(a_m1) + f_abs(dm0) + f_abs(a_p1) < K<T>::near_zero()) {
          bl = T(0);
          br = T(0);
        } else if (f_abs(T(3) * (bl + br)) > f_abs(bl - br)) {
          const T dq_m2 = T(2) * (qm - qm2);
          const T dq_m1 = T(2) * (q0 - qm);
          const T dq_0 = T(2) * (qp - q0);
          const T dq_p1 = T(2) * (qp2 - qp);
          const T pmp_2 = dq_m1;
          const T lac_2 = pmp_2 - T(0.75) * dq_m2;
          br = f_min(f_max(T(0), pmp_2, lac_2), f_max(br, f_min(T(0), pmp_2, lac_2)));
          const T pmp_1 = -dq_0;
          const T lac_1 = pmp_1 + T(0.75) * dq_p1;
          bl = f_min(f_max(T(0), pmp_1, lac_1), f_max(bl, f_min(T(0), pmp_1, lac_1)));
        }
      } else if (ORD == 7 || ORD == 12) {
        bl = al0 - q0;
        br = al1 - q0;
        const T a4 = T(-3) * (bl + br);
        const T da1 = br - bl;
        const bool ext5 = br * bl > T(0);
        const bool ext6 = f_abs(da1) < -a4;
        if (ext6) {
          if (q0 + T(0.25) / a4 * (da1 * da1) + a4 * K<T>::r12() < T(0)) {
            if (ext5) {
              br = T(0);
              bl = T(0);
            } else if (da1 > T(0)) {
              br = T(-2) * bl;
            } else {
              bl = T(-2) * br;
            }
          }
        }
      } else {
        bl = al0 - q0;
        br = al1 - q0;
        if (ORD == 9 || ORD == 13) pert_ppm1<T>(q0, bl, br, 0);
      }
    } else if (c >= 0 && c <= npx) {
      // tile-edge cells 0,1,2 and npx-2,npx-1,npx (tp_core.F90:636-674 with i <-> j)
      if (c == 0 || c == npx - 1) {
        // e0 = c+1: couples cells (c-1, c | c+1, c+2)
        T x = edge_value4<T>(qm1, q0, qp1, qp2, met(c - 1), met(c), met(c + 1), met(c + 2));
        x = f_max(x, f_min(qm1, q0, qp1, qp2));
        x = f_min(x, f_max(qm1, q0, qp1, qp2));
        if (c == 0) {
          bl = K<T>::s14() * a_m1 + K<T>::s11() * (qm1 - q0);
          br = x - q0;
        } else {
          bl = xt2 - q0;
          br = x - q0;
        }
        xt = x;
      } else if (c == 1) {
        xt2 = K<T>::s15() * q0 + K<T>::s11() * qp1 - K<T>::s14() * a_p1;
        bl = xt - q0;
        br = xt2 - q0;
      } else if (c == 2) {
        bl = xt2 - q0;
        br = al_p1 - q0;
      } else if (c == npx - 2) {
        xt2 = K<T>::s15() * qp1 + K<T>::s11() * q0 + K<T>::s14() * a_0;
        bl = al_0 - q0;
        br = xt2 - q0;
      } else {  // c == npx
        bl = xt - q0;
        br = K<T>::s11() * (qp1 - q0) - K<T>::s14() * a_p1;
      }
      pert_ppm1<T>(q0, bl, br, 1);
    }
  } else {
    // al(c+1) (tp_core.F90:377-400 with i <-> j)
    const int f = c + 1;
    if (EDGE && (f == 0 || f == npx - 1))
      a_p1 = K<T>::c1() * qm1 + K<T>::c2() * q0 + K<T>::c3() * qp1;
    else if (EDGE && (f == 1 || f == npx))
      a_p1 = edge_value4<T>(qm1, q0, qp1, qp2, met(c - 1), met(c), met(c + 1), met(c + 2));
    else if (EDGE && (f == 2 || f == npx + 1))
      a_p1 = K<T>::c3() * q0 + K<T>::c2() * qp1 + K<T>::c1() * qp2;
    else
      a_p1 = K<T>::p1() * (q0 + qp1) + K<T>::p2() * (qm1 + qp2);
    if (ORD < 0) a_p1 = f_max(T(0), a_p1);
    bl = a_0 - q0;
    br = a_p1 - q0;
    const T b0 = bl + br;
    if (mord == 1) {
      fl = f_abs(lim_fac * b0) < f_abs(bl - br);
    } else if (mord == 3 || mord == 4) {
      const T x0 = f_abs(b0);
      const T x1 = f_abs(bl - br);
      fl = (x0 < x1 ? 1 : 0) | (T(3) * x0 < x1 ? 2 : 0);
    } else if (ORD == 5) {
      fl = bl * br < T(0);
    } else if (ORD == -5) {
      fl = bl * br < T(0);
      const T da1 = br - bl;
      const T a4 = T(-3) * b0;
      if (f_abs(da1) < -a4) {
        if (q0 + T(0.25) * (da1 * da1) / a4 + a4 * K<T>::r12() < T(0)) {
          if (!fl) {
            br = T(0);
            bl = T(0);
          } else if (da1 > T(0)) {
            br = T(-2) * bl;
          } else {
            bl = T(-2) * br;
          }
        }
      }
    } else if (mord != 2) {
      fl = f_abs(T(3) * b0) < f_abs(bl - br);
    }
  }
  // flux at face c
  if (ORD >= 8) {
    // ppm_flux's single expression for both wind directions, with the upwind cell picked by three selects instead of index
    // comparisons in accessor lambdas (which the compiler turned into ~15 predicated moves per face)
    const bool up = cour > T(0);
    const T a = f_abs(cour);
    const T blu = up ? bl_m1 : bl, bru = up ? br_m1 : br, qu = up ? qm1 : q0;
    const T b = up ? bru : blu;
    return qu + (T(1) - a) * (b - a * (blu + bru));
  }
  const T qm1_ = qm1, q0_ = q0, blm = bl_m1, brm = br_m1;
  const T bl_ = bl, br_ = br;
  const int flm = fl_m1, fl_ = fl;
  const T alm = a_m1, al0_ = a_0, alp = a_p1;
  auto fq = [&](int gi) -> T { return gi == c ? q0_ : qm1_; };
  auto fbl = [&](int gi) -> T { return gi == c ? bl_ : blm; };
  auto fbr = [&](int gi) -> T { return gi == c ? br_ : brm; };
  auto ffl = [&](int gi) -> int { return gi == c ? fl_ : flm; };
  auto fal = [&](int gi) -> T { return gi == c ? al0_ : (gi == c - 1 ? alm : alp); };
  return ppm_flux<T, ORD>(c, cour, fq, fbl, fbr, ffl, fal);
}

// ---------------------------------------------------------------------------------------------------------------------
// Rolling 1-D PPM state of one column (yppm restated cell by cell).  When row r arrives the cell c = r-2 is
// reconstructed (its stencil q(c-2..c+2) is complete) and the flux at face c (between cells c-1 and c) is returned.
// ---------------------------------------------------------------------------------------------------------------------
template <class T, int ORD> struct YStream {
  T qm2, qm1, q0, qp1;  // q(c-2), q(c-1), q(c), q(c+1)
  T a_m1, a_0;          // ORD >= 7: dm(c-1), dm(c);  ORD < 7: al(c-1), al(c)
  T al_0;               // ORD >= 7: al(c)
  T bl_m1, br_m1;       // cell c-1
  int fl_m1;
  T xt, xt2;            // tile-edge values carried across the three edge cells

  FV3T_HD void init() {
    qm2 = qm1 = q0 = qp1 = a_m1 = a_0 = al_0 = bl_m1 = br_m1 = xt = xt2 = T(0);
    fl_m1 = 0;
  }

  // met(row) = metric along the sweep (dya(i,row)), read only at the tile edges.  q_cm1 returns q(c-1).
  template <class MF> FV3T_HD T push(int c, T qp2, T cour, int npx, T lim_fac, MF met, T& q_cm1) {
    T a_p1, al_p1, bl, br;
    int fl;
    const T flux = ystream_core<T, ORD, true>(c, qm2, qm1, q0, qp1, qp2, a_m1, a_0, al_0, bl_m1, br_m1, fl_m1, xt, xt2, cour, npx, lim_fac,
                                              met, a_p1, al_p1, bl, br, fl);
    q_cm1 = qm1;
    // shift to cell c+1
    qm2 = qm1;
    qm1 = q0;
    q0 = qp1;
    qp1 = qp2;
    a_m1 = a_0;
    a_0 = a_p1;
    al_0 = al_p1;
    bl_m1 = bl;
    br_m1 = br;
    fl_m1 = fl;
    return flux;
  }
};

// The same stream with its windows held in phase-indexed slots: step PH = (row step) mod 4 reads q(c-2..c+1) from slots
// PH..PH+3 (mod 4) and overwrites slot PH with the new row, so that a row loop unrolled by four rotates the windows by
// renaming only -- no register moves (they were 88 of the 710 instructions per cell-update of k_advect4,
// profiles/r01_advect4_c384_ncu.txt).
template <class T, int ORD> struct YWin {
  T q[4];
  T a[2];
  T al_0, bl_m1, br_m1;
  int fl_m1;
  FV3T_HD void init() {
    for (int k = 0; k < 4; ++k) q[k] = T(0);
    a[0] = a[1] = al_0 = bl_m1 = br_m1 = T(0);
    fl_m1 = 0;
  }
  // q(c-1) before the push = q(o) of the marching kernels' output row after it
  template <int PH> FV3T_HD T q_cm1() const { return q[(PH + 1) & 3]; }
  // edge2: two per-thread scratch slots (stride `es` elements) that carry the tile-edge values xt, xt2 across the three edge
  // cells; touched only by the EDGE instantiation, so the interior row loop does not keep them in registers
  template <int PH, bool EDGE, class MF> FV3T_HD T push(int c, T qp2, T cour, int npx, T lim_fac, MF met, T* edge2, int es) {
    T a_p1, al_p1, bl, br;
    int fl;
    T xt = T(0), xt2 = T(0);
    if (EDGE) {
      xt = edge2[0];
      xt2 = edge2[es];
    }
    const T flux = ystream_core<T, ORD, EDGE>(c, q[PH & 3], q[(PH + 1) & 3], q[(PH + 2) & 3], q[(PH + 3) & 3], qp2, a[PH & 1], a[(PH + 1) & 1],
                                              al_0, bl_m1, br_m1, fl_m1, xt, xt2, cour, npx, lim_fac, met, a_p1, al_p1, bl, br, fl);
    if (EDGE) {
      edge2[0] = xt;
      edge2[es] = xt2;
    }
    q[PH & 3] = qp2;
    a[PH & 1] = a_p1;
    al_0 = al_p1;
    bl_m1 = bl;
    br_m1 = br;
    fl_m1 = fl;
    return flux;
  }
};

// flux at x-face i from a row held in shared memory: q(gi), a(gi) (dm for ORD >= 7, al for ORD < 7) by global index
template <class T, int ORD, bool EDGE = true, class QF, class AF, class DF>
FV3T_HD T xface_flux(int i, T cour, int npx, T lim_fac, QF q, AF a, DF dxa) {
  if (ORD >= 8) {
    const bool up = cour > T(0);
    const int u = up ? i - 1 : i;
    T bl, br;
    int flg;
    ppm_blbr<T, ORD, EDGE>(u, npx, q, a, dxa, lim_fac, bl, br, flg);
    const T qu = q(u);
    const T a = f_abs(cour);  // see ppm_flux: one expression for both wind directions, bit-identical to the reference's two
    return qu + (T(1) - a) * ((up ? br : bl) - a * (bl + br));
  } else {
    T blm, brm, bl0, br0;
    int fm, f0;
    ppm_blbr<T, ORD, EDGE>(i - 1, npx, q, a, dxa, lim_fac, blm, brm, fm);
    ppm_blbr<T, ORD, EDGE>(i, npx, q, a, dxa, lim_fac, bl0, br0, f0);
    auto fbl = [&](int gi) -> T { return gi == i ? bl0 : blm; };
    auto fbr = [&](int gi) -> T { return gi == i ? br0 : brm; };
    auto ffl = [&](int gi) -> int { return gi == i ? f0 : fm; };
    return ppm_flux<T, ORD>(i, cour, q, fbl, fbr, ffl, a);
  }
}

template <class T, int ORD_IN, int ORD_OU> __global__ void __launch_bounds__(256) k_advect2(const Adv2Params<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw2[];
  const int NT = blockDim.x;
  T* sqa = reinterpret_cast<T*>(smem_raw2);  // q of row r          (inner x sweep)
  T* sda = sqa + NT;                         // dm / al of row r
  T* sqb = sda + NT;                         // q_i of row r-3      (outer x sweep)
  T* sdb = sqb + NT;
  T* sf1 = sdb + NT;                         // xfx*fx2 of row r
  T* sft = sf1 + NT;                         // 0.5*(fx+fx2)*mfx of row r-3

  const int n = p.n, npz = p.npz, npx = n + 1;
  const int iq = blockIdx.x, strip = blockIdx.y;
  const int t = blockIdx.z / npz, kz = blockIdx.z % npz;
  if (p.it > p.ksplt[kz]) return;
  const int tid = threadIdx.x;
  const int i0 = 1 + strip * p.W;
  const int nw = min(p.W, n - i0 + 1);
  const int i = i0 - 3 + tid;                         // this thread's column
  const bool col_ok = tid < nw + 6;                   // takes part in the y sweeps of q
  const bool face_ok = tid >= 3 && tid <= nw + 3;     // x-face i (i0 .. i1+1)
  const bool cell_ok = tid >= 3 && tid < nw + 3;      // compute column (i0 .. i1)
  const bool pre_ok = (ORD_IN >= 7) ? (tid >= 1 && tid <= nw + 4) : (tid >= 2 && tid <= nw + 4);
  const bool pre_ok_b = (ORD_OU >= 7) ? (tid >= 1 && tid <= nw + 4) : (tid >= 2 && tid <= nw + 4);
  const long nd = n + 6, plane = nd * nd;
  const long lev = (long)t * npz + kz;
  const T* qg = p.qin + (((long)t * p.nq + iq) * npz + kz) * plane;
  T* qo = p.qout + (((long)t * p.nq + iq) * npz + kz) * plane;
  const T* area = p.g.area + (long)t * plane;
  const T* rarea = p.g.rarea + (long)t * plane;
  const T* dxa = p.g.dxa + (long)t * plane;
  const T* dya = p.g.dya + (long)t * plane;
  const T* cxp = p.cx + lev * (long)(n + 1) * nd;
  const T* xfp = p.xfs + lev * (long)(n + 1) * nd;
  const T* cyp = p.cy + lev * nd * (long)(n + 1);
  const T* yfp = p.yfs + lev * nd * (long)(n + 1);
  const T* mfxp = p.mfx + lev * (long)(n + 1) * n;
  const T* mfyp = p.mfy + lev * (long)n * (n + 1);
  const T* dp1p = p.dp1 + lev * plane;
  const int gb = i0 - 3;  // global column of shared-memory slot 0
  const bool icor = (i < 1) || (i > n);

  YStream<T, ORD_IN> yin;
  YStream<T, ORD_OU> you;
  yin.init();
  you.init();
  T fx2_a = T(0), fx2_b = T(0), fx2_c = T(0);  // fx2(i, r-1), (r-2), (r-3)
  T fy2_prev = T(0), fyo_prev = T(0), fyy_prev = T(0), yf_prev = T(0), mfy_prev = T(0);

  auto met_y = [&](int row) -> T { return dya[(long)(row + 2) * nd + (i + 2)]; };

  for (int r = -2; r <= n + 3; ++r) {
    const int c = r - 2;  // y-face / cell completed by this row
    const int o = r - 3;  // output row
    const bool o_ok = o >= 1 && o <= n;
    const bool c_ok = c >= 1 && c <= n + 1;
    const long orow = (long)(r + 2) * nd + (i + 2);
    // ---- (a) this row of q: the x sweeps see the dir = 1 corner view, the y sweeps the dir = 2 view
    T qx = T(0), qy = T(0);
    if (col_ok) {
      if (icor && (r < 1 || r > n)) {
        int s1i, s1j, s2i, s2j;
        if (i < 1 && r < 1) {  // SW
          s1i = r, s1j = 1 - i, s2i = 1 - r, s2j = i;
        } else if (i > n && r < 1) {  // SE
          s1i = npx - r, s1j = i - npx + 1, s2i = npx + r - 1, s2j = npx - i;
        } else if (i > n) {  // NE
          s1i = r, s1j = 2 * npx - 1 - i, s2i = 2 * npx - 1 - r, s2j = i;
        } else {  // NW
          s1i = npx - r, s1j = i - 1 + npx, s2i = r + 1 - npx, s2j = npx - i;
        }
        qx = qg[(long)(s1j + 2) * nd + (s1i + 2)];
        qy = qg[(long)(s2j + 2) * nd + (s2i + 2)];
      } else {
        qx = qg[orow];
        qy = qx;
      }
    }
    // ---- (d) inner y sweep: flux at face c, then q_i of row o
    T cyv = T(0), yfv = T(0), fy2_c = T(0), fyy_c = T(0), q_o = T(0), qi = T(0);
    if (col_ok) {
      if (c_ok) {
        const long ocy = (long)(c - 1) * nd + (i + 2);
        cyv = cyp[ocy];
        yfv = yfp[ocy];
      }
      fy2_c = yin.push(c, qy, cyv, npx, p.lim_fac, met_y, q_o);
      fyy_c = yfv * fy2_c;
      if (o_ok) {
        const T ar = area[(long)(o + 2) * nd + (i + 2)];
        const T ra_y = ar + yf_prev - yfv;
        qi = (q_o * ar + fyy_prev - fyy_c) / ra_y;
      }
      sqa[tid] = qx;
      sqb[tid] = qi;
    }
    __syncthreads();
    // ---- phase 2: dm (ORD >= 7) or al (ORD < 7) of both rows
    {
      auto qa = [&](int gi) -> T { return sqa[gi - gb]; };
      auto qb = [&](int gi) -> T { return sqb[gi - gb]; };
      auto dxa_r = [&](int gi) -> T { return dxa[(long)(r + 2) * nd + (gi + 2)]; };
      auto dxa_o = [&](int gi) -> T { return dxa[(long)(o + 2) * nd + (gi + 2)]; };
      if (pre_ok) sda[tid] = ppm_pre<T, ORD_IN>(i, npx, qa, dxa_r);
      if (pre_ok_b && o_ok) sdb[tid] = ppm_pre<T, ORD_OU>(i, npx, qb, dxa_o);
    }
    __syncthreads();
    // ---- phase 3: x-face fluxes of both rows
    T xf_i = T(0), ar_r = T(0);
    if (face_ok) {
      auto qa = [&](int gi) -> T { return sqa[gi - gb]; };
      auto aa = [&](int gi) -> T { return sda[gi - gb]; };
      auto qb = [&](int gi) -> T { return sqb[gi - gb]; };
      auto ab = [&](int gi) -> T { return sdb[gi - gb]; };
      auto dxa_r = [&](int gi) -> T { return dxa[(long)(r + 2) * nd + (gi + 2)]; };
      auto dxa_o = [&](int gi) -> T { return dxa[(long)(o + 2) * nd + (gi + 2)]; };
      const long ocx = (long)(r + 2) * (n + 1) + (i - 1);
      const T cxv = cxp[ocx];
      xf_i = xfp[ocx];
      const T fx2 = xface_flux<T, ORD_IN>(i, cxv, npx, p.lim_fac, qa, aa, dxa_r);
      sf1[tid] = xf_i * fx2;
      if (o_ok) {
        const T cxo = cxp[(long)(o + 2) * (n + 1) + (i - 1)];
        const T fxo = xface_flux<T, ORD_OU>(i, cxo, npx, p.lim_fac, qb, ab, dxa_o);
        sft[tid] = T(0.5) * (fxo + fx2_c) * mfxp[(long)(o - 1) * (n + 1) + (i - 1)];
      }
      fx2_c = fx2_b;
      fx2_b = fx2_a;
      fx2_a = fx2;
    }
    __syncthreads();
    // ---- phase 4: q_j of row r, outer y flux at face c, flux-form update of row o
    if (cell_ok) {
      ar_r = area[orow];
      const T xf_ip = xfp[(long)(r + 2) * (n + 1) + i];
      const T ra_x = ar_r + xf_i - xf_ip;
      const T qj = (qx * ar_r + sf1[tid] - sf1[tid + 1]) / ra_x;
      T dummy;
      const T fyo_c = you.push(c, qj, cyv, npx, p.lim_fac, met_y, dummy);
      T mfy_c = T(0);
      if (c_ok) mfy_c = mfyp[(long)(c - 1) * n + (i - 1)];
      if (o_ok) {
        const long oo = (long)(o + 2) * nd + (i + 2);
        const T rar = rarea[oo];
        const T dp1v = dp1p[oo];
        const T mfx0 = mfxp[(long)(o - 1) * (n + 1) + (i - 1)];
        const T mfx1 = mfxp[(long)(o - 1) * (n + 1) + i];
        const T dp2 = dp1v + (mfx0 - mfx1 + mfy_prev - mfy_c) * rar;
        const T fxa = sft[tid], fxb = sft[tid + 1];
        const T fya = T(0.5) * (fyo_prev + fy2_prev) * mfy_prev;
        const T fyb = T(0.5) * (fyo_c + fy2_c) * mfy_c;
        qo[oo] = (q_o * dp1v + (fxa - fxb + fya - fyb) * rar) / dp2;
      }
      fyo_prev = fyo_c;
      mfy_prev = mfy_c;
    }
    fy2_prev = fy2_c;
    fyy_prev = fyy_c;
    yf_prev = yfv;
  }
}

}  // namespace fv3t
