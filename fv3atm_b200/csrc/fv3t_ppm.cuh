// fv3atm_b200: 1-D PPM building blocks shared by the x- and y-sweeps of the Lin-Rood operator.
//
// Behaviour follows xppm / yppm / pert_ppm of the reference (atmos_cubed_sphere/model/tp_core.F90:332-704,
// 707-1124, 1178-1236); the two reference routines differ only by i<->j, so one set of element-wise device
// functions serves both.  The functions are written per element (one cell or one face) so that a CTA can
// evaluate a whole shared-memory tile in three barrier-separated phases:
//    pre   : dm(i) (ORD >= 7)  or  al(i) (ORD < 7)
//    blbr  : bl(i), br(i) [+ limiter flags for ORD < 7]
//    flux  : upwind flux at face i
// `q`, `a`, `dxa` are accessors taking the *global* 1-based index along the sweep; npx is the tile extent
// (cells 1..npx-1), and the cubed-sphere tile-edge formulas (tp_core.F90:381-394, 636-674) are selected by
// index, exactly where the reference applies them for a rank that owns a whole tile edge.
#pragma once
#include "fv3t_common.cuh"

namespace fv3t {

// pert_ppm (tp_core.F90:1178-1236), one element
template <class T> FV3T_HD void pert_ppm1(T a0, T& al, T& ar, int iv) {
  if (iv == 0) {
    if (a0 <= T(0)) {
      al = T(0);
      ar = T(0);
    } else {
      const T a4 = T(-3) * (ar + al);
      const T da1 = ar - al;
      if (f_abs(da1) < -a4) {
        const T fmin = a0 + T(0.25) / a4 * (da1 * da1) + a4 * K<T>::r12();
        if (fmin < T(0)) {
          if (ar > T(0) && al > T(0)) {
            ar = T(0);
            al = T(0);
          } else if (da1 > T(0)) {
            ar = T(-2) * al;
          } else {
            al = T(-2) * ar;
          }
        }
      }
    }
  } else {
    if (al * ar < T(0)) {
      const T da1 = al - ar;
      const T da2 = da1 * da1;
      const T a6da = T(3) * (al + ar) * da1;
      if (a6da < -da2) {
        ar = T(-2) * al;
      } else if (a6da > da2) {
        al = T(-2) * ar;
      }
    } else {
      al = T(0);
      ar = T(0);
    }
  }
}

// two-sided edge value at a tile edge (tp_core.F90:384-385, 640-641); e0 = index of the first cell
// inside the tile edge's far side, i.e. the formula couples cells (e0-2, e0-1 | e0, e0+1)
template <class T, class QF, class DF> FV3T_HD T edge_value(int e0, QF q, DF dxa) {
  return T(0.5) * (((T(2) * dxa(e0 - 1) + dxa(e0 - 2)) * q(e0 - 1) - dxa(e0 - 1) * q(e0 - 2)) / (dxa(e0 - 2) + dxa(e0 - 1)) +
                   ((T(2) * dxa(e0) + dxa(e0 + 1)) * q(e0) - dxa(e0) * q(e0 + 1)) / (dxa(e0) + dxa(e0 + 1)));
}

// phase "pre": dm(i) for ORD >= 7 (tp_core.F90:563-567), al(i) for ORD < 7 (:377-400)
template <class T, int ORD, bool EDGE = true, class QF, class DF> FV3T_HD T ppm_pre(int i, int npx, QF q, DF dxa) {
  if (ORD >= 7) {
    // = sign(min(|xt|, max(qm,q0,qp) - q0, q0 - min(qm,q0,qp)), xt), bit-identical, three comparisons (see dm_of)
    const T qm = q(i - 1), q0 = q(i), qp = q(i + 1);
    const T xt = T(0.25) * (qp - qm);
    const bool up = qm < qp;
    const T lo = up ? qm : qp, hi = up ? qp : qm;
    return dm_limit<T>(xt, f_min(hi - q0, q0 - lo));
  } else {
    T al;
    if (!EDGE) {
      al = K<T>::p1() * (q(i - 1) + q(i)) + K<T>::p2() * (q(i - 2) + q(i + 1));
    } else if (i == 0) {
      al = K<T>::c1() * q(-2) + K<T>::c2() * q(-1) + K<T>::c3() * q(0);
    } else if (i == 1) {
      al = edge_value<T>(1, q, dxa);
    } else if (i == 2) {
      al = K<T>::c3() * q(1) + K<T>::c2() * q(2) + K<T>::c1() * q(3);
    } else if (i == npx - 1) {
      al = K<T>::c1() * q(npx - 3) + K<T>::c2() * q(npx - 2) + K<T>::c3() * q(npx - 1);
    } else if (i == npx) {
      al = edge_value<T>(npx, q, dxa);
    } else if (i == npx + 1) {
      al = K<T>::c3() * q(npx) + K<T>::c2() * q(npx + 1) + K<T>::c1() * q(npx + 2);
    } else {
      al = K<T>::p1() * (q(i - 1) + q(i)) + K<T>::p2() * (q(i - 2) + q(i + 1));
    }
    if (ORD < 0) al = f_max(T(0), al);
    return al;
  }
}

// phase "blbr" for cell i.  `a` = dm (ORD >= 7) or al (ORD < 7).  flg: bit0 smt5, bit1 smt6 (ORD < 7 only).
template <class T, int ORD, bool EDGE = true, class QF, class AF, class DF>
FV3T_HD void ppm_blbr(int i, int npx, QF q, AF a, DF dxa, T lim_fac, T& bl, T& br, int& flg) {
  flg = 0;
  const T q0 = q(i);
  if (ORD < 7) {
    constexpr int mord = ORD < 0 ? -ORD : ORD;
    bl = a(i) - q0;
    br = a(i + 1) - q0;
    const T b0 = bl + br;
    if (mord == 1) {
      flg = f_abs(lim_fac * b0) < f_abs(bl - br);
    } else if (mord == 3 || mord == 4) {
      const T x0 = f_abs(b0);
      const T xt = f_abs(bl - br);
      flg = (x0 < xt ? 1 : 0) | (T(3) * x0 < xt ? 2 : 0);
    } else if (ORD == 5) {
      flg = bl * br < T(0);
    } else if (ORD == -5) {
      flg = bl * br < T(0);
      const T da1 = br - bl;
      const T a4 = T(-3) * b0;
      if (f_abs(da1) < -a4) {
        if (q0 + T(0.25) * (da1 * da1) / a4 + a4 * K<T>::r12() < T(0)) {
          if (!flg) {
            br = T(0);
            bl = T(0);
          } else if (da1 > T(0)) {
            br = T(-2) * bl;
          } else {
            bl = T(-2) * br;
          }
        }
      }
    } else if (mord != 2) {  // ORD == 6 (and any other value < 7 falls in the reference's final else)
      flg = f_abs(T(3) * b0) < f_abs(bl - br);
    }
    return;
  }
  // ---- ORD >= 7 ----
  if (EDGE && i <= 2) {  // west / south tile edge, cells 0,1,2 (tp_core.F90:637-656)
    T xt = edge_value<T>(1, q, dxa);
    xt = f_max(xt, f_min(q(-1), q(0), q(1), q(2)));
    xt = f_min(xt, f_max(q(-1), q(0), q(1), q(2)));
    const T xt2 = K<T>::s15() * q(1) + K<T>::s11() * q(2) - K<T>::s14() * a(2);
    if (i == 0) {
      bl = K<T>::s14() * a(-1) + K<T>::s11() * (q(-1) - q(0));
      br = xt - q(0);
    } else if (i == 1) {
      bl = xt - q(1);
      br = xt2 - q(1);
    } else {
      const T al3 = T(0.5) * (q(2) + q(3)) + K<T>::r3() * (a(2) - a(3));
      bl = xt2 - q(2);
      br = al3 - q(2);
    }
    pert_ppm1<T>(q0, bl, br, 1);
    return;
  }
  if (EDGE && i >= npx - 2) {  // east / north tile edge, cells npx-2, npx-1, npx (tp_core.F90:657-674)
    const T xt2 = K<T>::s15() * q(npx - 1) + K<T>::s11() * q(npx - 2) + K<T>::s14() * a(npx - 2);
    T xt = edge_value<T>(npx, q, dxa);
    xt = f_max(xt, f_min(q(npx - 2), q(npx - 1), q(npx), q(npx + 1)));
    xt = f_min(xt, f_max(q(npx - 2), q(npx - 1), q(npx), q(npx + 1)));
    if (i == npx - 2) {
      const T aln = T(0.5) * (q(npx - 3) + q(npx - 2)) + K<T>::r3() * (a(npx - 3) - a(npx - 2));
      bl = aln - q(npx - 2);
      br = xt2 - q(npx - 2);
    } else if (i == npx - 1) {
      bl = xt2 - q(npx - 1);
      br = xt - q(npx - 1);
    } else {
      bl = xt - q(npx);
      br = K<T>::s11() * (q(npx + 1) - q(npx)) - K<T>::s14() * a(npx + 1);
    }
    pert_ppm1<T>(q0, bl, br, 1);
    return;
  }
  // interior cells 3 .. npx-3
  const T qm = q(i - 1), qp = q(i + 1);
  const T dm0 = a(i);
  const T al0 = T(0.5) * (qm + q0) + K<T>::r3() * (a(i - 1) - dm0);
  const T al1 = T(0.5) * (q0 + qp) + K<T>::r3() * (dm0 - a(i + 1));
  if (ORD == 8 || ORD == 11) {
    const T xt = (ORD == 8 ? T(2) : K<T>::ppm_fac()) * dm0;
    bl = -sign_min_abs<T>(xt, al0 - q0);
    br = sign_min_abs<T>(xt, al1 - q0);
  } else if (ORD == 10) {
    bl = al0 - q0;
    br = al1 - q0;
    if (f_abs(a(i - 1)) + f_abs(dm0) + f_abs(a(i + 1)) < K<T>::near_zero()) {
      bl = T(0);
      br = T(0);
    } else if (f_abs(T(3) * (bl + br)) > f_abs(bl - br)) {
      const T dq_m2 = T(2) * (qm - q(i - 2));  // dq(i-2)
      const T dq_m1 = T(2) * (q0 - qm);        // dq(i-1)
      const T dq_0 = T(2) * (qp - q0);         // dq(i)
      const T dq_p1 = T(2) * (q(i + 2) - qp);  // dq(i+1)
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
  } else {  // 9, 13 (and the reference's plain else)
    bl = al0 - q0;
    br = al1 - q0;
    if (ORD == 9 || ORD == 13) pert_ppm1<T>(q0, bl, br, 0);
  }
}

// phase "flux" at face i (between cells i-1 and i), Courant number c (tp_core.F90:402-554, 678-701)
template <class T, int ORD, class QF, class BLF, class BRF, class GF, class AF>
FV3T_HD T ppm_flux(int i, T c, QF q, BLF bl, BRF br, GF flg, AF al) {
  constexpr int mord = ORD < 0 ? -ORD : ORD;
  if (ORD >= 8) {
    // one evaluation for both wind directions: with a = |c| the upwind cell u gives q(u) + (1-a)*(b - a*(bl+br)), b = br for
    // c > 0 and bl otherwise -- bit-identical to the reference's two expressions (1+c = 1-a and bl + c*s = bl - a*s exactly)
    const bool up = c > T(0);
    const T a = f_abs(c);
    const T blu = up ? bl(i - 1) : bl(i), bru = up ? br(i - 1) : br(i), qu = up ? q(i - 1) : q(i);
    const T b = up ? bru : blu;
    return qu + (T(1) - a) * (b - a * (blu + bru));
  } else if (ORD == 7) {
    const T blm = bl(i - 1), brm = br(i - 1), bl0 = bl(i), br0 = br(i);
    const bool s = (blm * brm < T(0)) || (bl0 * br0 < T(0));
    T fx1, flux;
    if (c > T(0)) {
      fx1 = (T(1) - c) * (brm - c * (blm + brm));
      flux = q(i - 1);
    } else {
      fx1 = (T(1) + c) * (bl0 + c * (bl0 + br0));
      flux = q(i);
    }
    if (s) flux = flux + fx1;
    return flux;
  } else if (mord == 2) {
    if (c > T(0)) {
      const T qtmp = q(i - 1);
      return qtmp + (T(1) - c) * (al(i) - qtmp - c * (al(i - 1) + al(i) - (qtmp + qtmp)));
    } else {
      const T qtmp = q(i);
      return qtmp + (T(1) + c) * (al(i) - qtmp + c * (al(i) + al(i + 1) - (qtmp + qtmp)));
    }
  } else {
    const int fm = flg(i - 1), f0 = flg(i);
    bool use;
    if (mord == 3)
      use = (c > T(0)) ? ((fm & 1) || (f0 & 2)) : ((fm & 2) || (f0 & 1));
    else if (mord == 4)
      use = ((fm & 1) && (f0 & 1)) || ((fm & 2) || (f0 & 2));
    else
      use = (fm & 1) || (f0 & 1);
    T fx1, flux;
    if (c > T(0)) {
      const T blm = bl(i - 1), brm = br(i - 1);
      fx1 = (T(1) - c) * (brm - c * (blm + brm));
      flux = q(i - 1);
    } else {
      const T bl0 = bl(i), br0 = br(i);
      fx1 = (T(1) + c) * (bl0 + c * (bl0 + br0));
      flux = q(i);
    }
    if (use) flux = flux + fx1;
    return flux;
  }
}

}  // namespace fv3t
