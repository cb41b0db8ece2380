// =====================================================================================================
// TEST INFRASTRUCTURE ONLY -- CPU oracle for the FV3 tracer-transport hot path (horizontal part).
//
// A plain C++ restatement of the reference Fortran algorithm, used as the parity checker by tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline leg.  Nothing under fv3atm_b200/ may include,
// link or call this.  PARITY UNPINNED: the reference (NOAA-EMC/fv3atm, submodule atmos_cubed_sphere)
// cannot be compiled in this image (no Fortran compiler, no FMS/MPI) and ships no golden vectors for this
// path; the oracle is pinned to the Fortran source text line by line and cross-checked against the
// reference's own numpy notebook (docs/examples/tp_core.ipynb) and analytic invariants (see tests/).
//
// Reference files restated here (paths relative to atmos_cubed_sphere/):
//   model/tp_core.F90:62-98     constants
//   model/tp_core.F90:110-249   fv_tp_2d
//   model/tp_core.F90:253-330   copy_corners
//   model/tp_core.F90:332-704   xppm      } one 1-D line routine: the two differ only by i<->j
//   model/tp_core.F90:707-1124  yppm      }
//   model/tp_core.F90:1178-1236 pert_ppm
//   model/fv_tracer2d.F90:324-569 tracer_2d (steps A-E), executed for the six tiles of one mosaic in
//                               lock-step; mp_reduce_max -> max over tiles; group halo update -> gather table
//
// Arithmetic: every expression keeps the Fortran operation order; compile with -ffp-contract=off so
// that no FMA is formed (the reference's GNU -O2 x86-64 build forms none either).
// =====================================================================================================
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace fv3oracle {

// Fortran intrinsics with sharp edges
template <class T> static inline T f_sign(T a, T b) { return std::copysign(a, b); }
template <class T> static inline T f_max(T a, T b) { return a > b ? a : b; }
template <class T> static inline T f_min(T a, T b) { return a < b ? a : b; }
template <class T> static inline T f_max(T a, T b, T c) { return f_max(f_max(a, b), c); }
template <class T> static inline T f_min(T a, T b, T c) { return f_min(f_min(a, b), c); }
template <class T> static inline T f_max(T a, T b, T c, T d) { return f_max(f_max(f_max(a, b), c), d); }
template <class T> static inline T f_min(T a, T b, T c, T d) { return f_min(f_min(f_min(a, b), c), d); }
template <class T> static inline T f_abs(T a) { return std::fabs(a); }

// 1-D work array with a Fortran lower bound
template <class T> struct Line {
  std::vector<T> v;
  int lo = 0;
  void reset(int lo_, int hi_) {
    lo = lo_;
    const size_t nn = (size_t)(hi_ - lo_ + 1);
    if (v.size() != nn) v.assign(nn, T(0));
  }
  T& operator()(int i) { return v[(size_t)(i - lo)]; }
};

// strided view of a 1-D line inside a 2-D Fortran array
template <class T> struct SLine {
  T* p;
  int lo;
  long stride;
  T& operator()(int i) const { return p[(long)(i - lo) * stride]; }
};

// 2-D Fortran array view a(lo1:, lo2:)
template <class T> struct V2 {
  T* p;
  int lo1, lo2;
  long ld;
  T& operator()(int i, int j) const { return p[(long)(i - lo1) + (long)(j - lo2) * ld]; }
};

template <class T> struct PpmConst {
  // tp_core.F90:62-98 (default-real literals)
  static constexpr T r3 = T(1) / T(3);
  static constexpr T r12 = T(1) / T(12);
  static constexpr T near_zero = T(1.0e-25);
  static constexpr T ppm_fac = T(1.5);
  static constexpr T s11 = T(11) / T(14), s14 = T(4) / T(7), s15 = T(3) / T(14);
  static constexpr T c1 = T(-2) / T(14), c2 = T(11) / T(14), c3 = T(5) / T(14);
  static constexpr T p1 = T(7) / T(12), p2 = T(-1) / T(12);
};

// pert_ppm (tp_core.F90:1178-1236) for one element
template <class T> static inline void pert_ppm1(T a0, T& al, T& ar, int iv) {
  const T r12 = PpmConst<T>::r12;
  if (iv == 0) {
    if (a0 <= T(0)) {
      al = T(0);
      ar = T(0);
    } else {
      T a4 = T(-3) * (ar + al);
      T da1 = ar - al;
      if (f_abs(da1) < -a4) {
        T fmin = a0 + T(0.25) / a4 * (da1 * da1) + a4 * r12;
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
      T da1 = al - ar;
      T da2 = da1 * da1;
      T a6da = T(3) * (al + ar) * da1;
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

// Thread-private scratch for one PPM line
template <class T> struct PpmScratch {
  Line<T> bl, br, b0, a4, da1, al, dm, dq, fx1;
  std::vector<char> smt5, smt6, ext5, ext6;
};

// One line of xppm / yppm.  q1 on isd:ied, c and flux on is:ie+1, dxa on isd:ied (the metric along the
// sweep direction: dxa(:,j) for xppm, dya(i,:) for yppm).  `edges` = (.not.bounded_domain .and. grid_type<3).
template <class T>
static void ppm_line(SLine<T> flux, SLine<const T> q1, SLine<const T> c, SLine<const T> dxa, int iord, int is, int ie,
                     int npx, bool edges, T lim_fac, PpmScratch<T>& w) {
  using K = PpmConst<T>;
  const T r3 = K::r3, r12 = K::r12;
  int is1, ie3, ie1;
  if (edges) {
    is1 = std::max(3, is - 1);
    ie3 = std::min(npx - 2, ie + 2);
    ie1 = std::min(npx - 3, ie + 1);
  } else {
    is1 = is - 1;
    ie3 = ie + 2;
    ie1 = ie + 1;
  }
  const int mord = std::abs(iord);
  Line<T>&bl = w.bl, &br = w.br, &b0 = w.b0, &a4 = w.a4, &da1 = w.da1, &al = w.al, &dm = w.dm, &dq = w.dq, &fx1 = w.fx1;
  bl.reset(is - 1, ie + 1);
  br.reset(is - 1, ie + 1);
  b0.reset(is - 1, ie + 1);
  a4.reset(is - 1, ie + 1);
  da1.reset(is - 1, ie + 1);
  al.reset(is - 1, ie + 2);
  dm.reset(is - 2, ie + 2);
  dq.reset(is - 3, ie + 2);
  fx1.reset(is, ie + 1);
  const int nl = ie - is + 3;
  if ((int)w.smt5.size() != nl) {
    w.smt5.assign(nl, 0);
    w.smt6.assign(nl, 0);
    w.ext5.assign(nl, 0);
    w.ext6.assign(nl, 0);
  }
  auto smt5 = [&](int i) -> char& { return w.smt5[i - (is - 1)]; };
  auto smt6 = [&](int i) -> char& { return w.smt6[i - (is - 1)]; };
  auto ext5 = [&](int i) -> char& { return w.ext5[i - (is - 1)]; };
  auto ext6 = [&](int i) -> char& { return w.ext6[i - (is - 1)]; };

  if (iord < 7) {
    for (int i = is1; i <= ie3; ++i) al(i) = K::p1 * (q1(i - 1) + q1(i)) + K::p2 * (q1(i - 2) + q1(i + 1));
    if (edges) {
      if (is == 1) {
        al(0) = K::c1 * q1(-2) + K::c2 * q1(-1) + K::c3 * q1(0);
        al(1) = T(0.5) * (((T(2) * dxa(0) + dxa(-1)) * q1(0) - dxa(0) * q1(-1)) / (dxa(-1) + dxa(0)) +
                          ((T(2) * dxa(1) + dxa(2)) * q1(1) - dxa(1) * q1(2)) / (dxa(1) + dxa(2)));
        al(2) = K::c3 * q1(1) + K::c2 * q1(2) + K::c1 * q1(3);
      }
      if ((ie + 1) == npx) {
        al(npx - 1) = K::c1 * q1(npx - 3) + K::c2 * q1(npx - 2) + K::c3 * q1(npx - 1);
        al(npx) = T(0.5) * (((T(2) * dxa(npx - 1) + dxa(npx - 2)) * q1(npx - 1) - dxa(npx - 1) * q1(npx - 2)) /
                                (dxa(npx - 2) + dxa(npx - 1)) +
                            ((T(2) * dxa(npx) + dxa(npx + 1)) * q1(npx) - dxa(npx) * q1(npx + 1)) /
                                (dxa(npx) + dxa(npx + 1)));
        al(npx + 1) = K::c3 * q1(npx) + K::c2 * q1(npx + 1) + K::c1 * q1(npx + 2);
      }
    }
    if (iord < 0) {
      for (int i = is - 1; i <= ie + 2; ++i) al(i) = f_max(T(0), al(i));
    }

    if (mord == 1) {
      for (int i = is - 1; i <= ie + 1; ++i) {
        bl(i) = al(i) - q1(i);
        br(i) = al(i + 1) - q1(i);
        b0(i) = bl(i) + br(i);
        smt5(i) = f_abs(lim_fac * b0(i)) < f_abs(bl(i) - br(i));
      }
      for (int i = is; i <= ie + 1; ++i) {
        if (c(i) > T(0)) {
          fx1(i) = (T(1) - c(i)) * (br(i - 1) - c(i) * b0(i - 1));
          flux(i) = q1(i - 1);
        } else {
          fx1(i) = (T(1) + c(i)) * (bl(i) + c(i) * b0(i));
          flux(i) = q1(i);
        }
        if (smt5(i - 1) || smt5(i)) flux(i) = flux(i) + fx1(i);
      }
    } else if (mord == 2) {
      for (int i = is; i <= ie + 1; ++i) {
        T xt = c(i);
        if (xt > T(0)) {
          T qtmp = q1(i - 1);
          flux(i) = qtmp + (T(1) - xt) * (al(i) - qtmp - xt * (al(i - 1) + al(i) - (qtmp + qtmp)));
        } else {
          T qtmp = q1(i);
          flux(i) = qtmp + (T(1) + xt) * (al(i) - qtmp + xt * (al(i) + al(i + 1) - (qtmp + qtmp)));
        }
      }
    } else if (mord == 3) {
      for (int i = is - 1; i <= ie + 1; ++i) {
        bl(i) = al(i) - q1(i);
        br(i) = al(i + 1) - q1(i);
        b0(i) = bl(i) + br(i);
        T x0 = f_abs(b0(i));
        T xt = f_abs(bl(i) - br(i));
        smt5(i) = x0 < xt;
        smt6(i) = T(3) * x0 < xt;
      }
      for (int i = is; i <= ie + 1; ++i) {
        T xt1 = c(i);
        if (xt1 > T(0)) {
          if (smt5(i - 1) || smt6(i)) {
            flux(i) = q1(i - 1) + (T(1) - xt1) * (br(i - 1) - xt1 * b0(i - 1));
          } else {
            flux(i) = q1(i - 1);
          }
        } else {
          if (smt6(i - 1) || smt5(i)) {
            flux(i) = q1(i) + (T(1) + xt1) * (bl(i) + xt1 * b0(i));
          } else {
            flux(i) = q1(i);
          }
        }
      }
    } else if (mord == 4) {
      for (int i = is - 1; i <= ie + 1; ++i) {
        bl(i) = al(i) - q1(i);
        br(i) = al(i + 1) - q1(i);
        b0(i) = bl(i) + br(i);
        T x0 = f_abs(b0(i));
        T xt = f_abs(bl(i) - br(i));
        smt5(i) = x0 < xt;
        smt6(i) = T(3) * x0 < xt;
      }
      for (int i = is; i <= ie + 1; ++i) {
        T xt1 = c(i);
        bool hi5 = smt5(i - 1) && smt5(i);
        bool hi6 = smt6(i - 1) || smt6(i);
        hi5 = hi5 || hi6;
        if (xt1 > T(0)) {
          fx1(i) = (T(1) - xt1) * (br(i - 1) - xt1 * b0(i - 1));
          flux(i) = q1(i - 1);
        } else {
          fx1(i) = (T(1) + xt1) * (bl(i) + xt1 * b0(i));
          flux(i) = q1(i);
        }
        if (hi5) flux(i) = flux(i) + fx1(i);
      }
    } else {
      if (iord == 5) {
        for (int i = is - 1; i <= ie + 1; ++i) {
          bl(i) = al(i) - q1(i);
          br(i) = al(i + 1) - q1(i);
          b0(i) = bl(i) + br(i);
          smt5(i) = bl(i) * br(i) < T(0);
        }
      } else if (iord == -5) {
        for (int i = is - 1; i <= ie + 1; ++i) {
          bl(i) = al(i) - q1(i);
          br(i) = al(i + 1) - q1(i);
          b0(i) = bl(i) + br(i);
          smt5(i) = bl(i) * br(i) < T(0);
          da1(i) = br(i) - bl(i);
          a4(i) = T(-3) * b0(i);
        }
        for (int i = is - 1; i <= ie + 1; ++i) {
          if (f_abs(da1(i)) < -a4(i)) {
            if (q1(i) + T(0.25) * (da1(i) * da1(i)) / a4(i) + a4(i) * r12 < T(0)) {
              if (!smt5(i)) {
                br(i) = T(0);
                bl(i) = T(0);
                b0(i) = T(0);
              } else if (da1(i) > T(0)) {
                br(i) = T(-2) * bl(i);
                b0(i) = -bl(i);
              } else {
                bl(i) = T(-2) * br(i);
                b0(i) = -br(i);
              }
            }
          }
        }
      } else {
        for (int i = is - 1; i <= ie + 1; ++i) {
          bl(i) = al(i) - q1(i);
          br(i) = al(i + 1) - q1(i);
          b0(i) = bl(i) + br(i);
          smt5(i) = f_abs(T(3) * b0(i)) < f_abs(bl(i) - br(i));
        }
      }
      for (int i = is; i <= ie + 1; ++i) {
        if (c(i) > T(0)) {
          fx1(i) = (T(1) - c(i)) * (br(i - 1) - c(i) * b0(i - 1));
          flux(i) = q1(i - 1);
        } else {
          fx1(i) = (T(1) + c(i)) * (bl(i) + c(i) * b0(i));
          flux(i) = q1(i);
        }
        if (smt5(i - 1) || smt5(i)) flux(i) = flux(i) + fx1(i);
      }
    }
    return;
  }

  // ---- iord >= 7: monotone / positive-definite family (tp_core.F90:556-701) ----
  for (int i = is - 2; i <= ie + 2; ++i) {
    T xt = T(0.25) * (q1(i + 1) - q1(i - 1));
    dm(i) = f_sign(f_min(f_abs(xt), f_max(q1(i - 1), q1(i), q1(i + 1)) - q1(i), q1(i) - f_min(q1(i - 1), q1(i), q1(i + 1))), xt);
  }
  for (int i = is1; i <= ie1 + 1; ++i) al(i) = T(0.5) * (q1(i - 1) + q1(i)) + r3 * (dm(i - 1) - dm(i));

  if (iord == 8) {
    for (int i = is1; i <= ie1; ++i) {
      T xt = T(2) * dm(i);
      bl(i) = -f_sign(f_min(f_abs(xt), f_abs(al(i) - q1(i))), xt);
      br(i) = f_sign(f_min(f_abs(xt), f_abs(al(i + 1) - q1(i))), xt);
    }
  } else if (iord == 10) {
    for (int i = is1 - 2; i <= ie1 + 1; ++i) dq(i) = T(2) * (q1(i + 1) - q1(i));
    for (int i = is1; i <= ie1; ++i) {
      bl(i) = al(i) - q1(i);
      br(i) = al(i + 1) - q1(i);
      if (f_abs(dm(i - 1)) + f_abs(dm(i)) + f_abs(dm(i + 1)) < K::near_zero) {
        bl(i) = T(0);
        br(i) = T(0);
      } else if (f_abs(T(3) * (bl(i) + br(i))) > f_abs(bl(i) - br(i))) {
        T pmp_2 = dq(i - 1);
        T lac_2 = pmp_2 - T(0.75) * dq(i - 2);
        br(i) = f_min(f_max(T(0), pmp_2, lac_2), f_max(br(i), f_min(T(0), pmp_2, lac_2)));
        T pmp_1 = -dq(i);
        T lac_1 = pmp_1 + T(0.75) * dq(i + 1);
        bl(i) = f_min(f_max(T(0), pmp_1, lac_1), f_max(bl(i), f_min(T(0), pmp_1, lac_1)));
      }
    }
  } else if (iord == 11) {
    for (int i = is1; i <= ie1; ++i) {
      T xt = K::ppm_fac * dm(i);
      bl(i) = -f_sign(f_min(f_abs(xt), f_abs(al(i) - q1(i))), xt);
      br(i) = f_sign(f_min(f_abs(xt), f_abs(al(i + 1) - q1(i))), xt);
    }
  } else if (iord == 7 || iord == 12) {
    for (int i = is1; i <= ie1; ++i) {
      bl(i) = al(i) - q1(i);
      br(i) = al(i + 1) - q1(i);
      a4(i) = T(-3) * (bl(i) + br(i));
      da1(i) = br(i) - bl(i);
      ext5(i) = br(i) * bl(i) > T(0);
      ext6(i) = f_abs(da1(i)) < -a4(i);
    }
    for (int i = is1; i <= ie1; ++i) {
      if (ext6(i)) {
        if (q1(i) + T(0.25) / a4(i) * (da1(i) * da1(i)) + a4(i) * r12 < T(0)) {
          if (ext5(i)) {
            br(i) = T(0);
            bl(i) = T(0);
          } else if (da1(i) > T(0)) {
            br(i) = T(-2) * bl(i);
          } else {
            bl(i) = T(-2) * br(i);
          }
        }
      }
    }
  } else {
    for (int i = is1; i <= ie1; ++i) {
      bl(i) = al(i) - q1(i);
      br(i) = al(i + 1) - q1(i);
    }
  }
  if (iord == 9 || iord == 13) {
    for (int i = is1; i <= ie1; ++i) pert_ppm1(q1(i), bl(i), br(i), 0);
  }

  if (edges) {
    if (is == 1) {
      bl(0) = K::s14 * dm(-1) + K::s11 * (q1(-1) - q1(0));
      T xt = T(0.5) * (((T(2) * dxa(0) + dxa(-1)) * q1(0) - dxa(0) * q1(-1)) / (dxa(-1) + dxa(0)) +
                       ((T(2) * dxa(1) + dxa(2)) * q1(1) - dxa(1) * q1(2)) / (dxa(1) + dxa(2)));
      xt = f_max(xt, f_min(q1(-1), q1(0), q1(1), q1(2)));
      xt = f_min(xt, f_max(q1(-1), q1(0), q1(1), q1(2)));
      br(0) = xt - q1(0);
      bl(1) = xt - q1(1);
      xt = K::s15 * q1(1) + K::s11 * q1(2) - K::s14 * dm(2);
      br(1) = xt - q1(1);
      bl(2) = xt - q1(2);
      br(2) = al(3) - q1(2);
      for (int i = 0; i <= 2; ++i) pert_ppm1(q1(i), bl(i), br(i), 1);
    }
    if ((ie + 1) == npx) {
      bl(npx - 2) = al(npx - 2) - q1(npx - 2);
      T xt = K::s15 * q1(npx - 1) + K::s11 * q1(npx - 2) + K::s14 * dm(npx - 2);
      br(npx - 2) = xt - q1(npx - 2);
      bl(npx - 1) = xt - q1(npx - 1);
      xt = T(0.5) * (((T(2) * dxa(npx - 1) + dxa(npx - 2)) * q1(npx - 1) - dxa(npx - 1) * q1(npx - 2)) /
                         (dxa(npx - 2) + dxa(npx - 1)) +
                     ((T(2) * dxa(npx) + dxa(npx + 1)) * q1(npx) - dxa(npx) * q1(npx + 1)) / (dxa(npx) + dxa(npx + 1)));
      xt = f_max(xt, f_min(q1(npx - 2), q1(npx - 1), q1(npx), q1(npx + 1)));
      xt = f_min(xt, f_max(q1(npx - 2), q1(npx - 1), q1(npx), q1(npx + 1)));
      br(npx - 1) = xt - q1(npx - 1);
      bl(npx) = xt - q1(npx);
      br(npx) = K::s11 * (q1(npx + 1) - q1(npx)) - K::s14 * dm(npx + 1);
      for (int i = npx - 2; i <= npx; ++i) pert_ppm1(q1(i), bl(i), br(i), 1);
    }
  }

  if (iord == 7) {
    for (int i = is - 1; i <= ie + 1; ++i) {
      b0(i) = bl(i) + br(i);
      smt5(i) = bl(i) * br(i) < T(0);
    }
    for (int i = is; i <= ie + 1; ++i) {
      if (c(i) > T(0)) {
        fx1(i) = (T(1) - c(i)) * (br(i - 1) - c(i) * b0(i - 1));
        flux(i) = q1(i - 1);
      } else {
        fx1(i) = (T(1) + c(i)) * (bl(i) + c(i) * b0(i));
        flux(i) = q1(i);
      }
      if (smt5(i - 1) || smt5(i)) flux(i) = flux(i) + fx1(i);
    }
  } else {
    for (int i = is; i <= ie + 1; ++i) {
      if (c(i) > T(0)) {
        flux(i) = q1(i - 1) + (T(1) - c(i)) * (br(i - 1) - c(i) * (bl(i - 1) + br(i - 1)));
      } else {
        flux(i) = q1(i) + (T(1) + c(i)) * (bl(i) + c(i) * (bl(i) + br(i)));
      }
    }
  }
}

// Bounds of one tile owned by one rank (fv_grid_bounds_type, fv_arrays.F90:1178-1186), global domain.
struct Bounds {
  int is, ie, js, je, isd, ied, jsd, jed, ng, npx, npy;
  static Bounds tile(int n, int ng = 3) { return Bounds{1, n, 1, n, 1 - ng, n + ng, 1 - ng, n + ng, ng, n + 1, n + 1}; }
};

// The fields of fv_grid_type read on the path, one tile, `real` precision.
template <class T> struct GridT {
  const T *area, *rarea, *dx, *dy, *dxa, *dya, *sin_sg;  // sin_sg: (isd:ied, jsd:jed, 5)
  const T *del6_u = nullptr, *del6_v = nullptr;          // (isd:ied, jsd:jed+1) / (isd:ied+1, jsd:jed), fv_arrays.F90:124; deln_flux only
  T da_min = T(0);                                       // fv_arrays.F90:183
  bool bounded_domain = false;
  int grid_type = 0;
  bool sw_corner = true, se_corner = true, nw_corner = true, ne_corner = true;
};

// copy_corners (tp_core.F90:253-330)
template <class T> static void copy_corners(V2<T> q, int npx, int npy, int dir, const Bounds& bd, const GridT<T>& g) {
  const int ng = bd.ng;
  if (g.bounded_domain) return;
  if (dir == 1) {
    if (g.sw_corner)
      for (int j = 1 - ng; j <= 0; ++j)
        for (int i = 1 - ng; i <= 0; ++i) q(i, j) = q(j, 1 - i);
    if (g.se_corner)
      for (int j = 1 - ng; j <= 0; ++j)
        for (int i = npx; i <= npx + ng - 1; ++i) q(i, j) = q(npy - j, i - npx + 1);
    if (g.ne_corner)
      for (int j = npy; j <= npy + ng - 1; ++j)
        for (int i = npx; i <= npx + ng - 1; ++i) q(i, j) = q(j, 2 * npx - 1 - i);
    if (g.nw_corner)
      for (int j = npy; j <= npy + ng - 1; ++j)
        for (int i = 1 - ng; i <= 0; ++i) q(i, j) = q(npy - j, i - 1 + npx);
  } else if (dir == 2) {
    if (g.sw_corner)
      for (int j = 1 - ng; j <= 0; ++j)
        for (int i = 1 - ng; i <= 0; ++i) q(i, j) = q(1 - j, i);
    if (g.se_corner)
      for (int j = 1 - ng; j <= 0; ++j)
        for (int i = npx; i <= npx + ng - 1; ++i) q(i, j) = q(npy + j - 1, npx - i);
    if (g.ne_corner)
      for (int j = npy; j <= npy + ng - 1; ++j)
        for (int i = npx; i <= npx + ng - 1; ++i) q(i, j) = q(2 * npy - 1 - j, i);
    if (g.nw_corner)
      for (int j = npy; j <= npy + ng - 1; ++j)
        for (int i = 1 - ng; i <= 0; ++i) q(i, j) = q(j + 1 - npx, npy - i);
  }
}

// xppm: rows jfirst:jlast; q(isd:ied, jfirst:jlast), c & flux (is:ie+1, jfirst:jlast), dxa(isd:ied, jsd:jed)
template <class T>
static void xppm(V2<T> flux, V2<const T> q, V2<const T> c, int iord, const Bounds& bd, int jfirst, int jlast,
                 V2<const T> dxa, const GridT<T>& g, T lim_fac, PpmScratch<T>& w) {
  const bool edges = !g.bounded_domain && g.grid_type < 3;
  for (int j = jfirst; j <= jlast; ++j) {
    ppm_line<T>(SLine<T>{&flux(bd.is, j), bd.is, 1}, SLine<const T>{&q(bd.isd, j), bd.isd, 1},
                SLine<const T>{&c(bd.is, j), bd.is, 1}, SLine<const T>{&dxa(bd.isd, j), bd.isd, 1}, iord, bd.is, bd.ie,
                bd.npx, edges, lim_fac, w);
  }
}

// yppm: columns ifirst:ilast; q(ifirst:ilast, jsd:jed), c(isd:ied, js:je+1), flux(ifirst:ilast, js:je+1)
template <class T>
static void yppm(V2<T> flux, V2<const T> q, V2<const T> c, int jord, int ifirst, int ilast, const Bounds& bd,
                 V2<const T> dya, const GridT<T>& g, T lim_fac, PpmScratch<T>& w) {
  const bool edges = !g.bounded_domain && g.grid_type < 3;
  for (int i = ifirst; i <= ilast; ++i) {
    ppm_line<T>(SLine<T>{&flux(i, bd.js), bd.js, flux.ld}, SLine<const T>{&q(i, bd.jsd), bd.jsd, q.ld},
                SLine<const T>{&c(i, bd.js), bd.js, c.ld}, SLine<const T>{&dya(i, bd.jsd), bd.jsd, dya.ld}, jord, bd.js,
                bd.je, bd.npy, edges, lim_fac, w);
  }
}

// Per-thread 2-D temporaries of fv_tp_2d / tracer_2d
template <class T> struct Tp2dScratch {
  std::vector<T> q_i, q_j, fx2, fy2, fyy, fx1, fx, fy, ra_x, ra_y, dp2;
  PpmScratch<T> ppm;
  void size(const Bounds& b) {
    const size_t nxd = b.ied - b.isd + 1, nyd = b.jed - b.jsd + 1, nx = b.ie - b.is + 1, ny = b.je - b.js + 1;
    q_i.resize(nxd * ny);
    q_j.resize(nx * nyd);
    fx2.resize((nx + 1) * nyd);
    fy2.resize(nxd * (ny + 1));
    fyy.resize(nxd * (ny + 1));
    fx1.resize(nx + 1);
    fx.resize((nx + 1) * ny);
    fy.resize(nx * (ny + 1));
    ra_x.resize(nx * nyd);
    ra_y.resize(nxd * ny);
    dp2.resize(nx * ny);
  }
};

// deln_flux (tp_core.F90:1239-1387), both forms (`mass` present / absent), USE_SG undefined (CMake default): del-(2 nord + 2) damping
// fluxes of the cell means (mass weighted when `mass` is given), added to fx, fy.  q is ghosted on input and carries the dir = 1
// corner view that fv_tp_2d left in it (tp_core.F90:189).
template <class T>
static void deln_flux(int nord, int npx, int npy, T damp, V2<const T> q, V2<T> fx, V2<T> fy, const GridT<T>& g, const Bounds& bd,
                      const T* mass_p) {
  const int is = bd.is, ie = bd.ie, js = bd.js, je = bd.je, isd = bd.isd, ied = bd.ied, jsd = bd.jsd, jed = bd.jed;
  const long nxd = ied - isd + 1, nyd = jed - jsd + 1;
  std::vector<T> fx2v((size_t)(nxd + 1) * nyd, T(0)), fy2v((size_t)nxd * (nyd + 1), T(0)), d2v((size_t)nxd * nyd, T(0));
  V2<T> fx2{fx2v.data(), isd, jsd, nxd + 1}, fy2{fy2v.data(), isd, jsd, nxd}, d2{d2v.data(), isd, jsd, nxd};
  V2<const T> del6_v{g.del6_v, isd, jsd, nxd + 1}, del6_u{g.del6_u, isd, jsd, nxd}, rarea{g.rarea, isd, jsd, nxd};
  V2<const T> mass{mass_p, isd, jsd, nxd};
  const int i1 = is - 1 - nord, i2 = ie + 1 + nord, j1 = js - 1 - nord, j2 = je + 1 + nord;
  if (!mass_p) {  // tp_core.F90:1275-1281
    for (int j = j1; j <= j2; ++j)
      for (int i = i1; i <= i2; ++i) d2(i, j) = damp * q(i, j);
  } else {  // `mass` present: damp is applied with the mass weighting below
    for (int j = j1; j <= j2; ++j)
      for (int i = i1; i <= i2; ++i) d2(i, j) = q(i, j);
  }
  if (nord > 0) copy_corners(d2, npx, npy, 1, bd, g);
  for (int j = js - nord; j <= je + nord; ++j)
    for (int i = is - nord; i <= ie + nord + 1; ++i) fx2(i, j) = del6_v(i, j) * (d2(i - 1, j) - d2(i, j));
  if (nord > 0) copy_corners(d2, npx, npy, 2, bd, g);
  for (int j = js - nord; j <= je + nord + 1; ++j)
    for (int i = is - nord; i <= ie + nord; ++i) fy2(i, j) = del6_u(i, j) * (d2(i, j - 1) - d2(i, j));
  for (int n = 1; n <= nord; ++n) {
    const int nt = nord - n;
    for (int j = js - nt - 1; j <= je + nt + 1; ++j)
      for (int i = is - nt - 1; i <= ie + nt + 1; ++i)
        d2(i, j) = (fx2(i, j) - fx2(i + 1, j) + fy2(i, j) - fy2(i, j + 1)) * rarea(i, j);
    copy_corners(d2, npx, npy, 1, bd, g);
    for (int j = js - nt; j <= je + nt; ++j)
      for (int i = is - nt; i <= ie + nt + 1; ++i) fx2(i, j) = del6_v(i, j) * (d2(i, j) - d2(i - 1, j));
    copy_corners(d2, npx, npy, 2, bd, g);
    for (int j = js - nt; j <= je + nt + 1; ++j)
      for (int i = is - nt; i <= ie + nt; ++i) fy2(i, j) = del6_u(i, j) * (d2(i, j) - d2(i, j - 1));
  }
  if (!mass_p) {  // tp_core.F90:1372-1383
    for (int j = js; j <= je; ++j)
      for (int i = is; i <= ie + 1; ++i) fx(i, j) = fx(i, j) + fx2(i, j);
    for (int j = js; j <= je + 1; ++j)
      for (int i = is; i <= ie; ++i) fy(i, j) = fy(i, j) + fy2(i, j);
    return;
  }
  const T damp2 = T(0.5) * damp;
  for (int j = js; j <= je; ++j)
    for (int i = is; i <= ie + 1; ++i) fx(i, j) = fx(i, j) + damp2 * (mass(i - 1, j) + mass(i, j)) * fx2(i, j);
  for (int j = js; j <= je + 1; ++j)
    for (int i = is; i <= ie; ++i) fy(i, j) = fy(i, j) + damp2 * (mass(i, j - 1) + mass(i, j)) * fy2(i, j);
}

// fv_tp_2d (tp_core.F90:110-249).  mfx/mfy present -> tracer branch; nullptr -> xfx/yfx branch (delp, vorticity).
// The Fortran optionals: mass_p == nullptr / nord < 0 / damp_c <= 1e-4 stand for "absent".
template <class T>
static void fv_tp_2d(V2<T> q, V2<const T> crx, V2<const T> cry, int npx, int npy, int hord, V2<T> fx, V2<T> fy,
                     V2<const T> xfx, V2<const T> yfx, const GridT<T>& g, const Bounds& bd, V2<const T> ra_x,
                     V2<const T> ra_y, T lim_fac, const T* mfx_p, const T* mfy_p, Tp2dScratch<T>& w, const T* mass_p = nullptr,
                     int nord = 0, T damp_c = T(0)) {
  const int is = bd.is, ie = bd.ie, js = bd.js, je = bd.je, isd = bd.isd, ied = bd.ied, jsd = bd.jsd, jed = bd.jed;
  const long nxd = ied - isd + 1, nx = ie - is + 1;
  V2<T> q_i{w.q_i.data(), isd, js, nxd};
  V2<T> q_j{w.q_j.data(), is, jsd, nx};
  V2<T> fx2{w.fx2.data(), is, jsd, nx + 1};
  V2<T> fy2{w.fy2.data(), isd, js, nxd};
  V2<T> fyy{w.fyy.data(), isd, js, nxd};
  V2<const T> area{g.area, isd, jsd, nxd};
  V2<const T> dxa{g.dxa, isd, jsd, nxd};
  V2<const T> dya{g.dya, isd, jsd, nxd};
  auto cv = [](V2<T> v) { return V2<const T>{v.p, v.lo1, v.lo2, v.ld}; };

  const int ord_in = (hord == 10) ? 8 : hord;
  const int ord_ou = hord;

  if (!g.bounded_domain) copy_corners(q, npx, npy, 2, bd, g);
  yppm<T>(fy2, cv(q), cry, ord_in, isd, ied, bd, dya, g, lim_fac, w.ppm);
  for (int j = js; j <= je + 1; ++j)
    for (int i = isd; i <= ied; ++i) fyy(i, j) = yfx(i, j) * fy2(i, j);
  for (int j = js; j <= je; ++j)
    for (int i = isd; i <= ied; ++i) q_i(i, j) = (q(i, j) * area(i, j) + fyy(i, j) - fyy(i, j + 1)) / ra_y(i, j);

  xppm<T>(fx, cv(q_i), crx, ord_ou, bd, js, je, dxa, g, lim_fac, w.ppm);

  if (!g.bounded_domain) copy_corners(q, npx, npy, 1, bd, g);
  xppm<T>(fx2, cv(q), crx, ord_in, bd, jsd, jed, dxa, g, lim_fac, w.ppm);
  T* fx1 = w.fx1.data() - is;
  for (int j = jsd; j <= jed; ++j) {
    for (int i = is; i <= ie + 1; ++i) fx1[i] = xfx(i, j) * fx2(i, j);
    for (int i = is; i <= ie; ++i) q_j(i, j) = (q(i, j) * area(i, j) + fx1[i] - fx1[i + 1]) / ra_x(i, j);
  }
  yppm<T>(fy, cv(q_j), cry, ord_ou, is, ie, bd, dya, g, lim_fac, w.ppm);

  if (mfx_p && mfy_p) {
    V2<const T> mfx{mfx_p, is, js, nx + 1};
    V2<const T> mfy{mfy_p, is, js, nx};
    for (int j = js; j <= je; ++j)
      for (int i = is; i <= ie + 1; ++i) fx(i, j) = T(0.5) * (fx(i, j) + fx2(i, j)) * mfx(i, j);
    for (int j = js; j <= je + 1; ++j)
      for (int i = is; i <= ie; ++i) fy(i, j) = T(0.5) * (fy(i, j) + fy2(i, j)) * mfy(i, j);
    if (mass_p && damp_c > T(1.e-4)) {  // tp_core.F90:229-234
      const T damp = std::pow(damp_c * g.da_min, (T)(nord + 1));
      deln_flux<T>(nord, npx, npy, damp, cv(q), fx, fy, g, bd, mass_p);
    }
  } else {
    for (int j = js; j <= je; ++j)
      for (int i = is; i <= ie + 1; ++i) fx(i, j) = T(0.5) * (fx(i, j) + fx2(i, j)) * xfx(i, j);
    for (int j = js; j <= je + 1; ++j)
      for (int i = is; i <= ie; ++i) fy(i, j) = T(0.5) * (fy(i, j) + fy2(i, j)) * yfx(i, j);
    if (nord >= 0 && damp_c > T(1.e-4)) {  // tp_core.F90:243-248: delp / vorticity, no mass weighting
      const T damp = std::pow(damp_c * g.da_min, (T)(nord + 1));
      deln_flux<T>(nord, npx, npy, damp, cv(q), fx, fy, g, bd, nullptr);
    }
  }
}

// -----------------------------------------------------------------------------------------------------
// tracer_2d (fv_tracer2d.F90:324-569) for a mosaic of `ntiles` tiles, each owned by one "rank".
// Arrays are tile-major: q[t](isd:ied,jsd:jed,npz,nq) etc.  The halo update (complete_group_halo_update,
// :499) is the gather  qplane[dst[m]] = qplane[src[m]]  applied to every (k,iq) plane; dst/src are flat
// offsets into the tile-major stack of (nxd*nyd) planes (built from the contact table by the caller).
// q_split==0 only is bit-meaningful (cmax is unset otherwise); q_split/=0 uses ksplt=nsplt.
// -----------------------------------------------------------------------------------------------------
template <class T> struct Mosaic {
  int ntiles, n, npz, nq;
  T *q, *dp1, *mfx, *mfy, *cx, *cy;  // tile-major
  const GridT<T>* grid;              // [ntiles]
  const int64_t *halo_dst, *halo_src;
  int64_t halo_len;
};

template <class T> static void halo_update(const Mosaic<T>& m) {
  const Bounds bd = Bounds::tile(m.n);
  const int64_t plane = (int64_t)(bd.ied - bd.isd + 1) * (bd.jed - bd.jsd + 1);
  const int64_t tile_stride = plane * m.npz * m.nq;
  const int64_t nplanes = (int64_t)m.npz * m.nq;
#pragma omp parallel for schedule(static)
  for (int64_t pl = 0; pl < nplanes; ++pl) {
    for (int64_t e = 0; e < m.halo_len; ++e) {
      const int64_t d = m.halo_dst[e], s = m.halo_src[e];
      const int64_t dt = d / plane, dr = d % plane, st = s / plane, sr = s % plane;
      m.q[dt * tile_stride + pl * plane + dr] = m.q[st * tile_stride + pl * plane + sr];
    }
  }
}

template <class T>
static void tracer_2d_mosaic(const Mosaic<T>& m, int hord, int q_split, T lim_fac, int* nsplt_out, int* ksplt_out,
                             T* cmax_out, int nord_tr = 0, T trdm = T(0)) {
  const Bounds bd = Bounds::tile(m.n);
  const int is = bd.is, ie = bd.ie, js = bd.js, je = bd.je, isd = bd.isd, ied = bd.ied, jsd = bd.jsd, jed = bd.jed;
  const int npx = bd.npx, npy = bd.npy, npz = m.npz, nq = m.nq;
  const long nxd = ied - isd + 1, nyd = jed - jsd + 1, nx = ie - is + 1, ny = je - js + 1;
  const long sz_q2 = nxd * nyd, sz_cx2 = (nx + 1) * nyd, sz_cy2 = nxd * (ny + 1), sz_mfx2 = (nx + 1) * ny,
             sz_mfy2 = nx * (ny + 1);
  std::vector<T> xfx_all((size_t)m.ntiles * sz_cx2 * npz), yfx_all((size_t)m.ntiles * sz_cy2 * npz);
  std::vector<T> cmax(npz, T(0));
  std::vector<std::vector<T>> cmax_t(m.ntiles, std::vector<T>(npz, T(0)));
  std::vector<int> ksplt(npz, 1);

  // ---- step A (per tile == per rank): xfx, yfx, local cmax  (:387-427)
  for (int t = 0; t < m.ntiles; ++t) {
    const GridT<T>& g = m.grid[t];
    V2<const T> dxa{g.dxa, isd, jsd, nxd}, dya{g.dya, isd, jsd, nxd}, dx{g.dx, isd, jsd, nxd}, dy{g.dy, isd, jsd, nxd + 1};
    auto sin_sg = [&](int i, int j, int p) -> T { return g.sin_sg[(long)(i - isd) + (long)(j - jsd) * nxd + (long)(p - 1) * sz_q2]; };
#pragma omp parallel for schedule(static)
    for (int k = 1; k <= npz; ++k) {
      V2<const T> cx{m.cx + (size_t)t * sz_cx2 * npz + (size_t)(k - 1) * sz_cx2, is, jsd, nx + 1};
      V2<const T> cy{m.cy + (size_t)t * sz_cy2 * npz + (size_t)(k - 1) * sz_cy2, isd, js, nxd};
      V2<T> xfx{xfx_all.data() + (size_t)t * sz_cx2 * npz + (size_t)(k - 1) * sz_cx2, is, jsd, nx + 1};
      V2<T> yfx{yfx_all.data() + (size_t)t * sz_cy2 * npz + (size_t)(k - 1) * sz_cy2, isd, js, nxd};
      for (int j = jsd; j <= jed; ++j)
        for (int i = is; i <= ie + 1; ++i) {
          if (cx(i, j) > T(0))
            xfx(i, j) = cx(i, j) * dxa(i - 1, j) * dy(i, j) * sin_sg(i - 1, j, 3);
          else
            xfx(i, j) = cx(i, j) * dxa(i, j) * dy(i, j) * sin_sg(i, j, 1);
        }
      for (int j = js; j <= je + 1; ++j)
        for (int i = isd; i <= ied; ++i) {
          if (cy(i, j) > T(0))
            yfx(i, j) = cy(i, j) * dya(i, j - 1) * dx(i, j) * sin_sg(i, j - 1, 4);
          else
            yfx(i, j) = cy(i, j) * dya(i, j) * dx(i, j) * sin_sg(i, j, 2);
        }
      if (q_split == 0) {
        T cm = T(0);
        if (k < npz / 6) {
          for (int j = js; j <= je; ++j)
            for (int i = is; i <= ie; ++i) cm = f_max(cm, f_abs(cx(i, j)), f_abs(cy(i, j)));
        } else {
          for (int j = js; j <= je; ++j)
            for (int i = is; i <= ie; ++i)
              cm = f_max(cm, f_max(f_abs(cx(i, j)), f_abs(cy(i, j))) + T(1) - sin_sg(i, j, 5));
        }
        cmax_t[t][k - 1] = cm;
      }
    }
  }
  // ---- step B: mp_reduce_max over ranks, nsplt (:432-445)
  int nsplt;
  if (q_split == 0) {
    for (int k = 0; k < npz; ++k) {
      T c = cmax_t[0][k];
      for (int t = 1; t < m.ntiles; ++t) c = f_max(c, cmax_t[t][k]);
      cmax[k] = c;
    }
    T c_global = cmax[0];
    if (npz != 1)
      for (int k = 1; k < npz; ++k) c_global = f_max(cmax[k], c_global);
    nsplt = (int)(T(1) + c_global);
  } else {
    nsplt = q_split;
  }
  // ---- step C: per-level sub-step count and in-place scaling (:449-486)
  if (nsplt != 1) {
    for (int k = 1; k <= npz; ++k) ksplt[k - 1] = (q_split == 0) ? (int)(T(1) + cmax[k - 1]) : nsplt;
    for (int t = 0; t < m.ntiles; ++t) {
#pragma omp parallel for schedule(static)
      for (int k = 1; k <= npz; ++k) {
        const T frac = T(1) / (T)ksplt[k - 1];
        T* cx = m.cx + (size_t)t * sz_cx2 * npz + (size_t)(k - 1) * sz_cx2;
        T* xf = xfx_all.data() + (size_t)t * sz_cx2 * npz + (size_t)(k - 1) * sz_cx2;
        for (long e = 0; e < sz_cx2; ++e) {
          cx[e] = cx[e] * frac;
          xf[e] = xf[e] * frac;
        }
        T* mfx = m.mfx + (size_t)t * sz_mfx2 * npz + (size_t)(k - 1) * sz_mfx2;
        for (long e = 0; e < sz_mfx2; ++e) mfx[e] = mfx[e] * frac;
        T* cy = m.cy + (size_t)t * sz_cy2 * npz + (size_t)(k - 1) * sz_cy2;
        T* yf = yfx_all.data() + (size_t)t * sz_cy2 * npz + (size_t)(k - 1) * sz_cy2;
        for (long e = 0; e < sz_cy2; ++e) {
          cy[e] = cy[e] * frac;
          yf[e] = yf[e] * frac;
        }
        T* mfy = m.mfy + (size_t)t * sz_mfy2 * npz + (size_t)(k - 1) * sz_mfy2;
        for (long e = 0; e < sz_mfy2; ++e) mfy[e] = mfy[e] * frac;
      }
    }
  }
  if (nsplt_out) *nsplt_out = nsplt;
  if (ksplt_out) std::memcpy(ksplt_out, ksplt.data(), sizeof(int) * npz);
  if (cmax_out) std::memcpy(cmax_out, cmax.data(), sizeof(T) * npz);

  // ---- tracer damping: the edge halo of dp1 (complete_group_halo_update(dp1_pack), :487-494)
  if (trdm > T(1.e-4)) {
    const int64_t plane = sz_q2;
    for (int64_t pl = 0; pl < npz; ++pl)
      for (int64_t e = 0; e < m.halo_len; ++e) {
        const int64_t d = m.halo_dst[e], s = m.halo_src[e];
        m.dp1[(d / plane) * plane * npz + pl * plane + d % plane] = m.dp1[(s / plane) * plane * npz + pl * plane + s % plane];
      }
  }
  // ---- step E: sub-cycled transport (:496-566)
  for (int it = 1; it <= nsplt; ++it) {
    halo_update(m);  // complete_group_halo_update(q_pack)
    for (int t = 0; t < m.ntiles; ++t) {
      const GridT<T>& g = m.grid[t];
      V2<const T> area{g.area, isd, jsd, nxd}, rarea{g.rarea, isd, jsd, nxd};
#pragma omp parallel
      {
        Tp2dScratch<T> w;
        w.size(bd);
#pragma omp for schedule(dynamic, 1)
        for (int k = 1; k <= npz; ++k) {
          if (it <= ksplt[k - 1]) {
            V2<T> dp1{m.dp1 + (size_t)t * sz_q2 * npz + (size_t)(k - 1) * sz_q2, isd, jsd, nxd};
            V2<const T> mfx{m.mfx + (size_t)t * sz_mfx2 * npz + (size_t)(k - 1) * sz_mfx2, is, js, nx + 1};
            V2<const T> mfy{m.mfy + (size_t)t * sz_mfy2 * npz + (size_t)(k - 1) * sz_mfy2, is, js, nx};
            V2<const T> cx{m.cx + (size_t)t * sz_cx2 * npz + (size_t)(k - 1) * sz_cx2, is, jsd, nx + 1};
            V2<const T> cy{m.cy + (size_t)t * sz_cy2 * npz + (size_t)(k - 1) * sz_cy2, isd, js, nxd};
            V2<const T> xfx{xfx_all.data() + (size_t)t * sz_cx2 * npz + (size_t)(k - 1) * sz_cx2, is, jsd, nx + 1};
            V2<const T> yfx{yfx_all.data() + (size_t)t * sz_cy2 * npz + (size_t)(k - 1) * sz_cy2, isd, js, nxd};
            V2<T> dp2{w.dp2.data(), is, js, nx};
            V2<T> ra_x{w.ra_x.data(), is, jsd, nx};
            V2<T> ra_y{w.ra_y.data(), isd, js, nxd};
            V2<T> fx{w.fx.data(), is, js, nx + 1};
            V2<T> fy{w.fy.data(), is, js, nx};
            for (int j = js; j <= je; ++j)
              for (int i = is; i <= ie; ++i)
                dp2(i, j) = dp1(i, j) + (mfx(i, j) - mfx(i + 1, j) + mfy(i, j) - mfy(i, j + 1)) * rarea(i, j);
            for (int j = jsd; j <= jed; ++j)
              for (int i = is; i <= ie; ++i) ra_x(i, j) = area(i, j) + xfx(i, j) - xfx(i + 1, j);
            for (int j = js; j <= je; ++j)
              for (int i = isd; i <= ied; ++i) ra_y(i, j) = area(i, j) + yfx(i, j) - yfx(i, j + 1);
            for (int iq = 1; iq <= nq; ++iq) {
              V2<T> q{m.q + (size_t)t * sz_q2 * npz * nq + ((size_t)(iq - 1) * npz + (k - 1)) * sz_q2, isd, jsd, nxd};
              if (it == 1 && trdm > T(1.e-4))  // :527-532: damping with the first sub-step only
                fv_tp_2d<T>(q, cx, cy, npx, npy, hord, fx, fy, xfx, yfx, g, bd, V2<const T>{ra_x.p, is, jsd, nx},
                            V2<const T>{ra_y.p, isd, js, nxd}, lim_fac, mfx.p, mfy.p, w, dp1.p, nord_tr, trdm);
              else
                fv_tp_2d<T>(q, cx, cy, npx, npy, hord, fx, fy, xfx, yfx, g, bd, V2<const T>{ra_x.p, is, jsd, nx},
                            V2<const T>{ra_y.p, isd, js, nxd}, lim_fac, mfx.p, mfy.p, w);
              for (int j = js; j <= je; ++j)
                for (int i = is; i <= ie; ++i)
                  q(i, j) = (q(i, j) * dp1(i, j) + (fx(i, j) - fx(i + 1, j) + fy(i, j) - fy(i, j + 1)) * rarea(i, j)) / dp2(i, j);
            }
            if (it != nsplt) {
              for (int j = js; j <= je; ++j)
                for (int i = is; i <= ie; ++i) dp1(i, j) = dp2(i, j);
            }
          }
        }
      }
    }
    // (start_group_halo_update for the next sub-step happens at the top of the loop)
  }
}

// -----------------------------------------------------------------------------------------------------
// tracer_2d_1L (fv_tracer2d.F90:92-321): the level-at-a-time variant fv_dynamics calls when z_tracer is set
// (fv_dynamics.F90:696-698).  Same arithmetic as tracer_2d but a different driver: the sub-step count is
// per level (nsplt = int(1 + cmax(k)), :201-202, :254), the level's tracers are staged in qn2 (:263-270), q is
// written only by the last sub-step (:282-288) and qn2's halo is refreshed with a blocking update between the
// sub-steps of that level (:314).  q_split is not used by this routine.  Restated independently of
// tracer_2d_mosaic so that the two can pin each other: for trdm2 = 0 every output must be
// bit-identical.
// -----------------------------------------------------------------------------------------------------
template <class T>
static void tracer_2d_1L_mosaic(const Mosaic<T>& m, int hord, T lim_fac, int* nsplt_max_out, int* ksplt_out, T* cmax_out) {
  const Bounds bd = Bounds::tile(m.n);
  const int is = bd.is, ie = bd.ie, js = bd.js, je = bd.je, isd = bd.isd, ied = bd.ied, jsd = bd.jsd, jed = bd.jed;
  const int npx = bd.npx, npy = bd.npy, npz = m.npz, nq = m.nq;
  const long nxd = ied - isd + 1, nyd = jed - jsd + 1, nx = ie - is + 1, ny = je - js + 1;
  const long sz_q2 = nxd * nyd, sz_cx2 = (nx + 1) * nyd, sz_cy2 = nxd * (ny + 1), sz_mfx2 = (nx + 1) * ny,
             sz_mfy2 = nx * (ny + 1);
  std::vector<T> xfx_all((size_t)m.ntiles * sz_cx2 * npz), yfx_all((size_t)m.ntiles * sz_cy2 * npz);
  std::vector<T> cmax(npz, T(0));

  // xfx, yfx, cmax per level (:179-214), every tile playing one rank; mp_reduce_max (:225) = max over tiles
  for (int k = 1; k <= npz; ++k) {
    T cm_all = T(0);
    for (int t = 0; t < m.ntiles; ++t) {
      const GridT<T>& g = m.grid[t];
      V2<const T> dxa{g.dxa, isd, jsd, nxd}, dya{g.dya, isd, jsd, nxd}, dx{g.dx, isd, jsd, nxd}, dy{g.dy, isd, jsd, nxd + 1};
      auto sin_sg = [&](int i, int j, int p) -> T { return g.sin_sg[(long)(i - isd) + (long)(j - jsd) * nxd + (long)(p - 1) * sz_q2]; };
      V2<const T> cx{m.cx + (size_t)t * sz_cx2 * npz + (size_t)(k - 1) * sz_cx2, is, jsd, nx + 1};
      V2<const T> cy{m.cy + (size_t)t * sz_cy2 * npz + (size_t)(k - 1) * sz_cy2, isd, js, nxd};
      V2<T> xfx{xfx_all.data() + (size_t)t * sz_cx2 * npz + (size_t)(k - 1) * sz_cx2, is, jsd, nx + 1};
      V2<T> yfx{yfx_all.data() + (size_t)t * sz_cy2 * npz + (size_t)(k - 1) * sz_cy2, isd, js, nxd};
      for (int j = jsd; j <= jed; ++j)
        for (int i = is; i <= ie + 1; ++i)
          xfx(i, j) = (cx(i, j) > T(0)) ? cx(i, j) * dxa(i - 1, j) * dy(i, j) * sin_sg(i - 1, j, 3)
                                        : cx(i, j) * dxa(i, j) * dy(i, j) * sin_sg(i, j, 1);
      for (int j = js; j <= je + 1; ++j)
        for (int i = isd; i <= ied; ++i)
          yfx(i, j) = (cy(i, j) > T(0)) ? cy(i, j) * dya(i, j - 1) * dx(i, j) * sin_sg(i, j - 1, 4)
                                        : cy(i, j) * dya(i, j) * dx(i, j) * sin_sg(i, j, 2);
      T cm = T(0);
      if (k < npz / 6) {
        for (int j = js; j <= je; ++j)
          for (int i = is; i <= ie; ++i) cm = f_max(cm, f_abs(cx(i, j)), f_abs(cy(i, j)));
      } else {
        for (int j = js; j <= je; ++j)
          for (int i = is; i <= ie; ++i) cm = f_max(cm, f_max(f_abs(cx(i, j)), f_abs(cy(i, j))) + T(1) - sin_sg(i, j, 5));
      }
      cm_all = (t == 0) ? cm : f_max(cm_all, cm);
    }
    cmax[k - 1] = cm_all;
  }
  // per-level sub-step count and in-place scaling (:227-259)
  int nsplt_max = 1;
  for (int k = 1; k <= npz; ++k) {
    const int nsplt = (int)(T(1) + cmax[k - 1]);
    if (ksplt_out) ksplt_out[k - 1] = nsplt;
    nsplt_max = nsplt > nsplt_max ? nsplt : nsplt_max;
    if (nsplt > 1) {
      const T frac = T(1) / (T)nsplt;
      for (int t = 0; t < m.ntiles; ++t) {
        T* cx = m.cx + (size_t)t * sz_cx2 * npz + (size_t)(k - 1) * sz_cx2;
        T* xf = xfx_all.data() + (size_t)t * sz_cx2 * npz + (size_t)(k - 1) * sz_cx2;
        for (long e = 0; e < sz_cx2; ++e) {
          cx[e] = cx[e] * frac;
          xf[e] = xf[e] * frac;
        }
        T* mfx = m.mfx + (size_t)t * sz_mfx2 * npz + (size_t)(k - 1) * sz_mfx2;
        for (long e = 0; e < sz_mfx2; ++e) mfx[e] = mfx[e] * frac;
        T* cy = m.cy + (size_t)t * sz_cy2 * npz + (size_t)(k - 1) * sz_cy2;
        T* yf = yfx_all.data() + (size_t)t * sz_cy2 * npz + (size_t)(k - 1) * sz_cy2;
        for (long e = 0; e < sz_cy2; ++e) {
          cy[e] = cy[e] * frac;
          yf[e] = yf[e] * frac;
        }
        T* mfy = m.mfy + (size_t)t * sz_mfy2 * npz + (size_t)(k - 1) * sz_mfy2;
        for (long e = 0; e < sz_mfy2; ++e) mfy[e] = mfy[e] * frac;
      }
    }
  }
  if (nsplt_max_out) *nsplt_max_out = nsplt_max;
  if (cmax_out) std::memcpy(cmax_out, cmax.data(), sizeof(T) * npz);

  halo_update(m);  // complete_group_halo_update(q_pack) (:261)

  // level loop (:268-319); qn2(isd:ied, jsd:jed, nq) per tile
  std::vector<T> qn2((size_t)m.ntiles * sz_q2 * nq);
  Tp2dScratch<T> w;
  w.size(bd);
  for (int k = 1; k <= npz; ++k) {
    const int nsplt = (int)(T(1) + cmax[k - 1]);
    for (int it = 1; it <= nsplt; ++it) {
      for (int t = 0; t < m.ntiles; ++t) {
        const GridT<T>& g = m.grid[t];
        V2<const T> area{g.area, isd, jsd, nxd}, rarea{g.rarea, isd, jsd, nxd};
        V2<T> dp1{m.dp1 + (size_t)t * sz_q2 * npz + (size_t)(k - 1) * sz_q2, isd, jsd, nxd};
        V2<const T> mfx{m.mfx + (size_t)t * sz_mfx2 * npz + (size_t)(k - 1) * sz_mfx2, is, js, nx + 1};
        V2<const T> mfy{m.mfy + (size_t)t * sz_mfy2 * npz + (size_t)(k - 1) * sz_mfy2, is, js, nx};
        V2<const T> cx{m.cx + (size_t)t * sz_cx2 * npz + (size_t)(k - 1) * sz_cx2, is, jsd, nx + 1};
        V2<const T> cy{m.cy + (size_t)t * sz_cy2 * npz + (size_t)(k - 1) * sz_cy2, isd, js, nxd};
        V2<const T> xfx{xfx_all.data() + (size_t)t * sz_cx2 * npz + (size_t)(k - 1) * sz_cx2, is, jsd, nx + 1};
        V2<const T> yfx{yfx_all.data() + (size_t)t * sz_cy2 * npz + (size_t)(k - 1) * sz_cy2, isd, js, nxd};
        V2<T> dp2{w.dp2.data(), is, js, nx};
        V2<T> ra_x{w.ra_x.data(), is, jsd, nx};
        V2<T> ra_y{w.ra_y.data(), isd, js, nxd};
        V2<T> fx{w.fx.data(), is, js, nx + 1};
        V2<T> fy{w.fy.data(), is, js, nx};
        for (int j = jsd; j <= jed; ++j) {  // :270-279 (hoisted out of the it loop in the reference; same values)
          for (int i = is; i <= ie; ++i) ra_x(i, j) = area(i, j) + xfx(i, j) - xfx(i + 1, j);
          if (j >= js && j <= je)
            for (int i = isd; i <= ied; ++i) ra_y(i, j) = area(i, j) + yfx(i, j) - yfx(i, j + 1);
        }
        for (int j = js; j <= je; ++j)
          for (int i = is; i <= ie; ++i)
            dp2(i, j) = dp1(i, j) + (mfx(i, j) - mfx(i + 1, j) + mfy(i, j) - mfy(i, j + 1)) * rarea(i, j);
        for (int iq = 1; iq <= nq; ++iq) {
          V2<T> q{m.q + (size_t)t * sz_q2 * npz * nq + ((size_t)(iq - 1) * npz + (k - 1)) * sz_q2, isd, jsd, nxd};
          V2<T> qn{qn2.data() + ((size_t)t * nq + (iq - 1)) * sz_q2, isd, jsd, nxd};
          if (nsplt != 1) {
            if (it == 1)
              for (int j = jsd; j <= jed; ++j)
                for (int i = isd; i <= ied; ++i) qn(i, j) = q(i, j);
            fv_tp_2d<T>(qn, cx, cy, npx, npy, hord, fx, fy, xfx, yfx, g, bd, V2<const T>{ra_x.p, is, jsd, nx},
                        V2<const T>{ra_y.p, isd, js, nxd}, lim_fac, mfx.p, mfy.p, w);
            V2<T> dst = (it < nsplt) ? qn : q;
            for (int j = js; j <= je; ++j)
              for (int i = is; i <= ie; ++i)
                dst(i, j) = (qn(i, j) * dp1(i, j) + (fx(i, j) - fx(i + 1, j) + fy(i, j) - fy(i, j + 1)) * rarea(i, j)) / dp2(i, j);
          } else {
            fv_tp_2d<T>(q, cx, cy, npx, npy, hord, fx, fy, xfx, yfx, g, bd, V2<const T>{ra_x.p, is, jsd, nx},
                        V2<const T>{ra_y.p, isd, js, nxd}, lim_fac, mfx.p, mfy.p, w);
            for (int j = js; j <= je; ++j)
              for (int i = is; i <= ie; ++i)
                q(i, j) = (q(i, j) * dp1(i, j) + (fx(i, j) - fx(i + 1, j) + fy(i, j) - fy(i, j + 1)) * rarea(i, j)) / dp2(i, j);
          }
        }
        if (it < nsplt)
          for (int j = js; j <= je; ++j)
            for (int i = is; i <= ie; ++i) dp1(i, j) = dp2(i, j);
      }
      if (it < nsplt) {  // mpp_update_domains(qn2, domain) (:314): edge halos of every tracer slab of this level
        const int64_t plane = sz_q2, tile_stride = plane * nq;
        for (int64_t pl = 0; pl < nq; ++pl)
          for (int64_t e = 0; e < m.halo_len; ++e) {
            const int64_t d = m.halo_dst[e], s = m.halo_src[e];
            qn2[(d / plane) * tile_stride + pl * plane + d % plane] = qn2[(s / plane) * tile_stride + pl * plane + s % plane];
          }
      }
    }
  }
}

}  // namespace fv3oracle
