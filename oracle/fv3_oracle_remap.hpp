// =====================================================================================================
// TEST INFRASTRUCTURE ONLY -- CPU oracle for the FV3 tracer-transport hot path (vertical remap part).
// See fv3_oracle_advect.hpp for the status of this code (parity unpinned; pinned to the Fortran text).
//
// Reference files restated here (paths relative to atmos_cubed_sphere/):
//   model/fv_mapz.F90:110         r3, r23, r12
//   model/fv_mapz.F90:261-273,343-368,407-426  Lagrangian_to_Eulerian, tracer part of the j loop
//   model/fv_mapz.F90:1386-1499   mapn_tracer
//   model/fv_mapz.F90:1502-1592   map1_q2
//   model/fv_mapz.F90:1691-2096   scalar_profile
//   model/fv_mapz.F90:2098-2498   cs_profile
//   model/fv_mapz.F90:2501-2576   cs_limiters
//   model/fv_mapz.F90:2580-2837   ppm_profile
//   model/fv_mapz.F90:2840-2916   ppm_limiters
//   model/fv_fill.F90:86-153      fillz (the #else branch: DEV_GFS_PHYS is not defined by any CMake file)
//
// The Fortran routines are written i-vectorised over a row of columns; every column is independent, so
// they are restated here per column (loop interchange only; the per-column operation order is the
// reference's).  Arrays are 1-based like the Fortran.
// =====================================================================================================
#pragma once
#include "fv3_oracle_advect.hpp"

namespace fv3oracle {

template <class T> struct MapConst {
  static constexpr T r3 = T(1) / T(3), r23 = T(2) / T(3), r12 = T(1) / T(12);
};

// a4(4, km) for one column, 1-based: A(c,k), c = 1..4
template <class T> struct ColA4 {
  std::vector<T> v;
  int km = 0;
  void size(int km_) {
    km = km_;
    v.resize((size_t)4 * (km_ + 2));
  }
  T& operator()(int c, int k) { return v[(size_t)(c - 1) + (size_t)4 * k]; }
};

// cs_limiters (fv_mapz.F90:2501-2576) for one element
template <class T> static inline void cs_limiters1(bool extm, T& a1, T& a2, T& a3, T& a4, int iv) {
  const T r12 = MapConst<T>::r12;
  if (iv == 0) {
    if (a1 <= T(0)) {
      a2 = a1;
      a3 = a1;
      a4 = T(0);
    } else {
      if (f_abs(a3 - a2) < -a4) {
        if ((a1 + T(0.25) * ((a3 - a2) * (a3 - a2)) / a4 + a4 * r12) < T(0)) {
          if (a1 < a3 && a1 < a2) {
            a3 = a1;
            a2 = a1;
            a4 = T(0);
          } else if (a3 > a2) {
            a4 = T(3) * (a2 - a1);
            a3 = a2 - a4;
          } else {
            a4 = T(3) * (a3 - a1);
            a2 = a3 - a4;
          }
        }
      }
    }
  } else if (iv == 1) {
    if ((a1 - a2) * (a1 - a3) >= T(0)) {
      a2 = a1;
      a3 = a1;
      a4 = T(0);
    } else {
      T da1 = a3 - a2;
      T da2 = da1 * da1;
      T a6da = a4 * da1;
      if (a6da < -da2) {
        a4 = T(3) * (a2 - a1);
        a3 = a2 - a4;
      } else if (a6da > da2) {
        a4 = T(3) * (a3 - a1);
        a2 = a3 - a4;
      }
    }
  } else {
    if (extm) {
      a2 = a1;
      a3 = a1;
      a4 = T(0);
    } else {
      T da1 = a3 - a2;
      T da2 = da1 * da1;
      T a6da = a4 * da1;
      if (a6da < -da2) {
        a4 = T(3) * (a2 - a1);
        a3 = a2 - a4;
      } else if (a6da > da2) {
        a4 = T(3) * (a3 - a1);
        a2 = a3 - a4;
      }
    }
  }
}

template <class T> struct ProfScratch {
  std::vector<T> gam, q, dc, h2, delq, df2, d4;
  std::vector<char> extm, ext5, ext6;
  void size(int km) {
    gam.resize(km + 3);
    q.resize(km + 3);
    dc.resize(km + 3);
    h2.resize(km + 3);
    delq.resize(km + 3);
    df2.resize(km + 3);
    d4.resize(km + 3);
    extm.resize(km + 3);
    ext5.resize(km + 3);
    ext6.resize(km + 3);
  }
};

// scalar_profile (fv_mapz.F90:1691-2096) and cs_profile (:2098-2498) for one column.
// `scalar` selects scalar_profile (uses qmin in the |kord| = 9, 11(else), 15 branches and the
// 3*(2*a1-(a2+a3)) form of a4 for |kord| = 9) versus cs_profile (no qmin; 6*a1-3*(a2+a3) for |kord| = 9).
template <class T>
static void cs_profile_col(bool scalar, T qs, ColA4<T>& a4, const T* delp /*1-based*/, int km, int iv, int kord, T qmin,
                           ProfScratch<T>& w) {
  T* gam = w.gam.data();
  T* q = w.q.data();
  char* extm = w.extm.data();
  char* ext5 = w.ext5.data();
  char* ext6 = w.ext6.data();
  T d4 = T(0);
  const int akord = std::abs(kord);

  if (iv == -2) {
    gam[2] = T(0.5);
    q[1] = T(1.5) * a4(1, 1);
    for (int k = 2; k <= km - 1; ++k) {
      T grat = delp[k - 1] / delp[k];
      T bet = T(2) + grat + grat - gam[k];
      q[k] = (T(3) * (a4(1, k - 1) + a4(1, k)) - q[k - 1]) / bet;
      gam[k + 1] = grat / bet;
    }
    T grat = delp[km - 1] / delp[km];
    q[km] = (T(3) * (a4(1, km - 1) + a4(1, km)) - grat * qs - q[km - 1]) / (T(2) + grat + grat - gam[km]);
    q[km + 1] = qs;
    for (int k = km - 1; k >= 1; --k) q[k] = q[k] - gam[k + 1] * q[k + 1];
  } else {
    T grat = delp[2] / delp[1];
    T bet = grat * (grat + T(0.5));
    q[1] = ((grat + grat) * (grat + T(1)) * a4(1, 1) + a4(1, 2)) / bet;
    gam[1] = (T(1) + grat * (grat + T(1.5))) / bet;
    for (int k = 2; k <= km; ++k) {
      d4 = delp[k - 1] / delp[k];
      bet = T(2) + d4 + d4 - gam[k - 1];
      q[k] = (T(3) * (a4(1, k - 1) + d4 * a4(1, k)) - q[k - 1]) / bet;
      gam[k] = d4 / bet;
    }
    T a_bot = T(1) + d4 * (d4 + T(1.5));
    q[km + 1] = (T(2) * d4 * (d4 + T(1)) * a4(1, km) + a4(1, km - 1) - a_bot * q[km]) / (d4 * (d4 + T(0.5)) - a_bot * gam[km]);
    for (int k = km; k >= 1; --k) q[k] = q[k] - gam[k] * q[k + 1];
  }

  if (akord > 16) {
    for (int k = 1; k <= km; ++k) {
      a4(2, k) = q[k];
      a4(3, k) = q[k + 1];
      a4(4, k) = T(3) * (T(2) * a4(1, k) - (a4(2, k) + a4(3, k)));
    }
    return;
  }

  // large-scale constraints
  q[2] = f_min(q[2], f_max(a4(1, 1), a4(1, 2)));
  q[2] = f_max(q[2], f_min(a4(1, 1), a4(1, 2)));
  for (int k = 2; k <= km; ++k) gam[k] = a4(1, k) - a4(1, k - 1);
  for (int k = 3; k <= km - 1; ++k) {
    if (gam[k - 1] * gam[k + 1] > T(0)) {
      q[k] = f_min(q[k], f_max(a4(1, k - 1), a4(1, k)));
      q[k] = f_max(q[k], f_min(a4(1, k - 1), a4(1, k)));
    } else {
      if (gam[k - 1] > T(0)) {
        q[k] = f_max(q[k], f_min(a4(1, k - 1), a4(1, k)));
      } else {
        q[k] = f_min(q[k], f_max(a4(1, k - 1), a4(1, k)));
        if (iv == 0) q[k] = f_max(T(0), q[k]);
      }
    }
  }
  q[km] = f_min(q[km], f_max(a4(1, km - 1), a4(1, km)));
  q[km] = f_max(q[km], f_min(a4(1, km - 1), a4(1, km)));

  for (int k = 1; k <= km; ++k) {
    a4(2, k) = q[k];
    a4(3, k) = q[k + 1];
  }
  for (int k = 1; k <= km; ++k) {
    if (k == 1 || k == km)
      extm[k] = (a4(2, k) - a4(1, k)) * (a4(3, k) - a4(1, k)) > T(0);
    else
      extm[k] = gam[k] * gam[k + 1] < T(0);
    if (akord > 9) {
      T x0 = T(2) * a4(1, k) - (a4(2, k) + a4(3, k));
      T x1 = f_abs(a4(2, k) - a4(3, k));
      a4(4, k) = T(3) * x0;
      ext5[k] = f_abs(x0) > x1;
      ext6[k] = f_abs(a4(4, k)) > x1;
    }
  }

  // top layer
  if (iv == 0) {
    a4(2, 1) = f_max(T(0), a4(2, 1));
  } else if (iv == -1) {
    if (a4(2, 1) * a4(1, 1) <= T(0)) a4(2, 1) = T(0);
  } else if (iv == 2) {
    a4(2, 1) = a4(1, 1);
    a4(3, 1) = a4(1, 1);
    a4(4, 1) = T(0);
  }
  if (iv != 2) {
    a4(4, 1) = T(3) * (T(2) * a4(1, 1) - (a4(2, 1) + a4(3, 1)));
    cs_limiters1<T>(extm[1], a4(1, 1), a4(2, 1), a4(3, 1), a4(4, 1), 1);
  }
  a4(4, 2) = T(3) * (T(2) * a4(1, 2) - (a4(2, 2) + a4(3, 2)));
  cs_limiters1<T>(extm[2], a4(1, 2), a4(2, 2), a4(3, 2), a4(4, 2), 2);

  auto huynh = [&](int k) {
    T pmp_1 = a4(1, k) - T(2) * gam[k + 1];
    T lac_1 = pmp_1 + T(1.5) * gam[k + 2];
    a4(2, k) = f_min(f_max(a4(2, k), f_min(a4(1, k), pmp_1, lac_1)), f_max(a4(1, k), pmp_1, lac_1));
    T pmp_2 = a4(1, k) + T(2) * gam[k];
    T lac_2 = pmp_2 - T(1.5) * gam[k - 1];
    a4(3, k) = f_min(f_max(a4(3, k), f_min(a4(1, k), pmp_2, lac_2)), f_max(a4(1, k), pmp_2, lac_2));
  };
  auto set_a6 = [&](int k) { a4(4, k) = T(3) * (T(2) * a4(1, k) - (a4(2, k) + a4(3, k))); };
  auto flat = [&](int k) {
    a4(2, k) = a4(1, k);
    a4(3, k) = a4(1, k);
  };

  for (int k = 3; k <= km - 2; ++k) {
    if (akord < 9) {
      huynh(k);
      set_a6(k);
    } else if (akord == 9) {
      if (extm[k] && extm[k - 1]) {
        flat(k);
        a4(4, k) = T(0);
      } else if (extm[k] && extm[k + 1]) {
        flat(k);
        a4(4, k) = T(0);
      } else if (scalar && extm[k] && a4(1, k) < qmin) {
        flat(k);
        a4(4, k) = T(0);
      } else {
        if (scalar)
          set_a6(k);
        else
          a4(4, k) = T(6) * a4(1, k) - T(3) * (a4(2, k) + a4(3, k));
        if (f_abs(a4(4, k)) > f_abs(a4(2, k) - a4(3, k))) {
          huynh(k);
          if (scalar)
            set_a6(k);
          else
            a4(4, k) = T(6) * a4(1, k) - T(3) * (a4(2, k) + a4(3, k));
        }
      }
    } else if (akord == 10) {
      if (ext5[k]) {
        if (ext5[k - 1] || ext5[k + 1]) {
          flat(k);
        } else if (ext6[k - 1] || ext6[k + 1]) {
          huynh(k);
        }
      } else if (ext6[k]) {
        if (ext5[k - 1] || ext5[k + 1]) huynh(k);
      }
      set_a6(k);
    } else if (akord == 12) {
      if (extm[k]) {
        flat(k);
        a4(4, k) = T(0);
      } else {
        a4(4, k) = T(6) * a4(1, k) - T(3) * (a4(2, k) + a4(3, k));
        if (f_abs(a4(4, k)) > f_abs(a4(2, k) - a4(3, k))) {
          huynh(k);
          a4(4, k) = T(6) * a4(1, k) - T(3) * (a4(2, k) + a4(3, k));
        }
      }
    } else if (akord == 13) {
      if (ext6[k]) {
        if (ext6[k - 1] && ext6[k + 1]) flat(k);
      }
      set_a6(k);
    } else if (akord == 14) {
      set_a6(k);
    } else if (akord == 15) {
      if (scalar) {
        if (ext5[k] && ext5[k - 1]) {
          flat(k);
        } else if (ext5[k] && ext5[k + 1]) {
          flat(k);
        } else if (ext5[k] && a4(1, k) < qmin) {
          flat(k);
        } else if (ext6[k]) {
          huynh(k);
        }
      } else {
        if (ext5[k]) {
          if (ext5[k - 1] || ext5[k + 1]) flat(k);
        } else if (ext6[k]) {
          huynh(k);
        }
      }
      set_a6(k);
    } else if (akord == 16) {
      if (ext5[k]) {
        if (ext5[k - 1] || ext5[k + 1]) {
          flat(k);
        } else if (ext6[k - 1] || ext6[k + 1]) {
          huynh(k);
        }
      }
      set_a6(k);
    } else {  // |kord| = 11
      const bool noisy = scalar ? (ext5[k] && (ext5[k - 1] || ext5[k + 1] || a4(1, k) < qmin))
                                : (ext5[k] && (ext5[k - 1] || ext5[k + 1]));
      if (noisy) {
        flat(k);
        a4(4, k) = T(0);
      } else {
        set_a6(k);
      }
    }
    if (iv == 0) cs_limiters1<T>(extm[k], a4(1, k), a4(2, k), a4(3, k), a4(4, k), 0);
  }

  // bottom
  if (iv == 0) {
    a4(3, km) = f_max(T(0), a4(3, km));
  } else if (iv == -1) {
    if (a4(3, km) * a4(1, km) <= T(0)) a4(3, km) = T(0);
  }
  for (int k = km - 1; k <= km; ++k) {
    set_a6(k);
    if (k == km - 1) cs_limiters1<T>(extm[k], a4(1, k), a4(2, k), a4(3, k), a4(4, k), 2);
    if (k == km) cs_limiters1<T>(extm[k], a4(1, k), a4(2, k), a4(3, k), a4(4, k), 1);
  }
}

// ppm_limiters (fv_mapz.F90:2840-2916) for one element
template <class T> static inline void ppm_limiters1(T dm, T& a1, T& a2, T& a3, T& a4, int lmt) {
  const T r12 = MapConst<T>::r12;
  if (lmt == 3) return;
  if (lmt == 0) {
    if (dm == T(0)) {
      a2 = a1;
      a3 = a1;
      a4 = T(0);
    } else {
      T da1 = a3 - a2;
      T da2 = da1 * da1;
      T a6da = a4 * da1;
      if (a6da < -da2) {
        a4 = T(3) * (a2 - a1);
        a3 = a2 - a4;
      } else if (a6da > da2) {
        a4 = T(3) * (a3 - a1);
        a2 = a3 - a4;
      }
    }
  } else if (lmt == 1) {
    T qmp = T(2) * dm;
    a2 = a1 - f_sign(f_min(f_abs(qmp), f_abs(a2 - a1)), qmp);
    a3 = a1 + f_sign(f_min(f_abs(qmp), f_abs(a3 - a1)), qmp);
    a4 = T(3) * (T(2) * a1 - (a2 + a3));
  } else if (lmt == 2) {
    if (f_abs(a3 - a2) < -a4) {
      T fmin = a1 + T(0.25) * ((a3 - a2) * (a3 - a2)) / a4 + a4 * r12;
      if (fmin < T(0)) {
        if (a1 < a3 && a1 < a2) {
          a3 = a1;
          a2 = a1;
          a4 = T(0);
        } else if (a3 > a2) {
          a4 = T(3) * (a2 - a1);
          a3 = a2 - a4;
        } else {
          a4 = T(3) * (a3 - a1);
          a2 = a3 - a4;
        }
      }
    }
  }
}

// ppm_profile (fv_mapz.F90:2580-2837) for one column (BOT_MONO not defined)
template <class T>
static void ppm_profile_col(ColA4<T>& a4, const T* delp /*1-based*/, int km, int iv, int kord, ProfScratch<T>& w) {
  T* dc = w.dc.data();
  T* h2 = w.h2.data();
  T* delq = w.delq.data();
  T* df2 = w.df2.data();
  T* d4 = w.d4.data();
  const int km1 = km - 1;
  for (int k = 2; k <= km; ++k) {
    delq[k - 1] = a4(1, k) - a4(1, k - 1);
    d4[k] = delp[k - 1] + delp[k];
  }
  for (int k = 2; k <= km1; ++k) {
    T c1 = (delp[k - 1] + T(0.5) * delp[k]) / d4[k + 1];
    T c2 = (delp[k + 1] + T(0.5) * delp[k]) / d4[k];
    df2[k] = delp[k] * (c1 * delq[k] + c2 * delq[k - 1]) / (d4[k] + delp[k + 1]);
    dc[k] = f_sign(f_min(f_abs(df2[k]), f_max(a4(1, k - 1), a4(1, k), a4(1, k + 1)) - a4(1, k),
                         a4(1, k) - f_min(a4(1, k - 1), a4(1, k), a4(1, k + 1))),
                   df2[k]);
  }
  for (int k = 3; k <= km1; ++k) {
    T c1 = delq[k - 1] * delp[k - 1] / d4[k];
    T a1 = d4[k - 1] / (d4[k] + delp[k - 1]);
    T a2 = d4[k + 1] / (d4[k] + delp[k]);
    a4(2, k) = a4(1, k - 1) + c1 +
               T(2) / (d4[k - 1] + d4[k + 1]) * (delp[k] * (c1 * (a1 - a2) + a2 * dc[k - 1]) - delp[k - 1] * a1 * dc[k]);
  }
  {  // top
    T d1 = delp[1], d2 = delp[2];
    T qm = (d2 * a4(1, 1) + d1 * a4(1, 2)) / (d1 + d2);
    T dq = T(2) * (a4(1, 2) - a4(1, 1)) / (d1 + d2);
    T c1 = T(4) * (a4(2, 3) - qm - d2 * dq) / (d2 * (T(2) * d2 * d2 + d1 * (d2 + T(3) * d1)));
    T c3 = dq - T(0.5) * c1 * (d2 * (T(5) * d1 + d2) - T(3) * d1 * d1);
    a4(2, 2) = qm - T(0.25) * c1 * d1 * d2 * (d2 + T(3) * d1);
    a4(2, 1) = d1 * (T(2) * c1 * (d1 * d1) - c3) + a4(2, 2);
    a4(2, 2) = f_max(a4(2, 2), f_min(a4(1, 1), a4(1, 2)));
    a4(2, 2) = f_min(a4(2, 2), f_max(a4(1, 1), a4(1, 2)));
    dc[1] = T(0.5) * (a4(2, 2) - a4(1, 1));
  }
  if (iv == 0) {
    a4(2, 1) = f_max(T(0), a4(2, 1));
    a4(2, 2) = f_max(T(0), a4(2, 2));
  } else if (iv == -1) {
    if (a4(2, 1) * a4(1, 1) <= T(0)) a4(2, 1) = T(0);
  } else if (std::abs(iv) == 2) {
    a4(2, 1) = a4(1, 1);
    a4(3, 1) = a4(1, 1);
  }
  {  // bottom
    T d1 = delp[km], d2 = delp[km1];
    T qm = (d2 * a4(1, km) + d1 * a4(1, km1)) / (d1 + d2);
    T dq = T(2) * (a4(1, km1) - a4(1, km)) / (d1 + d2);
    T c1 = (a4(2, km1) - qm - d2 * dq) / (d2 * (T(2) * d2 * d2 + d1 * (d2 + T(3) * d1)));
    T c3 = dq - T(2.0) * c1 * (d2 * (T(5) * d1 + d2) - T(3) * d1 * d1);
    a4(2, km) = qm - c1 * d1 * d2 * (d2 + T(3) * d1);
    a4(3, km) = d1 * (T(8) * c1 * (d1 * d1) - c3) + a4(2, km);
    a4(2, km) = f_max(a4(2, km), f_min(a4(1, km), a4(1, km1)));
    a4(2, km) = f_min(a4(2, km), f_max(a4(1, km), a4(1, km1)));
    dc[km] = T(0.5) * (a4(1, km) - a4(2, km));
  }
  if (iv == 0) {
    a4(2, km) = f_max(T(0), a4(2, km));
    a4(3, km) = f_max(T(0), a4(3, km));
  } else if (iv < 0) {
    if (a4(1, km) * a4(3, km) <= T(0)) a4(3, km) = T(0);
  }
  for (int k = 1; k <= km1; ++k) a4(3, k) = a4(2, k + 1);

  for (int k = 1; k <= 2; ++k) {
    a4(4, k) = T(3) * (T(2) * a4(1, k) - (a4(2, k) + a4(3, k)));
    ppm_limiters1<T>(dc[k], a4(1, k), a4(2, k), a4(3, k), a4(4, k), 0);
  }
  if (kord >= 7) {
    for (int k = 2; k <= km1; ++k) {
      h2[k] = T(2) * (dc[k + 1] / delp[k + 1] - dc[k - 1] / delp[k - 1]) / (delp[k] + T(0.5) * (delp[k - 1] + delp[k + 1])) *
              (delp[k] * delp[k]);
    }
    const T fac = T(1.5);
    for (int k = 3; k <= km - 2; ++k) {
      T pmp = T(2) * dc[k];
      T qmp = a4(1, k) + pmp;
      T lac = a4(1, k) + fac * h2[k - 1] + dc[k];
      a4(3, k) = f_min(f_max(a4(3, k), f_min(a4(1, k), qmp, lac)), f_max(a4(1, k), qmp, lac));
      qmp = a4(1, k) - pmp;
      lac = a4(1, k) + fac * h2[k + 1] - dc[k];
      a4(2, k) = f_min(f_max(a4(2, k), f_min(a4(1, k), qmp, lac)), f_max(a4(1, k), qmp, lac));
      a4(4, k) = T(3) * (T(2) * a4(1, k) - (a4(2, k) + a4(3, k)));
      if (iv == 0 && kord >= 6) ppm_limiters1<T>(dc[k], a4(1, k), a4(2, k), a4(3, k), a4(4, k), 2);
    }
  } else {
    int lmt = kord - 3;
    lmt = std::max(0, lmt);
    if (iv == 0) lmt = std::min(2, lmt);
    for (int k = 3; k <= km - 2; ++k) {
      if (kord != 4) a4(4, k) = T(3) * (T(2) * a4(1, k) - (a4(2, k) + a4(3, k)));
      if (kord != 6) ppm_limiters1<T>(dc[k], a4(1, k), a4(2, k), a4(3, k), a4(4, k), lmt);
    }
  }
  for (int k = km1; k <= km; ++k) {
    a4(4, k) = T(3) * (T(2) * a4(1, k) - (a4(2, k) + a4(3, k)));
    ppm_limiters1<T>(dc[k], a4(1, k), a4(2, k), a4(3, k), a4(4, k), 0);
  }
}

// fillz (fv_fill.F90:86-153) for one column, all tracers: q(k, ic) 1-based via accessor
template <class T, class Q> static void fillz_col(int km, int nq, Q&& q, const T* dp /*1-based*/, std::vector<T>& dm) {
  dm.resize(km + 2);
  for (int ic = 1; ic <= nq; ++ic) {
    if (q(1, ic) < T(0)) {
      q(2, ic) = q(2, ic) + q(1, ic) * dp[1] / dp[2];
      q(1, ic) = T(0);
    }
    bool zfix = false;
    for (int k = 2; k <= km - 1; ++k) {
      if (q(k, ic) < T(0)) {
        zfix = true;
        if (q(k - 1, ic) > T(0)) {
          T dq = f_min(q(k - 1, ic) * dp[k - 1], -q(k, ic) * dp[k]);
          q(k - 1, ic) = q(k - 1, ic) - dq / dp[k - 1];
          q(k, ic) = q(k, ic) + dq / dp[k];
        }
        if (q(k, ic) < T(0.0) && q(k + 1, ic) > T(0)) {
          T dq = f_min(q(k + 1, ic) * dp[k + 1], -q(k, ic) * dp[k]);
          q(k + 1, ic) = q(k + 1, ic) - dq / dp[k + 1];
          q(k, ic) = q(k, ic) + dq / dp[k];
        }
      }
    }
    {
      const int k = km;
      if (q(k, ic) < T(0) && q(k - 1, ic) > T(0)) {
        zfix = true;
        T qup = q(k - 1, ic) * dp[k - 1];
        T qly = -q(k, ic) * dp[k];
        T dup = f_min(qly, qup);
        q(k - 1, ic) = q(k - 1, ic) - dup / dp[k - 1];
        q(k, ic) = q(k, ic) + dup / dp[k];
      }
    }
    if (zfix) {
      T sum0 = T(0);
      for (int k = 2; k <= km; ++k) {
        dm[k] = q(k, ic) * dp[k];
        sum0 = sum0 + dm[k];
      }
      if (sum0 > T(0)) {
        T sum1 = T(0);
        for (int k = 2; k <= km; ++k) sum1 = sum1 + f_max(T(0), dm[k]);
        T fac = sum0 / sum1;
        for (int k = 2; k <= km; ++k) q(k, ic) = f_max(T(0), fac * dm[k] / dp[k]);
      }
    }
  }
}

template <class T> struct RemapScratch {
  std::vector<ColA4<T>> a4;  // per tracer
  std::vector<T> pe1, pe2, dp1, dp2, q2, qsum, dm;
  ProfScratch<T> prof;
  void size(int km, int nq) {
    a4.resize(nq);
    for (auto& a : a4) a.size(km);
    pe1.resize(km + 3);
    pe2.resize(km + 3);
    dp1.resize(km + 3);
    dp2.resize(km + 3);
    q2.resize((size_t)(km + 2) * nq);
    qsum.resize(nq);
    prof.size(km);
  }
};

// mapn_tracer (fv_mapz.F90:1386-1499) for one column i of row j.  a4[iq](1,k) must hold q1(i,j,k,iq);
// on return q2(k,iq) holds the remapped (and, if fill, filled) values.
template <class T>
static void mapn_tracer_col(int nq, int km, const T* pe1, const T* pe2, const T* dp2, const int* kord, T q_min, bool fill,
                            RemapScratch<T>& w) {
  const T r3 = MapConst<T>::r3, r23 = MapConst<T>::r23;
  T* dp1 = w.dp1.data();
  for (int k = 1; k <= km; ++k) dp1[k] = pe1[k + 1] - pe1[k];
  for (int iq = 0; iq < nq; ++iq) cs_profile_col<T>(true, T(0) /*qs unset in the reference*/, w.a4[iq], dp1, km, 0, kord[iq], q_min, w.prof);
  auto q2 = [&](int k, int iq1) -> T& { return w.q2[(size_t)(iq1 - 1) * (km + 2) + k]; };
  T* qsum = w.qsum.data();
  int k0 = 1;
  for (int k = 1; k <= km; ++k) {
    bool direct = false;
    for (int l = k0; l <= km; ++l) {
      if (pe2[k] >= pe1[l] && pe2[k] <= pe1[l + 1]) {
        T pl = (pe2[k] - pe1[l]) / dp1[l];
        if (pe2[k + 1] <= pe1[l + 1]) {
          T pr = (pe2[k + 1] - pe1[l]) / dp1[l];
          T fac1 = pr + pl;
          T fac2 = r3 * (pr * fac1 + pl * pl);
          fac1 = T(0.5) * fac1;
          for (int iq = 0; iq < nq; ++iq) {
            ColA4<T>& q4 = w.a4[iq];
            q2(k, iq + 1) = q4(2, l) + (q4(4, l) + q4(3, l) - q4(2, l)) * fac1 - q4(4, l) * fac2;
          }
          k0 = l;
          direct = true;
          break;  // goto 555
        } else {
          T dp = pe1[l + 1] - pe2[k];
          T fac1 = T(1) + pl;
          T fac2 = r3 * (T(1) + pl * fac1);
          fac1 = T(0.5) * fac1;
          for (int iq = 0; iq < nq; ++iq) {
            ColA4<T>& q4 = w.a4[iq];
            qsum[iq] = dp * (q4(2, l) + (q4(4, l) + q4(3, l) - q4(2, l)) * fac1 - q4(4, l) * fac2);
          }
          for (int m = l + 1; m <= km; ++m) {
            if (pe2[k + 1] > pe1[m + 1]) {
              for (int iq = 0; iq < nq; ++iq) qsum[iq] = qsum[iq] + dp1[m] * w.a4[iq](1, m);
            } else {
              dp = pe2[k + 1] - pe1[m];
              T esl = dp / dp1[m];
              fac1 = T(0.5) * esl;
              fac2 = T(1) - r23 * esl;
              for (int iq = 0; iq < nq; ++iq) {
                ColA4<T>& q4 = w.a4[iq];
                qsum[iq] = qsum[iq] + dp * (q4(2, m) + fac1 * (q4(3, m) - q4(2, m) + q4(4, m) * fac2));
              }
              k0 = m;
              break;  // goto 123
            }
          }
          break;  // goto 123
        }
      }
    }
    if (!direct) {
      for (int iq = 0; iq < nq; ++iq) q2(k, iq + 1) = qsum[iq] / dp2[k];
    }
  }
  if (fill) fillz_col<T>(km, nq, q2, dp2, w.dm);
}

// map1_q2 (fv_mapz.F90:1502-1592) for one column: a4(1,k) holds q1; q2out(k) 1-based.  The same body serves map_scalar
// (:1199-1290: scalar_profile with the bottom value qs) and map1_ppm (:1293-1383: cs_profile, `scalar` = false): the three
// reference routines differ only in the profile they call and in dp2 being passed (map1_q2) or formed as pe2(k+1)-pe2(k).
template <class T>
static void map1_q2_col(int km, const T* pe1, ColA4<T>& q4, int kn, const T* pe2, T* q2out, const T* dp2, int iv, int kord,
                        T q_min, RemapScratch<T>& w, bool scalar = true, T qs = T(0)) {
  const T r3 = MapConst<T>::r3, r23 = MapConst<T>::r23;
  T* dp1 = w.dp1.data();
  for (int k = 1; k <= km; ++k) dp1[k] = pe1[k + 1] - pe1[k];
  if (kord > 7)
    cs_profile_col<T>(scalar, qs, q4, dp1, km, iv, kord, q_min, w.prof);
  else
    ppm_profile_col<T>(q4, dp1, km, iv, kord, w.prof);
  int k0 = 1;
  T qsum = T(0);
  for (int k = 1; k <= kn; ++k) {
    bool direct = false;
    for (int l = k0; l <= km; ++l) {
      if (pe2[k] >= pe1[l] && pe2[k] <= pe1[l + 1]) {
        T pl = (pe2[k] - pe1[l]) / dp1[l];
        if (pe2[k + 1] <= pe1[l + 1]) {
          T pr = (pe2[k + 1] - pe1[l]) / dp1[l];
          q2out[k] = q4(2, l) + T(0.5) * (q4(4, l) + q4(3, l) - q4(2, l)) * (pr + pl) - q4(4, l) * r3 * (pr * (pr + pl) + pl * pl);
          k0 = l;
          direct = true;
          break;
        } else {
          qsum = (pe1[l + 1] - pe2[k]) * (q4(2, l) + T(0.5) * (q4(4, l) + q4(3, l) - q4(2, l)) * (T(1) + pl) -
                                          q4(4, l) * (r3 * (T(1) + pl * (T(1) + pl))));
          for (int m = l + 1; m <= km; ++m) {
            if (pe2[k + 1] > pe1[m + 1]) {
              qsum = qsum + dp1[m] * q4(1, m);
            } else {
              T dp = pe2[k + 1] - pe1[m];
              T esl = dp / dp1[m];
              qsum = qsum + dp * (q4(2, m) + T(0.5) * esl * (q4(3, m) - q4(2, m) + q4(4, m) * (T(1) - r23 * esl)));
              k0 = m;
              break;
            }
          }
          break;
        }
      }
    }
    if (!direct) q2out[k] = qsum / dp2[k];
  }
}

// Tracer part of Lagrangian_to_Eulerian's j loop (fv_mapz.F90:261-273, 343-368, 407-426) for one tile:
//   pe(is-1:ie+1, km+1, js-1:je+1), q(isd:ied, jsd:jed, km, nq), delp(isd:ied, jsd:jed, km)
template <class T>
static void remap_tracers_tile(int n, int km, int nq, const T* pe, const T* ak, const T* bk, T ptop, T* q, T* delp,
                               const int* kord_tr, bool fill) {
  const Bounds bd = Bounds::tile(n);
  const int is = bd.is, ie = bd.ie, js = bd.js, je = bd.je, isd = bd.isd, jsd = bd.jsd;
  const long nxd = bd.ied - bd.isd + 1, nyd = bd.jed - bd.jsd + 1, plane = nxd * nyd;
  const long pe_ld1 = (ie + 1) - (is - 1) + 1, pe_ld2 = pe_ld1 * (km + 1);
#pragma omp parallel
  {
    RemapScratch<T> w;
    w.size(km, nq);
    std::vector<T> q2one(km + 2);
#pragma omp for schedule(static)
    for (int j = js; j <= je; ++j) {
      for (int i = is; i <= ie; ++i) {
        T* pe1 = w.pe1.data();
        T* pe2 = w.pe2.data();
        T* dp2 = w.dp2.data();
        auto PE = [&](int ii, int k, int jj) -> T { return pe[(long)(ii - (is - 1)) + (long)(k - 1) * pe_ld1 + (long)(jj - (js - 1)) * pe_ld2]; };
        for (int k = 1; k <= km + 1; ++k) pe1[k] = PE(i, k, j);
        pe2[1] = ptop;
        pe2[km + 1] = PE(i, km + 1, j);
        for (int k = 2; k <= km; ++k) pe2[k] = ak[k - 1] + bk[k - 1] * PE(i, km + 1, j);
        for (int k = 1; k <= km; ++k) dp2[k] = pe2[k + 1] - pe2[k];
        const long col = (long)(i - isd) + (long)(j - jsd) * nxd;
        for (int k = 1; k <= km; ++k) delp[col + (long)(k - 1) * plane] = dp2[k];
        if (nq > 5) {
          for (int iq = 0; iq < nq; ++iq)
            for (int k = 1; k <= km; ++k) w.a4[iq](1, k) = q[col + ((long)iq * km + (k - 1)) * plane];
          mapn_tracer_col<T>(nq, km, pe1, pe2, dp2, kord_tr, T(0), fill, w);
          for (int iq = 0; iq < nq; ++iq)
            for (int k = 1; k <= km; ++k) q[col + ((long)iq * km + (k - 1)) * plane] = w.q2[(size_t)iq * (km + 2) + k];
        } else if (nq > 0) {
          for (int iq = 0; iq < nq; ++iq) {
            for (int k = 1; k <= km; ++k) w.a4[0](1, k) = q[col + ((long)iq * km + (k - 1)) * plane];
            map1_q2_col<T>(km, pe1, w.a4[0], km, pe2, q2one.data(), dp2, 0, kord_tr[iq], T(0), w);
            if (fill) {
              auto qq = [&](int k, int) -> T& { return q2one[k]; };
              fillz_col<T>(km, 1, qq, dp2, w.dm);
            }
            for (int k = 1; k <= km; ++k) q[col + ((long)iq * km + (k - 1)) * plane] = q2one[k];
          }
        }
      }
    }
  }
}

}  // namespace fv3oracle
