// fv3atm_b200: vertical Lagrangian-to-Eulerian tracer remap kernels for sm_100a.
//
// Replaces the tracer part of Lagrangian_to_Eulerian's j loop (atmos_cubed_sphere/model/fv_mapz.F90:261-273,
// 343-368, 407-426): mapn_tracer (:1386-1499) / map1_q2 (:1502-1592) with scalar_profile (:1691-2096),
// cs_limiters (:2501-2576), ppm_profile / ppm_limiters (:2580-2916) and fillz (model/fv_fill.F90:86-153).
// One thread owns one column (i,j); threads of a warp own 32 consecutive i, so every load/store of
// q(i,j,k,iq), pe(i,k,j) and delp(i,j,k) is a coalesced i-contiguous row segment.  The column state lives in
// thread-private arrays; per-level values are processed in the reference's operation order.
#pragma once
#include "fv3t_common.cuh"

namespace fv3t {

template <class T> struct RemapParams {
  const T* q0;   // ping-pong buffers of q (level kz currently lives in buffer par[kz])
  const T* q1;
  T* qout;       // destination buffer (buffer 0)
  const int* par;  // [km]
  const T* pe;     // (is-1:ie+1, km+1, js-1:je+1), tile-major
  const T *ak, *bk;
  T* delp;         // (isd:ied, jsd:jed, km), tile-major
  const int* kord;  // [nq]
  T ptop;
  int n, km, nq, ntiles, fill;
  int j_first, j_count;  // rows js+j_first .. (row-granular compatibility entry uses j_count = 1)
};

// cs_limiters (fv_mapz.F90:2501-2576), one element
template <class T> FV3T_HD void cs_limiters1(bool extm, T a1, T& a2, T& a3, T& a4, int iv) {
  if (iv == 0) {
    if (a1 <= T(0)) {
      a2 = a1;
      a3 = a1;
      a4 = T(0);
    } else if (f_abs(a3 - a2) < -a4) {
      if ((a1 + T(0.25) * ((a3 - a2) * (a3 - a2)) / a4 + a4 * K<T>::r12()) < T(0)) {
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
  } else {
    const bool flat = (iv == 1) ? ((a1 - a2) * (a1 - a3) >= T(0)) : extm;
    if (flat) {
      a2 = a1;
      a3 = a1;
      a4 = T(0);
    } else {
      const T da1 = a3 - a2;
      const T da2 = da1 * da1;
      const T a6da = a4 * da1;
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

// ppm_limiters (fv_mapz.F90:2840-2916), one element
template <class T> FV3T_HD void ppm_limiters1(T dm, T a1, T& a2, T& a3, T& a4, int lmt) {
  if (lmt == 3) return;
  if (lmt == 0) {
    if (dm == T(0)) {
      a2 = a1;
      a3 = a1;
      a4 = T(0);
    } else {
      const T da1 = a3 - a2;
      const T da2 = da1 * da1;
      const T a6da = a4 * da1;
      if (a6da < -da2) {
        a4 = T(3) * (a2 - a1);
        a3 = a2 - a4;
      } else if (a6da > da2) {
        a4 = T(3) * (a3 - a1);
        a2 = a3 - a4;
      }
    }
  } else if (lmt == 1) {
    const T qmp = T(2) * dm;
    a2 = a1 - f_sign(f_min(f_abs(qmp), f_abs(a2 - a1)), qmp);
    a3 = a1 + f_sign(f_min(f_abs(qmp), f_abs(a3 - a1)), qmp);
    a4 = T(3) * (T(2) * a1 - (a2 + a3));
  } else if (lmt == 2) {
    if (f_abs(a3 - a2) < -a4) {
      const T fmin = a1 + T(0.25) * ((a3 - a2) * (a3 - a2)) / a4 + a4 * K<T>::r12();
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

// Column workspace (thread-private; 1-based like the Fortran)
template <class T, int KM> struct ColWork {
  T a1[KM + 2], a2[KM + 2], a3[KM + 2], a4[KM + 2];
  T gam[KM + 3], qv[KM + 3];
  T pe1[KM + 2], pe2[KM + 2], dp1[KM + 2], dp2[KM + 2];
  T q2[KM + 2];
  unsigned char fl[KM + 3];  // bit0 extm, bit1 ext5, bit2 ext6
};

// scalar_profile (fv_mapz.F90:1691-2096; SCALAR = true) and cs_profile (:2098-2498; SCALAR = false) for one column.  The two
// reference routines share the spline, the interface constraints and the limiter cascade; cs_profile has no qmin (the
// |kord| = 9, 11, 15 branches differ) and forms a6 of |kord| = 9 as 6*a1 - 3*(a2+a3) (:2324) where scalar_profile has
// 3*(2*a1 - (a2+a3)) (:1921) -- a different rounding.  qs (bottom boundary value) is read for iv = -2 only.
template <class T, int KM, bool SCALAR> __device__ void profile_col(ColWork<T, KM>& w, int km, int iv, int kord, T qmin, T qs) {
  T* a1 = w.a1;
  T* a2 = w.a2;
  T* a3 = w.a3;
  T* a4 = w.a4;
  T* gam = w.gam;
  T* q = w.qv;
  const T* delp = w.dp1;
  unsigned char* fl = w.fl;
  const int akord = kord < 0 ? -kord : kord;
  T d4 = T(0);
  if (iv == -2) {  // spline with the bottom interface value given (fv_mapz.F90:1719-1735 / :2126-2142)
    gam[2] = T(0.5);
    q[1] = T(1.5) * a1[1];
    for (int k = 2; k <= km - 1; ++k) {
      const T grat = delp[k - 1] / delp[k];
      const T bet = T(2) + grat + grat - gam[k];
      q[k] = (T(3) * (a1[k - 1] + a1[k]) - q[k - 1]) / bet;
      gam[k + 1] = grat / bet;
    }
    const T grat = delp[km - 1] / delp[km];
    q[km] = (T(3) * (a1[km - 1] + a1[km]) - grat * qs - q[km - 1]) / (T(2) + grat + grat - gam[km]);
    q[km + 1] = qs;
    for (int k = km - 1; k >= 1; --k) q[k] = q[k] - gam[k + 1] * q[k + 1];
  } else {
    const T grat = delp[2] / delp[1];
    T bet = grat * (grat + T(0.5));
    q[1] = ((grat + grat) * (grat + T(1)) * a1[1] + a1[2]) / bet;
    gam[1] = (T(1) + grat * (grat + T(1.5))) / bet;
    for (int k = 2; k <= km; ++k) {
      d4 = delp[k - 1] / delp[k];
      bet = T(2) + d4 + d4 - gam[k - 1];
      q[k] = (T(3) * (a1[k - 1] + d4 * a1[k]) - q[k - 1]) / bet;
      gam[k] = d4 / bet;
    }
    const T a_bot = T(1) + d4 * (d4 + T(1.5));
    q[km + 1] = (T(2) * d4 * (d4 + T(1)) * a1[km] + a1[km - 1] - a_bot * q[km]) / (d4 * (d4 + T(0.5)) - a_bot * gam[km]);
    for (int k = km; k >= 1; --k) q[k] = q[k] - gam[k] * q[k + 1];
  }
  if (akord > 16) {
    for (int k = 1; k <= km; ++k) {
      a2[k] = q[k];
      a3[k] = q[k + 1];
      a4[k] = T(3) * (T(2) * a1[k] - (a2[k] + a3[k]));
    }
    return;
  }
  q[2] = f_min(q[2], f_max(a1[1], a1[2]));
  q[2] = f_max(q[2], f_min(a1[1], a1[2]));
  for (int k = 2; k <= km; ++k) gam[k] = a1[k] - a1[k - 1];
  for (int k = 3; k <= km - 1; ++k) {
    if (gam[k - 1] * gam[k + 1] > T(0)) {
      q[k] = f_min(q[k], f_max(a1[k - 1], a1[k]));
      q[k] = f_max(q[k], f_min(a1[k - 1], a1[k]));
    } else if (gam[k - 1] > T(0)) {
      q[k] = f_max(q[k], f_min(a1[k - 1], a1[k]));
    } else {
      q[k] = f_min(q[k], f_max(a1[k - 1], a1[k]));
      if (iv == 0) q[k] = f_max(T(0), q[k]);
    }
  }
  q[km] = f_min(q[km], f_max(a1[km - 1], a1[km]));
  q[km] = f_max(q[km], f_min(a1[km - 1], a1[km]));
  for (int k = 1; k <= km; ++k) {
    a2[k] = q[k];
    a3[k] = q[k + 1];
  }
  for (int k = 1; k <= km; ++k) {
    int f;
    if (k == 1 || k == km)
      f = ((a2[k] - a1[k]) * (a3[k] - a1[k]) > T(0)) ? 1 : 0;
    else
      f = (gam[k] * gam[k + 1] < T(0)) ? 1 : 0;
    if (akord > 9) {
      const T x0 = T(2) * a1[k] - (a2[k] + a3[k]);
      const T x1 = f_abs(a2[k] - a3[k]);
      a4[k] = T(3) * x0;
      if (f_abs(x0) > x1) f |= 2;
      if (f_abs(a4[k]) > x1) f |= 4;
    }
    fl[k] = (unsigned char)f;
  }
  if (iv == 0) {
    a2[1] = f_max(T(0), a2[1]);
  } else if (iv == -1) {
    if (a2[1] * a1[1] <= T(0)) a2[1] = T(0);
  } else if (iv == 2) {
    a2[1] = a1[1];
    a3[1] = a1[1];
    a4[1] = T(0);
  }
  if (iv != 2) {
    a4[1] = T(3) * (T(2) * a1[1] - (a2[1] + a3[1]));
    cs_limiters1<T>(fl[1] & 1, a1[1], a2[1], a3[1], a4[1], 1);
  }
  a4[2] = T(3) * (T(2) * a1[2] - (a2[2] + a3[2]));
  cs_limiters1<T>(fl[2] & 1, a1[2], a2[2], a3[2], a4[2], 2);

  for (int k = 3; k <= km - 2; ++k) {
    const T a1k = a1[k];
    T a2k = a2[k], a3k = a3[k], a4k = a4[k];
    const bool extm = fl[k] & 1, ext5 = fl[k] & 2, ext6 = fl[k] & 4;
    const bool extm_m = fl[k - 1] & 1, ext5_m = fl[k - 1] & 2, ext6_m = fl[k - 1] & 4;
    const bool extm_p = fl[k + 1] & 1, ext5_p = fl[k + 1] & 2, ext6_p = fl[k + 1] & 4;
    auto huynh = [&]() {
      const T pmp_1 = a1k - T(2) * gam[k + 1];
      const T lac_1 = pmp_1 + T(1.5) * gam[k + 2];
      a2k = f_min(f_max(a2k, f_min(a1k, pmp_1, lac_1)), f_max(a1k, pmp_1, lac_1));
      const T pmp_2 = a1k + T(2) * gam[k];
      const T lac_2 = pmp_2 - T(1.5) * gam[k - 1];
      a3k = f_min(f_max(a3k, f_min(a1k, pmp_2, lac_2)), f_max(a1k, pmp_2, lac_2));
    };
    auto set_a6 = [&]() { a4k = T(3) * (T(2) * a1k - (a2k + a3k)); };
    auto flat = [&]() {
      a2k = a1k;
      a3k = a1k;
    };
    if (akord < 9) {
      huynh();
      set_a6();
    } else if (akord == 9) {
      if ((extm && extm_m) || (extm && extm_p) || (SCALAR && extm && a1k < qmin)) {
        flat();
        a4k = T(0);
      } else {
        if (SCALAR) set_a6(); else a4k = T(6) * a1k - T(3) * (a2k + a3k);
        if (f_abs(a4k) > f_abs(a2k - a3k)) {
          huynh();
          if (SCALAR) set_a6(); else a4k = T(6) * a1k - T(3) * (a2k + a3k);
        }
      }
    } else if (akord == 10) {
      if (ext5) {
        if (ext5_m || ext5_p)
          flat();
        else if (ext6_m || ext6_p)
          huynh();
      } else if (ext6) {
        if (ext5_m || ext5_p) huynh();
      }
      set_a6();
    } else if (akord == 12) {
      if (extm) {
        flat();
        a4k = T(0);
      } else {
        a4k = T(6) * a1k - T(3) * (a2k + a3k);
        if (f_abs(a4k) > f_abs(a2k - a3k)) {
          huynh();
          a4k = T(6) * a1k - T(3) * (a2k + a3k);
        }
      }
    } else if (akord == 13) {
      if (ext6 && ext6_m && ext6_p) flat();
      set_a6();
    } else if (akord == 14) {
      set_a6();
    } else if (akord == 15) {
      if (SCALAR) {
        if ((ext5 && ext5_m) || (ext5 && ext5_p) || (ext5 && a1k < qmin))
          flat();
        else if (ext6)
          huynh();
      } else {  // cs_profile (:2394-2404): the ext6 branch is reached only when ext5 is false
        if (ext5) {
          if (ext5_m || ext5_p) flat();
        } else if (ext6) {
          huynh();
        }
      }
      set_a6();
    } else if (akord == 16) {
      if (ext5) {
        if (ext5_m || ext5_p)
          flat();
        else if (ext6_m || ext6_p)
          huynh();
      }
      set_a6();
    } else {  // 11
      if (ext5 && (ext5_m || ext5_p || (SCALAR && a1k < qmin))) {
        flat();
        a4k = T(0);
      } else {
        set_a6();
      }
    }
    if (iv == 0) cs_limiters1<T>(extm, a1k, a2k, a3k, a4k, 0);
    a2[k] = a2k;
    a3[k] = a3k;
    a4[k] = a4k;
  }
  if (iv == 0) {
    a3[km] = f_max(T(0), a3[km]);
  } else if (iv == -1) {
    if (a3[km] * a1[km] <= T(0)) a3[km] = T(0);
  }
  for (int k = km - 1; k <= km; ++k) {
    a4[k] = T(3) * (T(2) * a1[k] - (a2[k] + a3[k]));
    cs_limiters1<T>(fl[k] & 1, a1[k], a2[k], a3[k], a4[k], k == km ? 1 : 2);
  }
}

template <class T, int KM> __device__ void scalar_profile_col(ColWork<T, KM>& w, int km, int iv, int kord, T qmin) {
  profile_col<T, KM, true>(w, km, iv, kord, qmin, T(0));
}

// ppm_profile (fv_mapz.F90:2580-2837), reached through map1_q2 when kord <= 7.  Scratch: gam <- dc,
// qv <- h2 / delq, q2 <- df2 / d4 are reused as work arrays.
template <class T, int KM> __device__ void ppm_profile_col(ColWork<T, KM>& w, int km, int iv, int kord) {
  T* a1 = w.a1;
  T* a2 = w.a2;
  T* a3 = w.a3;
  T* a4 = w.a4;
  const T* delp = w.dp1;
  T* dc = w.gam;
  T* h2 = w.qv;
  // d4(k) = delp(k-1)+delp(k) and delq(k) = a1(k+1)-a1(k) are cheap: recomputed where used (same rounding)
  auto d4 = [&](int k) -> T { return delp[k - 1] + delp[k]; };
  auto delq = [&](int k) -> T { return a1[k + 1] - a1[k]; };
  const int km1 = km - 1;
  for (int k = 2; k <= km1; ++k) {
    const T c1 = (delp[k - 1] + T(0.5) * delp[k]) / d4(k + 1);
    const T c2 = (delp[k + 1] + T(0.5) * delp[k]) / d4(k);
    const T df2 = delp[k] * (c1 * delq(k) + c2 * delq(k - 1)) / (d4(k) + delp[k + 1]);
    dc[k] = f_sign(f_min(f_abs(df2), f_max(a1[k - 1], a1[k], a1[k + 1]) - a1[k], a1[k] - f_min(a1[k - 1], a1[k], a1[k + 1])), df2);
  }
  for (int k = 3; k <= km1; ++k) {
    const T c1 = delq(k - 1) * delp[k - 1] / d4(k);
    const T aa1 = d4(k - 1) / (d4(k) + delp[k - 1]);
    const T aa2 = d4(k + 1) / (d4(k) + delp[k]);
    a2[k] = a1[k - 1] + c1 + T(2) / (d4(k - 1) + d4(k + 1)) * (delp[k] * (c1 * (aa1 - aa2) + aa2 * dc[k - 1]) - delp[k - 1] * aa1 * dc[k]);
  }
  {
    const T d1 = delp[1], d2 = delp[2];
    const T qm = (d2 * a1[1] + d1 * a1[2]) / (d1 + d2);
    const T dq = T(2) * (a1[2] - a1[1]) / (d1 + d2);
    const T c1 = T(4) * (a2[3] - qm - d2 * dq) / (d2 * (T(2) * d2 * d2 + d1 * (d2 + T(3) * d1)));
    const T c3 = dq - T(0.5) * c1 * (d2 * (T(5) * d1 + d2) - T(3) * d1 * d1);
    a2[2] = qm - T(0.25) * c1 * d1 * d2 * (d2 + T(3) * d1);
    a2[1] = d1 * (T(2) * c1 * (d1 * d1) - c3) + a2[2];
    a2[2] = f_max(a2[2], f_min(a1[1], a1[2]));
    a2[2] = f_min(a2[2], f_max(a1[1], a1[2]));
    dc[1] = T(0.5) * (a2[2] - a1[1]);
  }
  if (iv == 0) {
    a2[1] = f_max(T(0), a2[1]);
    a2[2] = f_max(T(0), a2[2]);
  } else if (iv == -1) {
    if (a2[1] * a1[1] <= T(0)) a2[1] = T(0);
  } else if (iv == 2 || iv == -2) {
    a2[1] = a1[1];
    a3[1] = a1[1];
  }
  {
    const T d1 = delp[km], d2 = delp[km1];
    const T qm = (d2 * a1[km] + d1 * a1[km1]) / (d1 + d2);
    const T dq = T(2) * (a1[km1] - a1[km]) / (d1 + d2);
    const T c1 = (a2[km1] - qm - d2 * dq) / (d2 * (T(2) * d2 * d2 + d1 * (d2 + T(3) * d1)));
    const T c3 = dq - T(2.0) * c1 * (d2 * (T(5) * d1 + d2) - T(3) * d1 * d1);
    a2[km] = qm - c1 * d1 * d2 * (d2 + T(3) * d1);
    a3[km] = d1 * (T(8) * c1 * (d1 * d1) - c3) + a2[km];
    a2[km] = f_max(a2[km], f_min(a1[km], a1[km1]));
    a2[km] = f_min(a2[km], f_max(a1[km], a1[km1]));
    dc[km] = T(0.5) * (a1[km] - a2[km]);
  }
  if (iv == 0) {
    a2[km] = f_max(T(0), a2[km]);
    a3[km] = f_max(T(0), a3[km]);
  } else if (iv < 0) {
    if (a1[km] * a3[km] <= T(0)) a3[km] = T(0);
  }
  for (int k = 1; k <= km1; ++k) a3[k] = a2[k + 1];
  for (int k = 1; k <= 2; ++k) {
    a4[k] = T(3) * (T(2) * a1[k] - (a2[k] + a3[k]));
    ppm_limiters1<T>(dc[k], a1[k], a2[k], a3[k], a4[k], 0);
  }
  if (kord >= 7) {
    for (int k = 2; k <= km1; ++k)
      h2[k] = T(2) * (dc[k + 1] / delp[k + 1] - dc[k - 1] / delp[k - 1]) / (delp[k] + T(0.5) * (delp[k - 1] + delp[k + 1])) *
              (delp[k] * delp[k]);
    const T fac = T(1.5);
    for (int k = 3; k <= km - 2; ++k) {
      const T pmp = T(2) * dc[k];
      T qmp = a1[k] + pmp;
      T lac = a1[k] + fac * h2[k - 1] + dc[k];
      a3[k] = f_min(f_max(a3[k], f_min(a1[k], qmp, lac)), f_max(a1[k], qmp, lac));
      qmp = a1[k] - pmp;
      lac = a1[k] + fac * h2[k + 1] - dc[k];
      a2[k] = f_min(f_max(a2[k], f_min(a1[k], qmp, lac)), f_max(a1[k], qmp, lac));
      a4[k] = T(3) * (T(2) * a1[k] - (a2[k] + a3[k]));
      if (iv == 0 && kord >= 6) ppm_limiters1<T>(dc[k], a1[k], a2[k], a3[k], a4[k], 2);
    }
  } else {
    int lmt = kord - 3;
    lmt = lmt > 0 ? lmt : 0;
    if (iv == 0) lmt = lmt < 2 ? lmt : 2;
    for (int k = 3; k <= km - 2; ++k) {
      if (kord != 4) a4[k] = T(3) * (T(2) * a1[k] - (a2[k] + a3[k]));
      if (kord != 6) ppm_limiters1<T>(dc[k], a1[k], a2[k], a3[k], a4[k], lmt);
    }
  }
  for (int k = km1; k <= km; ++k) {
    a4[k] = T(3) * (T(2) * a1[k] - (a2[k] + a3[k]));
    ppm_limiters1<T>(dc[k], a1[k], a2[k], a3[k], a4[k], 0);
  }
}

// Overlap integration of one tracer's piecewise parabolae onto the target layers.
// MAPN = true: factored form of mapn_tracer (fv_mapz.F90:1434-1472); false: map1_q2's form (:1557-1581).
template <class T, int KM, bool MAPN> __device__ void map_col(ColWork<T, KM>& w, int km) {
  const T r3 = K<T>::r3(), r23 = K<T>::r23();
  const T *pe1 = w.pe1, *pe2 = w.pe2, *dp1 = w.dp1, *dp2 = w.dp2;
  const T *a1 = w.a1, *a2 = w.a2, *a3 = w.a3, *a4 = w.a4;
  int k0 = 1;
  T qsum = T(0);
  for (int k = 1; k <= km; ++k) {
    bool direct = false;
    for (int l = k0; l <= km; ++l) {
      if (pe2[k] >= pe1[l] && pe2[k] <= pe1[l + 1]) {
        const T pl = (pe2[k] - pe1[l]) / dp1[l];
        if (pe2[k + 1] <= pe1[l + 1]) {
          const T pr = (pe2[k + 1] - pe1[l]) / dp1[l];
          if (MAPN) {
            T fac1 = pr + pl;
            const T fac2 = r3 * (pr * fac1 + pl * pl);
            fac1 = T(0.5) * fac1;
            w.q2[k] = a2[l] + (a4[l] + a3[l] - a2[l]) * fac1 - a4[l] * fac2;
          } else {
            w.q2[k] = a2[l] + T(0.5) * (a4[l] + a3[l] - a2[l]) * (pr + pl) - a4[l] * r3 * (pr * (pr + pl) + pl * pl);
          }
          k0 = l;
          direct = true;
          break;
        } else {
          T dp = pe1[l + 1] - pe2[k];
          if (MAPN) {
            T fac1 = T(1) + pl;
            const T fac2 = r3 * (T(1) + pl * fac1);
            fac1 = T(0.5) * fac1;
            qsum = dp * (a2[l] + (a4[l] + a3[l] - a2[l]) * fac1 - a4[l] * fac2);
          } else {
            qsum = dp * (a2[l] + T(0.5) * (a4[l] + a3[l] - a2[l]) * (T(1) + pl) - a4[l] * (r3 * (T(1) + pl * (T(1) + pl))));
          }
          for (int m = l + 1; m <= km; ++m) {
            if (pe2[k + 1] > pe1[m + 1]) {
              qsum = qsum + dp1[m] * a1[m];
            } else {
              dp = pe2[k + 1] - pe1[m];
              const T esl = dp / dp1[m];
              if (MAPN) {
                const T fac1 = T(0.5) * esl;
                const T fac2 = T(1) - r23 * esl;
                qsum = qsum + dp * (a2[m] + fac1 * (a3[m] - a2[m] + a4[m] * fac2));
              } else {
                qsum = qsum + dp * (a2[m] + T(0.5) * esl * (a3[m] - a2[m] + a4[m] * (T(1) - r23 * esl)));
              }
              k0 = m;
              break;
            }
          }
          break;
        }
      }
    }
    if (!direct) w.q2[k] = qsum / dp2[k];
  }
}

// fillz (fv_fill.F90:86-153) for one tracer of one column, in place on q2
template <class T, int KM> __device__ void fillz_col(ColWork<T, KM>& w, int km) {
  T* q = w.q2;
  const T* dp = w.dp2;
  if (q[1] < T(0)) {
    q[2] = q[2] + q[1] * dp[1] / dp[2];
    q[1] = T(0);
  }
  bool zfix = false;
  for (int k = 2; k <= km - 1; ++k) {
    if (q[k] < T(0)) {
      zfix = true;
      if (q[k - 1] > T(0)) {
        const T dq = f_min(q[k - 1] * dp[k - 1], -q[k] * dp[k]);
        q[k - 1] = q[k - 1] - dq / dp[k - 1];
        q[k] = q[k] + dq / dp[k];
      }
      if (q[k] < T(0) && q[k + 1] > T(0)) {
        const T dq = f_min(q[k + 1] * dp[k + 1], -q[k] * dp[k]);
        q[k + 1] = q[k + 1] - dq / dp[k + 1];
        q[k] = q[k] + dq / dp[k];
      }
    }
  }
  {
    const int k = km;
    if (q[k] < T(0) && q[k - 1] > T(0)) {
      zfix = true;
      const T qup = q[k - 1] * dp[k - 1];
      const T qly = -q[k] * dp[k];
      const T dup = f_min(qly, qup);
      q[k - 1] = q[k - 1] - dup / dp[k - 1];
      q[k] = q[k] + dup / dp[k];
    }
  }
  if (zfix) {
    T sum0 = T(0);
    for (int k = 2; k <= km; ++k) sum0 = sum0 + q[k] * dp[k];
    if (sum0 > T(0)) {
      T sum1 = T(0);
      for (int k = 2; k <= km; ++k) sum1 = sum1 + f_max(T(0), q[k] * dp[k]);
      const T fac = sum0 / sum1;
      for (int k = 2; k <= km; ++k) q[k] = f_max(T(0), fac * (q[k] * dp[k]) / dp[k]);
    }
  }
}

template <class T, int KM> __global__ void __launch_bounds__(64) k_remap(const RemapParams<T> p) {
  const int n = p.n, km = p.km;
  const long nd = n + 6, plane = nd * nd;
  const int cols = n * p.j_count;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y;
  if (c >= cols) return;
  const int i = c % n + 1, j = c / n + 1 + p.j_first;
  ColWork<T, KM> w;
  // pe1 = pe(i,:,j); pe2 = ak + bk*ps; dp2; delp <- dp2   (fv_mapz.F90:263-272, 350-368)
  const long pe_ld1 = n + 2, pe_ld2 = pe_ld1 * (km + 1);
  const T* pe = p.pe + (long)t * pe_ld2 * (n + 2) + (long)i + (long)j * pe_ld2;  // (i-(is-1)) = i, (j-(js-1)) = j
  for (int k = 1; k <= km + 1; ++k) w.pe1[k] = pe[(long)(k - 1) * pe_ld1];
  const T ps = w.pe1[km + 1];
  w.pe2[1] = p.ptop;
  w.pe2[km + 1] = ps;
  for (int k = 2; k <= km; ++k) w.pe2[k] = p.ak[k - 1] + p.bk[k - 1] * ps;
  const long col = (long)(j + 2) * nd + (i + 2);
  T* delp = p.delp + (long)t * plane * km + col;
  for (int k = 1; k <= km; ++k) {
    w.dp2[k] = w.pe2[k + 1] - w.pe2[k];
    delp[(long)(k - 1) * plane] = w.dp2[k];
    w.dp1[k] = w.pe1[k + 1] - w.pe1[k];
  }
  const bool mapn = p.nq > 5;
  for (int iq = 0; iq < p.nq; ++iq) {
    const long qoff = (((long)t * p.nq + iq) * km) * plane + col;
    for (int k = 1; k <= km; ++k) {
      const T* src = p.par[k - 1] ? p.q1 : p.q0;
      w.a1[k] = src[qoff + (long)(k - 1) * plane];
    }
    const int kord = p.kord[iq];
    if (mapn || kord > 7)
      scalar_profile_col<T, KM>(w, km, 0, kord, T(0));
    else
      ppm_profile_col<T, KM>(w, km, 0, kord);
    if (mapn)
      map_col<T, KM, true>(w, km);
    else
      map_col<T, KM, false>(w, km);
    if (p.fill) fillz_col<T, KM>(w, km);
    for (int k = 1; k <= km; ++k) p.qout[qoff + (long)(k - 1) * plane] = w.q2[k];
  }
}

// map_scalar (fv_mapz.F90:1199-1290: scalar_profile) / map1_ppm (:1293-1383: cs_profile) for kn = km, all rows of the resident
// tiles at once: one field (isd:ied, jsd:jed, km) per tile, remapped in place from the resident Lagrangian pe onto
// pe2 = ak + bk*pe(km+1) (the target grid Lagrangian_to_Eulerian builds, :263-272).  kord <= 7 takes ppm_profile in both.
// These are the callers through which the reference reaches cs_profile (pt, w, delz, u, v: fv_mapz.F90:393-436, 610-660).
template <class T> struct MapFieldParams {
  T* q;           // (isd:ied, jsd:jed, km) tile-major, in/out
  const T* qs;    // (isd:ied, jsd:jed) tile-major bottom boundary value, read for iv = -2 only (may be null otherwise)
  const T* pe;    // (is-1:ie+1, km+1, js-1:je+1) tile-major
  const T *ak, *bk;
  T ptop, q_min;
  int n, km, ntiles, iv, kord, use_cs;
};
template <class T, int KM> __global__ void __launch_bounds__(64) k_map_field(const MapFieldParams<T> p) {
  const int n = p.n, km = p.km;
  const long nd = n + 6, plane = nd * nd;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y;
  if (c >= n * n) return;
  const int i = c % n + 1, j = c / n + 1;
  ColWork<T, KM> w;
  const long pe_ld1 = n + 2, pe_ld2 = pe_ld1 * (km + 1);
  const T* pe = p.pe + (long)t * pe_ld2 * (n + 2) + (long)i + (long)j * pe_ld2;
  for (int k = 1; k <= km + 1; ++k) w.pe1[k] = pe[(long)(k - 1) * pe_ld1];
  const T ps = w.pe1[km + 1];
  w.pe2[1] = p.ptop;
  w.pe2[km + 1] = ps;
  for (int k = 2; k <= km; ++k) w.pe2[k] = p.ak[k - 1] + p.bk[k - 1] * ps;
  const long col = (long)(j + 2) * nd + (i + 2);
  T* q = p.q + (long)t * plane * km + col;
  for (int k = 1; k <= km; ++k) {
    w.dp2[k] = w.pe2[k + 1] - w.pe2[k];
    w.dp1[k] = w.pe1[k + 1] - w.pe1[k];
    w.a1[k] = q[(long)(k - 1) * plane];
  }
  const T qs = (p.iv == -2 && p.qs) ? p.qs[(long)t * plane + col] : T(0);
  if (p.kord > 7) {
    if (p.use_cs)
      profile_col<T, KM, false>(w, km, p.iv, p.kord, T(0), qs);
    else
      profile_col<T, KM, true>(w, km, p.iv, p.kord, p.q_min, qs);
  } else {
    ppm_profile_col<T, KM>(w, km, p.iv, p.kord);
  }
  map_col<T, KM, false>(w, km);
  for (int k = 1; k <= km; ++k) q[(long)(k - 1) * plane] = w.q2[k];
}

}  // namespace fv3t
