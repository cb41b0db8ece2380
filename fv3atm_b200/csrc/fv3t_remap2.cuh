// fv3atm_b200: streaming vertical tracer remap (mapn_tracer / map1_q2 with scalar_profile) for sm_100a.
//
// Same arithmetic as fv3t_remap.cuh (operation order of atmos_cubed_sphere/model/fv_mapz.F90:1386-1499 mapn_tracer,
// :1502-1592 map1_q2, :1691-2096 scalar_profile with iv = 0, :2501-2576 cs_limiters, model/fv_fill.F90:86-153 fillz),
// re-scheduled so that a column never holds more than two full-column work arrays:
//   pass 1 (k = 1..km)      forward sweep of the cubic-spline tridiagonal system; q~(k) and gam(k) to scratch
//   pass 2 (k = km+1..1)    back-substitution; the large-scale interface constraints (fv_mapz.F90:1783-1818) are pointwise
//                           in k, so the constrained interface value is what is written back
//   pass 3 (l = 1..km)      source-layer-major sweep: the limited parabola (a2,a3,a4) of source layer l is generated
//                           from a rolling window and consumed at once by every target layer that overlaps it
//                           (equivalent to the reference's target-major search for monotone pe1); fillz runs one
//                           level behind the emission and the result streams straight to the output buffer.
// One thread owns one column and G tracers in lock-step: the spline matrix (d4, bet, gam), the overlap factors
// (pl, pr, fac1, fac2, esl) and every fp64 division they contain are computed once per G tracers, and G independent
// dependency chains hide the fp64 latency.  Input and output buffers must be distinct (q ping-pong buffers).
#pragma once
#include "fv3t_remap.cuh"

namespace fv3t {

template <class T> struct Remap2Params {
  const T* qsrc;   // (isd:ied, jsd:jed, km, nq) tile-major
  T* qdst;         // same layout, a different buffer
  const T* pe;     // (is-1:ie+1, km+1, js-1:je+1) tile-major
  const T *ak, *bk;
  T* delp;         // (isd:ied, jsd:jed, km) tile-major
  const int* kord; // [nq]
  T ptop;
  int n, km, nq, ntiles, fill;
  int j_first, j_count;
  // Row-granular mapn_tracer entry (fv_mapz.F90:1386-1402): the CALLER's target grid of row j_first, pe2 (i1:i2, km+1) and
  // dp2 (i1:i2, km) on the device; null = pe2 = ak + bk*pe1(km+1) as Lagrangian_to_Eulerian builds it (fv_mapz.F90:263-272)
  const T* pe2_ext = nullptr;
  const T* dp2_ext = nullptr;
};

// limited parabola of an interior layer 3 <= k <= km-2 (fv_mapz.F90:1886-2073), iv = 0, qmin = 0.
// g_m1..g_p2 = a1 differences gam(k-1..k+2); f_m, f_0, f_p = flags (bit0 extm, bit1 ext5, bit2 ext6) of k-1, k, k+1.
template <class T>
FV3T_HD void interior_parabola(int akord, T a1k, T& a2k, T& a3k, T& a4k, T g_m1, T g_0, T g_p1, T g_p2, int f_m, int f_0, int f_p,
                               T qmin) {
  const bool extm = f_0 & 1, ext5 = f_0 & 2, ext6 = f_0 & 4;
  const bool extm_m = f_m & 1, ext5_m = f_m & 2, ext6_m = f_m & 4;
  const bool extm_p = f_p & 1, ext5_p = f_p & 2, ext6_p = f_p & 4;
  auto huynh = [&]() {
    const T pmp_1 = a1k - T(2) * g_p1;
    const T lac_1 = pmp_1 + T(1.5) * g_p2;
    a2k = f_min(f_max(a2k, f_min(a1k, pmp_1, lac_1)), f_max(a1k, pmp_1, lac_1));
    const T pmp_2 = a1k + T(2) * g_0;
    const T lac_2 = pmp_2 - T(1.5) * g_m1;
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
    if ((extm && extm_m) || (extm && extm_p) || (extm && a1k < qmin)) {
      flat();
      a4k = T(0);
    } else {
      set_a6();
      if (f_abs(a4k) > f_abs(a2k - a3k)) {
        huynh();
        set_a6();
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
    if ((ext5 && ext5_m) || (ext5 && ext5_p) || (ext5 && a1k < qmin))
      flat();
    else if (ext6)
      huynh();
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
    if (ext5 && (ext5_m || ext5_p || a1k < qmin)) {
      flat();
      a4k = T(0);
    } else {
      set_a6();
    }
  }
  cs_limiters1<T>(extm, a1k, a2k, a3k, a4k, 0);
}

// flags of layer j (2 <= j <= km-1): extm from the a1 differences, ext5/ext6 from the interface values (fv_mapz.F90:1820-1846)
template <class T> FV3T_HD int layer_flags(int akord, T a1j, T cj, T cj1, T g_j, T g_j1) {
  int f = (g_j * g_j1 < T(0)) ? 1 : 0;
  if (akord > 9) {
    const T x0 = T(2) * a1j - (cj + cj1);
    const T x1 = f_abs(cj - cj1);
    const T a4 = T(3) * x0;
    if (f_abs(x0) > x1) f |= 2;
    if (f_abs(a4) > x1) f |= 4;
  }
  return f;
}

// One column, tracers iq0 .. iq0+G-1 (indices beyond nq-1 are clamped: they recompute the last tracer and store nothing).
template <class T, int G, bool MAPN, int KM>
FV3T_HD void remap_column(const Remap2Params<T>& p, int t, int i, int j, int iq0) {
  const int n = p.n, km = p.km;
  const long nd = n + 6, plane = nd * nd;
  const T r3 = K<T>::r3(), r23 = K<T>::r23();
  const long pe_ld1 = n + 2, pe_ld2 = pe_ld1 * (km + 1);
  const T* pe = p.pe + (long)t * pe_ld2 * (n + 2) + (long)i + (long)j * pe_ld2;  // pe1(k) = pe[(k-1)*pe_ld1]
  const long col = (long)(j + 2) * nd + (i + 2);
  const T* qs[G];
  T* qd[G];
  int akord[G];
  bool live[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int iq = (iq0 + g < p.nq) ? iq0 + g : p.nq - 1;
    live[g] = iq0 + g < p.nq;
    const long off = (((long)t * p.nq + iq) * km) * plane + col;
    qs[g] = p.qsrc + off;
    qd[g] = p.qdst + off;
    const int ko = p.kord[iq];
    akord[g] = ko < 0 ? -ko : ko;
  }
  auto A1 = [&](int g, int k) -> T { return qs[g][(long)(k - 1) * plane]; };
  auto PE1 = [&](int k) -> T { return pe[(long)(k - 1) * pe_ld1]; };
  constexpr int PF = 6;  // prefetch distance in levels: one thread walks its column through HBM with a plane-sized stride
  auto PFA1 = [&](int g, int k) {
    if (k >= 1 && k <= km) prefetch_l1(qs[g] + (long)(k - 1) * plane);
  };
  auto PFPE = [&](int k) {
    if (k <= km + 1) prefetch_l1(pe + (long)(k - 1) * pe_ld1);
  };
  const T ps = PE1(km + 1);
  auto PE2 = [&](int k) -> T {
    if (p.pe2_ext) return p.pe2_ext[(long)(k - 1) * n + (i - 1)];
    return k == 1 ? p.ptop : (k == km + 1 ? ps : p.ak[k - 1] + p.bk[k - 1] * ps);
  };
  // dp2(k) given its two interfaces: the caller's own array when the target grid is the caller's
  auto DP2 = [&](int k, T lo, T hi) -> T { return p.dp2_ext ? p.dp2_ext[(long)(k - 1) * n + (i - 1)] : hi - lo; };

  T gam[KM + 2];
  T qv[G][KM + 2];

  // ---- pass 1: forward sweep (fv_mapz.F90:1736-1750); delp(k) = pe1(k+1) - pe1(k)
  {
    T pa = PE1(1), pb = PE1(2), pc = PE1(3);
    T dpm = pb - pa;  // delp(k-1)
    T dpc = pc - pb;  // delp(k)
    const T grat = dpc / dpm;
    T bet = grat * (grat + T(0.5));
    const T ctop = (grat + grat) * (grat + T(1));
    T a1m[G], a1mm[G], qk[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      a1mm[g] = A1(g, 1);
      a1m[g] = A1(g, 2);
      qk[g] = (ctop * a1mm[g] + a1m[g]) / bet;
      qv[g][1] = qk[g];
    }
    T gprev = (T(1) + grat * (grat + T(1.5))) / bet;
    gam[1] = gprev;
    T d4 = T(0);
    // k = 2: a1(k-1) = a1mm, a1(k) = a1m
    for (int k = 2; k <= km; ++k) {
      PFPE(k + 2 + PF);
#pragma unroll
      for (int g = 0; g < G; ++g) PFA1(g, k + 1 + PF);
      d4 = dpm / dpc;
      bet = T(2) + d4 + d4 - gprev;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        qk[g] = (T(3) * (a1mm[g] + d4 * a1m[g]) - qk[g]) / bet;
        qv[g][k] = qk[g];
      }
      gprev = d4 / bet;
      gam[k] = gprev;
      if (k < km) {
        pb = pc;
        pc = PE1(k + 2);
        dpm = dpc;
        dpc = pc - pb;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          a1mm[g] = a1m[g];
          a1m[g] = A1(g, k + 1);
        }
      }
    }
    // here a1mm = a1(km-1), a1m = a1(km), d4 = delp(km-1)/delp(km), gprev = gam(km)
    const T a_bot = T(1) + d4 * (d4 + T(1.5));
    const T cbot = T(2) * d4 * (d4 + T(1));
    const T den = d4 * (d4 + T(0.5)) - a_bot * gprev;
#pragma unroll
    for (int g = 0; g < G; ++g) qv[g][km + 1] = (cbot * a1m[g] + a1mm[g] - a_bot * qk[g]) / den;
  }

  // ---- pass 2: back-substitution (fv_mapz.F90:1757-1762) + interface constraints (:1783-1818); iv = 0
#pragma unroll
  for (int g = 0; g < G; ++g) {
    T r = qv[g][km + 1];
    if (akord[g] > 16) {
      for (int k = km; k >= 1; --k) {
        r = qv[g][k] - gam[k] * r;
        qv[g][k] = r;
      }
    } else {
      // rolling a1(k+1), a1(k), a1(k-1), a1(k-2); the scratch loads of four levels are issued ahead of the
      // dependent recurrence (local memory lives in L2/HBM: its latency must not sit inside the chain)
      T ap = T(0), a0 = A1(g, km), am = A1(g, km - 1), amm = A1(g, km - 2);
      for (int k0 = km; k0 >= 1; k0 -= 4) {
        T qq[4], gg[4], an[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int k = k0 - u;
          PFA1(g, k - 3 - PF);
          qq[u] = k >= 1 ? qv[g][k] : T(0);
          gg[u] = k >= 1 ? gam[k] : T(0);
          an[u] = (k - 3 >= 1) ? A1(g, k - 3) : T(0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int k = k0 - u;
          if (k >= 1) {
            r = qq[u] - gg[u] * r;
            T c = r;
            if (k == km || k == 2) {
              c = f_min(c, f_max(am, a0));
              c = f_max(c, f_min(am, a0));
            } else if (k >= 3) {
              const T gm = am - amm;  // gam(k-1) = a1(k-1) - a1(k-2)
              const T gp = ap - a0;   // gam(k+1) = a1(k+1) - a1(k)
              if (gm * gp > T(0)) {
                c = f_min(c, f_max(am, a0));
                c = f_max(c, f_min(am, a0));
              } else if (gm > T(0)) {
                c = f_max(c, f_min(am, a0));
              } else {
                c = f_min(c, f_max(am, a0));
                c = f_max(T(0), c);
              }
            }
            qv[g][k] = c;
            ap = a0;
            a0 = am;
            am = amm;
            amm = an[u];
          }
        }
      }
    }
  }

  // ---- pass 3: source-layer-major sweep ------------------------------------------------------------------------------
  // generator state per tracer at source layer l
  T a_0[G], a_p1[G], a_p2[G];            // a1(l), a1(l+1), a1(l+2)
  T c_0[G], c_p1[G], c_p2[G];            // interface values q(l), q(l+1), q(l+2)
  T g_m1[G], g_0[G], g_p1[G], g_p2[G];   // a1 differences gam(l-1..l+2), gam(j) = a1(j) - a1(j-1)
  int f_m[G], f_0[G], f_p[G];            // flags of l-1, l, l+1
  // map / fillz state per tracer
  T qsum[G], xa[G], xb[G], sum0[G], sum1[G];
  bool zfix[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    a_0[g] = A1(g, 1);
    a_p1[g] = A1(g, 2);
    a_p2[g] = A1(g, 3);
    c_0[g] = qv[g][1];
    c_p1[g] = qv[g][2];
    c_p2[g] = qv[g][3];
    g_m1[g] = T(0);
    g_0[g] = T(0);
    g_p1[g] = a_p1[g] - a_0[g];
    g_p2[g] = a_p2[g] - a_p1[g];
    f_m[g] = 0;
    f_0[g] = 0;
    f_p[g] = layer_flags<T>(akord[g], a_p1[g], c_p1[g], c_p2[g], g_p1[g], g_p2[g]);  // flags(2)
    qsum[g] = T(0);
    xa[g] = xb[g] = sum0[g] = sum1[g] = T(0);
    zfix[g] = false;
  }
  // target-layer state (shared by the tracers): pe2(k), pe2(k+1), dp2(k-2..k)
  int k = 1;
  T pe2k = PE2(1), pe2k1 = PE2(2);
  T dpk = DP2(1, pe2k, pe2k1), dpk_m1 = T(0), dpk_m2 = T(0);
  bool started = false;
  T* delp = p.delp + (long)t * plane * km + col;
  const bool wdelp = iq0 == 0 && !p.pe2_ext;  // mapn_tracer itself never writes delp
  if (wdelp) delp[0] = dpk;

  auto finalize = [&](int g, int kk, T x, T dpkk) {
    if (live[g]) qd[g][(long)(kk - 1) * plane] = x;
    if (kk >= 2) {
      const T m = x * dpkk;
      sum0[g] = sum0[g] + m;
      sum1[g] = sum1[g] + f_max(T(0), m);
    }
  };
  // value v[g] of target layer k for every tracer -> fillz pipeline (fv_fill.F90:86-128) or straight to memory
  auto emit = [&](const T (&v)[G]) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      if (!p.fill) {
        if (live[g]) qd[g][(long)(k - 1) * plane] = v[g];
        continue;
      }
      if (k == 1) {
        xa[g] = v[g];
      } else if (k == 2) {
        xb[g] = v[g];
        if (xa[g] < T(0)) {
          xb[g] = xb[g] + xa[g] * dpk_m1 / dpk;
          xa[g] = T(0);
        }
      } else {
        T xc = v[g];
        // interior step j = k-1 on (q(k-2), q(k-1), q(k)) with dp(k-2), dp(k-1), dp(k)
        if (xb[g] < T(0)) {
          zfix[g] = true;
          if (xa[g] > T(0)) {
            const T dq = f_min(xa[g] * dpk_m2, -xb[g] * dpk_m1);
            xa[g] = xa[g] - dq / dpk_m2;
            xb[g] = xb[g] + dq / dpk_m1;
          }
          if (xb[g] < T(0) && xc > T(0)) {
            const T dq = f_min(xc * dpk, -xb[g] * dpk_m1);
            xc = xc - dq / dpk;
            xb[g] = xb[g] + dq / dpk_m1;
          }
        }
        finalize(g, k - 2, xa[g], dpk_m2);
        xa[g] = xb[g];
        xb[g] = xc;
        if (k == km) {
          if (xb[g] < T(0) && xa[g] > T(0)) {
            zfix[g] = true;
            const T qup = xa[g] * dpk_m1;
            const T qly = -xb[g] * dpk;
            const T dup = f_min(qly, qup);
            xa[g] = xa[g] - dup / dpk_m1;
            xb[g] = xb[g] + dup / dpk;
          }
          finalize(g, km - 1, xa[g], dpk_m1);
          finalize(g, km, xb[g], dpk);
        }
      }
    }
    // next target layer
    ++k;
    if (k <= km) {
      pe2k = pe2k1;
      pe2k1 = PE2(k + 1);
      dpk_m2 = dpk_m1;
      dpk_m1 = dpk;
      dpk = DP2(k, pe2k, pe2k1);
      if (wdelp) delp[(long)(k - 1) * plane] = dpk;
    }
  };

  T pe1lo = PE1(1), pe1hi = PE1(2);
  // l > km: no source layer left (malformed pe only) -- the reference then divides the stale qsum for the remaining targets
  for (int l = 1; k <= km; ++l) {
    const bool have = l <= km;
    const T dp1l = pe1hi - pe1lo;
    PFPE(l + 2 + PF);
#pragma unroll
    for (int g = 0; g < G; ++g) PFA1(g, l + 3 + PF);
    // ---- limited parabola of source layer l for every tracer
    T a2[G], a3[G], a4[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      a2[g] = c_0[g];
      a3[g] = c_p1[g];
      a4[g] = T(0);
      if (!have) {
      } else if (akord[g] > 16) {
        a4[g] = T(3) * (T(2) * a_0[g] - (a2[g] + a3[g]));
      } else if (l >= 3 && l <= km - 2) {
        interior_parabola<T>(akord[g], a_0[g], a2[g], a3[g], a4[g], g_m1[g], g_0[g], g_p1[g], g_p2[g], f_m[g], f_0[g], f_p[g], T(0));
      } else {
        if (l == 1) a2[g] = f_max(T(0), a2[g]);
        if (l == km) a3[g] = f_max(T(0), a3[g]);
        a4[g] = T(3) * (T(2) * a_0[g] - (a2[g] + a3[g]));
        cs_limiters1<T>(f_0[g] & 1, a_0[g], a2[g], a3[g], a4[g], (l == 1 || l == km) ? 1 : 2);
      }
    }
    // ---- consume every target layer that ends in (or passes through) source layer l
    while (k <= km) {
      T v[G];
      bool done = false;  // true: target layer k is complete, v holds its values
      if (!have) {
#pragma unroll
        for (int g = 0; g < G; ++g) v[g] = qsum[g] / dpk;
        done = true;
      } else if (!started) {
        if (pe2k > pe1hi) {
          // top edge of the target lies below this source layer: next l
        } else if (pe2k < pe1lo) {
          // not locatable (cannot happen when pe1(1) = pe2(1)): the reference leaves qsum as it is
#pragma unroll
          for (int g = 0; g < G; ++g) v[g] = qsum[g] / dpk;
          done = true;
        } else {
          const T pl = (pe2k - pe1lo) / dp1l;
          if (pe2k1 <= pe1hi) {
            const T pr = (pe2k1 - pe1lo) / dp1l;
            if (MAPN) {
              T fac1 = pr + pl;
              const T fac2 = r3 * (pr * fac1 + pl * pl);
              fac1 = T(0.5) * fac1;
#pragma unroll
              for (int g = 0; g < G; ++g) v[g] = a2[g] + (a4[g] + a3[g] - a2[g]) * fac1 - a4[g] * fac2;
            } else {
#pragma unroll
              for (int g = 0; g < G; ++g)
                v[g] = a2[g] + T(0.5) * (a4[g] + a3[g] - a2[g]) * (pr + pl) - a4[g] * r3 * (pr * (pr + pl) + pl * pl);
            }
            done = true;
          } else {
            const T dp = pe1hi - pe2k;
            if (MAPN) {
              T fac1 = T(1) + pl;
              const T fac2 = r3 * (T(1) + pl * fac1);
              fac1 = T(0.5) * fac1;
#pragma unroll
              for (int g = 0; g < G; ++g) qsum[g] = dp * (a2[g] + (a4[g] + a3[g] - a2[g]) * fac1 - a4[g] * fac2);
            } else {
#pragma unroll
              for (int g = 0; g < G; ++g)
                qsum[g] = dp * (a2[g] + T(0.5) * (a4[g] + a3[g] - a2[g]) * (T(1) + pl) - a4[g] * (r3 * (T(1) + pl * (T(1) + pl))));
            }
            started = true;
          }
        }
      } else if (pe2k1 > pe1hi) {  // whole layer
#pragma unroll
        for (int g = 0; g < G; ++g) qsum[g] = qsum[g] + dp1l * a_0[g];
      } else {
        const T dp = pe2k1 - pe1lo;
        const T esl = dp / dp1l;
        if (MAPN) {
          const T fac1 = T(0.5) * esl;
          const T fac2 = T(1) - r23 * esl;
#pragma unroll
          for (int g = 0; g < G; ++g) {
            qsum[g] = qsum[g] + dp * (a2[g] + fac1 * (a3[g] - a2[g] + a4[g] * fac2));
            v[g] = qsum[g] / dpk;
          }
        } else {
#pragma unroll
          for (int g = 0; g < G; ++g) {
            qsum[g] = qsum[g] + dp * (a2[g] + T(0.5) * esl * (a3[g] - a2[g] + a4[g] * (T(1) - r23 * esl)));
            v[g] = qsum[g] / dpk;
          }
        }
        started = false;
        done = true;
      }
      if (!done) break;
      emit(v);
    }
    // ---- advance the generators to source layer l+1
    if (l < km) {
      pe1lo = pe1hi;
      pe1hi = PE1(l + 2);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        a_0[g] = a_p1[g];
        a_p1[g] = a_p2[g];
        a_p2[g] = (l + 3 <= km) ? A1(g, l + 3) : T(0);
        c_0[g] = c_p1[g];
        c_p1[g] = c_p2[g];
        c_p2[g] = (l + 3 <= km + 1) ? qv[g][l + 3] : T(0);
        g_m1[g] = g_0[g];
        g_0[g] = g_p1[g];
        g_p1[g] = g_p2[g];
        g_p2[g] = a_p2[g] - a_p1[g];  // gam(l+3) = a1(l+3) - a1(l+2)
        f_m[g] = f_0[g];
        f_0[g] = f_p[g];
        // flags(l+2), needed while 2 <= l+2 <= km-1
        f_p[g] = (l + 2 <= km - 1) ? layer_flags<T>(akord[g], a_p1[g], c_p1[g], c_p2[g], g_p1[g], g_p2[g]) : 0;
      }
    }
  }
  // ---- fillz non-local rescale for the flagged columns (fv_fill.F90:131-152); re-reads this thread's own output
  if (p.fill) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      if (zfix[g] && sum0[g] > T(0) && live[g]) {
        const T fac = sum0[g] / sum1[g];
        for (int kk = 2; kk <= km; ++kk) {
          const T dp = DP2(kk, PE2(kk), PE2(kk + 1));
          const T x = qd[g][(long)(kk - 1) * plane];
          qd[g][(long)(kk - 1) * plane] = f_max(T(0), fac * (x * dp) / dp);
        }
      }
    }
  }
}

template <class T, int G, bool MAPN, int KM> __global__ void __launch_bounds__(128) k_remap2(const Remap2Params<T> p) {
  const int cols = p.n * p.j_count;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const int i = c % p.n + 1, j = c / p.n + 1 + p.j_first;
  remap_column<T, G, MAPN, KM>(p, blockIdx.y, i, j, blockIdx.z * G);
}

}  // namespace fv3t
