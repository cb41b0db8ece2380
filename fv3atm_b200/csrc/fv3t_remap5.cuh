// fv3atm_b200: two-pass form of the production vertical tracer remap (mapn_tracer with scalar_profile iv = 0, cs_limiters, fillz:
// atmos_cubed_sphere/model/fv_mapz.F90:1386-1499, 1691-2096, 2501-2576, model/fv_fill.F90:86-153).
//
// k_remap3 walks a column three times (forward elimination top-down, back-substitution bottom-up, then everything else
// top-down) because the reference eliminates the spline system from the top (:1736-1762).  The system
//      row 1      bet1 x1 + c1 x2                        = ctop a1 + a2
//      row k      x(k-1) + (2 + 2 d4_k) x_k + d4_k x(k+1) = 3 (a(k-1) + d4_k a_k)          k = 2..km,  d4_k = dp(k-1) / dp(k)
//      row km+1   a_bot x_km + D x(km+1)                  = cbot a_km + a(km-1)
// is diagonally dominant, so it can just as well be eliminated FROM THE BOTTOM:  x(k) = e_k - f_k x(k-1)  with
//      f(km+1) = a_bot / D,  e(km+1) = rhs(km+1) / D,     bet'_k = 2 + 2 d4_k - d4_k f(k+1),  f_k = 1 / bet'_k,
//      e_k = (rhs_k - d4_k e(k+1)) f_k,                    x1 = (rhs_1 - c1 e2) / (bet1 - c1 f2)
// and then the interface values come out TOP-DOWN -- the direction in which the constraints, the limited parabolas, the overlap
// integrals and fillz consume them.  Two walks instead of three:
//   pass A (bottom-up)  e_k into the per-thread scratch column
//   pass B (top-down)   x_k = e_k - f_k x(k-1), interface constraints (:1783-1818), limited parabola, overlap integration, fillz
// q is read twice instead of three times and the scratch column is written once and read once instead of twice each.  Same
// equations, another elimination order: the interface values differ from the reference's by rounding only (normalised 1e-16;
// the tests hold the 1e-12 bar), like every other "fast" kernel.  f, d4, 1/bet', the closures, 1/dp1, 1/dp2 and pe2 are tracer
// independent (k_remap_coef5, once per column); pe2(k+1) travels with 1/dp2(k) through the cp.async ring, so the target-layer
// loop evaluates no ak + bk*ps, and the fillz sums are taken only in the (rare) columns that need the non-local fix.
//
// MEASURED (B200, C768 L127 x9 fp64): 84.9 ms against 76.6 ms for k_remap3 -- parity-green but slower, so it is opt-in
// (FV3T_REMAP5=1).  The remap is latency-bound at 16 warps per SM (128 registers per thread); the fused walk puts the recurrence,
// the constraint and the flags of the NEXT interface on every source-layer iteration with one layer of look-ahead, whereas
// k_remap3's separate back-substitution walk issues eight levels of independent loads at a time.  Fewer instructions and fewer
// bytes, longer dependent chains per iteration: a loss.
#pragma once
#include <type_traits>

#include "fv3t_remap3.cuh"

namespace fv3t {

// Remap3Params is reused: P1 = {d4_k, 1/bet'_k} (level 1: {ctop, c1}; level km+1: {cbot, 1/D}), GAM = f_k (level 1: 1/(bet1 - c1 f2)),
// RD1 = 1/dp1, R2 (2 * km planes of T) = Pair{1/dp2(k), pe2(k+1)}

template <class T> FV3T_HD void remap_coef5_column(const Remap3Params<T>& p, int t, int i, int j) {
  const int n = p.n, km = p.km;
  const long nd = n + 6, plane = nd * nd;
  const long pe_ld1 = n + 2, pe_ld2 = pe_ld1 * (km + 1);
  const T* pe = p.pe + (long)t * pe_ld2 * (n + 2) + (long)i + (long)j * pe_ld2;
  const long col = (long)(j + 2) * nd + (i + 2);
  const int ipl = (int)plane, ipe = (int)pe_ld1;
  Pair<T>* P1 = p.P1 + (long)t * plane * (km + 1) + col;
  T* F = p.GAM + (long)t * plane * (km + 1) + col;
  T* RD1 = p.RD1 + (long)t * plane * km + col;
  Pair<T>* R2 = reinterpret_cast<Pair<T>*>(p.R2) + (long)t * plane * km + col;
  T* delp = p.delp + (long)t * plane * km + col;
  auto PE1 = [&](int k) -> T { return pe[(k - 1) * ipe]; };
  const T ps = PE1(km + 1);
  auto PE2 = [&](int k) -> T { return k == 1 ? p.ptop : (k == km + 1 ? ps : add_rn(p.ak[k - 1], mul_rn(p.bk[k - 1], ps))); };  // uncontracted: delp is caller-visible
  // bottom-up: dp(k), d4_k = dp(k-1)/dp(k), f, 1/bet'
  T plo = PE1(km), phi = ps;
  T dpk = phi - plo;              // dp(km)
  T pm = PE1(km - 1);
  T dpm = plo - pm;               // dp(km-1)
  T d4 = dpm / dpk;
  {
    const T a_bot = T(1) + d4 * (d4 + T(1.5));
    const T rD = T(1) / (d4 * (d4 + T(0.5)));
    P1[km * ipl] = Pair<T>{T(2) * d4 * (d4 + T(1)), rD};
    F[km * ipl] = a_bot * rD;
  }
  T f = F[km * ipl];
  RD1[(km - 1) * ipl] = T(1) / dpk;
  for (int k = km; k >= 2; --k) {
    // here dpk = dp(k), dpm = dp(k-1), d4 = d4_k, f = f(k+1)
    const T rb = T(1) / (T(2) + d4 + d4 - d4 * f);
    P1[(k - 1) * ipl] = Pair<T>{d4, rb};
    F[(k - 1) * ipl] = rb;
    f = rb;
    RD1[(k - 2) * ipl] = T(1) / dpm;
    if (k > 2) {
      dpk = dpm;
      const T pmm = PE1(k - 2);
      dpm = pm - pmm;
      pm = pmm;
      d4 = dpm / dpk;
    }
  }
  {  // top closure: here dpm = dp(1), dpk = dp(2) (km >= 3), f = f_2
    const T grat = dpk / dpm;
    const T bet1 = grat * (grat + T(0.5));
    const T c1 = T(1) + grat * (grat + T(1.5));
    P1[0] = Pair<T>{(grat + grat) * (grat + T(1)), c1};
    F[0] = T(1) / (bet1 - c1 * f);
  }
  T p2a = PE2(1);
  for (int k = 1; k <= km; ++k) {
    const T p2b = PE2(k + 1);
    const T dp2 = p2b - p2a;
    delp[(k - 1) * ipl] = dp2;
    R2[(k - 1) * ipl] = Pair<T>{T(1) / dp2, p2b};
    p2a = p2b;
  }
}

// ---- one column of one tracer; AK = abs(kord); ring: four Pair<T> slots, `rstride` Pairs apart ---------------------------
template <class T, int AK, int KM>
FV3T_HD void remap5_column(const Remap3Params<T>& p, Pair<T>* ring, int rstride, int t, int i, int j, int iq) {
  const int n = p.n, km = p.km;
  const long nd = n + 6, plane = nd * nd;
  const T r3 = K<T>::r3(), r23 = K<T>::r23();
  const long pe_ld1 = n + 2, pe_ld2 = pe_ld1 * (km + 1);
  const T* pe = p.pe + (long)t * pe_ld2 * (n + 2) + (long)i + (long)j * pe_ld2;  // pe1(k) = pe[(k-1)*pe_ld1]
  const long col = (long)(j + 2) * nd + (i + 2);
  const long off = (((long)t * p.nq + iq) * km) * plane + col;
  const int ipl = (int)plane, ipe = (int)pe_ld1;  // 32-bit offsets inside a column (checked by the launcher)
  const T* __restrict__ qs = p.qsrc + off;
  T* __restrict__ qd = p.qdst + off;
  const Pair<T>* P1 = p.P1 + (long)t * plane * (km + 1) + col;
  const T* F = p.GAM + (long)t * plane * (km + 1) + col;
  const T* RD1 = p.RD1 + (long)t * plane * km + col;
  const Pair<T>* R2 = reinterpret_cast<const Pair<T>*>(p.R2) + (long)t * plane * km + col;
  auto A1 = [&](int k) -> T { return qs[(k - 1) * ipl]; };
  auto PE1 = [&](int k) -> T { return pe[(k - 1) * ipe]; };

  T qv[KM + 2];  // e_k, k = 2..km+1
  constexpr int CH = 8;

  // ---- pass A (bottom-up): e_k.  Loads are issued a chunk of CH levels ahead of the dependent recurrence
  {
    T a0 = A1(km), am = A1(km - 1);
    const Pair<T> cb = P1[km * ipl];
    T e = (cb.a * a0 + am) * cb.b;
    qv[km + 1] = e;
    auto chunk = [&](auto full, int k0) {  // levels k0, k0-1, .., k0-CH+1; full: all of them >= 3
      constexpr bool FULL = decltype(full)::value;
      T an[CH];
      Pair<T> ck[CH];
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        const int k = k0 - u;
        ck[u] = (FULL || k >= 2) ? P1[(k - 1) * ipl] : Pair<T>{T(0), T(0)};
        an[u] = (FULL || k - 2 >= 1) ? A1(k - 2) : T(0);
      }
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        const int k = k0 - u;
        if (FULL || k >= 2) {
          e = (T(3) * (am + ck[u].a * a0) - ck[u].a * e) * ck[u].b;
          qv[k] = e;
          a0 = am;
          am = an[u];
        }
      }
    };
    int k0 = km;
    for (; k0 - CH + 1 >= 3; k0 -= CH) chunk(std::true_type{}, k0);
    for (; k0 >= 2; k0 -= CH) chunk(std::false_type{}, k0);
  }

  // interface value k from its raw spline value x: fv_mapz.F90:1783-1818 (iv = 0); amm, am, a0, ap = a1(k-2 .. k+1)
  auto constrain = [&](int kk, T x, T amm, T am, T a0, T ap) -> T {
    if (AK > 16) return x;
    T c = x;
    if (kk == km || kk == 2) {
      c = f_min(c, f_max(am, a0));
      c = f_max(c, f_min(am, a0));
    } else if (kk >= 3 && kk < km) {
      const T gm = am - amm;  // gam(k-1)
      const T gp = ap - a0;   // gam(k+1)
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
    return c;
  };

  // ---- pass B (top-down), source-layer major (see fv3t_remap2.cuh)
  T a_0 = A1(1), a_p1 = A1(2), a_p2 = A1(3), a_p3 = km >= 4 ? A1(4) : T(0);
  T xr, c_0, c_p1, c_p2;
  {
    const Pair<T> ct = P1[0];
    const T x1 = (ct.a * a_0 + a_p1 - ct.b * qv[2]) * F[0];
    const T x2 = qv[2] - F[ipl] * x1;
    xr = qv[3] - F[2 * ipl] * x2;
    c_0 = x1;
    c_p1 = constrain(2, x2, T(0), a_0, a_p1, a_p2);
    c_p2 = constrain(3, xr, a_0, a_p1, a_p2, a_p3);
  }
  T g_m1 = T(0), g_0 = T(0), g_p1 = a_p1 - a_0, g_p2 = a_p2 - a_p1;
  int f_m = 0, f_0 = 0, f_p = layer_flags<T>(AK, a_p1, c_p1, c_p2, g_p1, g_p2);
  T qsum = T(0), xa = T(0), xb = T(0);
  bool zfix = false;
  int k = 1;
  // {1/dp2(k), pe2(k+1)} of the target layers reaches the thread through a four-deep cp.async ring (slot = k mod 4): requested
  // three target layers ahead, completed with cp.async.wait_group
  auto r2_request = [&](int kk) {  // target layer kk (1-based), clamped
    const int kc = kk <= km ? kk : km;
    async_copy<sizeof(Pair<T>)>(ring + (kk & 3) * rstride, R2 + (kc - 1) * ipl);
    async_commit();
  };
  r2_request(1);
  r2_request(2);
  r2_request(3);
  async_wait_pending<2>();
  T pe2k = p.ptop;
  T rdpk = ring[(1 & 3) * rstride].a, pe2k1 = ring[(1 & 3) * rstride].b;
  T dpk = pe2k1 - pe2k, dpk_m1 = T(0), dpk_m2 = T(0);
  bool started = false;

  // value v of target layer k -> fillz pipeline (fv_fill.F90:86-128) or straight to memory
  auto emit = [&](T v) {
    if (!p.fill) {
      qd[(k - 1) * ipl] = v;
    } else if (k == 1) {
      xa = v;
    } else if (k == 2) {
      xb = v;
      if (xa < T(0)) {
        xb = xb + xa * dpk_m1 / dpk;
        xa = T(0);
      }
    } else {
      T xc = v;
      if (xb < T(0)) {
        zfix = true;
        if (xa > T(0)) {
          const T dq = f_min(xa * dpk_m2, -xb * dpk_m1);
          xa = xa - dq / dpk_m2;
          xb = xb + dq / dpk_m1;
        }
        if (xb < T(0) && xc > T(0)) {
          const T dq = f_min(xc * dpk, -xb * dpk_m1);
          xc = xc - dq / dpk;
          xb = xb + dq / dpk_m1;
        }
      }
      qd[(k - 3) * ipl] = xa;
      xa = xb;
      xb = xc;
      if (k == km) {
        if (xb < T(0) && xa > T(0)) {
          zfix = true;
          const T qup = xa * dpk_m1;
          const T qly = -xb * dpk;
          const T dup = f_min(qly, qup);
          xa = xa - dup / dpk_m1;
          xb = xb + dup / dpk;
        }
        qd[(km - 2) * ipl] = xa;
        qd[(km - 1) * ipl] = xb;
      }
    }
    ++k;
    if (k <= km) {
      pe2k = pe2k1;
      dpk_m2 = dpk_m1;
      dpk_m1 = dpk;
      r2_request(k + 2);
      async_wait_pending<2>();  // at most the requests for k+1, k+2 are still in flight: slot k has landed
      const Pair<T> r = ring[(k & 3) * rstride];
      rdpk = r.a;
      pe2k1 = r.b;
      dpk = pe2k1 - pe2k;
    }
  };

  T pe1lo = PE1(1), pe1hi = PE1(2);
  T rdp1 = RD1[0];
  for (int l = 1; k <= km; ++l) {
    const bool have = l <= km;
    const T dp1l = pe1hi - pe1lo;
    // inputs of the NEXT source layer, requested before this layer's arithmetic so that their latency hides behind it
    T n_pe = T(0), n_rd = T(0), n_a = T(0), n_e = T(0), n_f = T(0);
    if (l < km) {
      n_pe = PE1(l + 2);
      n_rd = RD1[l * ipl];
      n_a = (l + 4 <= km) ? A1(l + 4) : T(0);
      if (l + 3 <= km + 1) {
        n_e = qv[l + 3];
        n_f = F[(l + 2) * ipl];
      }
    }
    // ---- limited parabola of source layer l
    T a2 = c_0, a3 = c_p1, a4 = T(0);
    if (!have) {
    } else if (AK > 16) {
      a4 = T(3) * (T(2) * a_0 - (a2 + a3));
    } else if (l >= 3 && l <= km - 2) {
      interior_parabola<T>(AK, a_0, a2, a3, a4, g_m1, g_0, g_p1, g_p2, f_m, f_0, f_p, T(0));
    } else {
      if (l == 1) a2 = f_max(T(0), a2);
      if (l == km) a3 = f_max(T(0), a3);
      a4 = T(3) * (T(2) * a_0 - (a2 + a3));
      cs_limiters1<T>(f_0 & 1, a_0, a2, a3, a4, (l == 1 || l == km) ? 1 : 2);
    }
    // ---- consume every target layer that ends in (or passes through) source layer l
    while (k <= km) {
      T v = T(0);
      bool done = false;
      if (!have) {
        v = qsum * rdpk;
        done = true;
      } else if (!started) {
        if (pe2k > pe1hi) {
        } else if (pe2k < pe1lo) {
          v = qsum * rdpk;
          done = true;
        } else {
          const T pl = (pe2k - pe1lo) * rdp1;
          if (pe2k1 <= pe1hi) {
            const T pr = (pe2k1 - pe1lo) * rdp1;
            T fac1 = pr + pl;
            const T fac2 = r3 * (pr * fac1 + pl * pl);
            fac1 = T(0.5) * fac1;
            v = a2 + (a4 + a3 - a2) * fac1 - a4 * fac2;
            done = true;
          } else {
            const T dp = pe1hi - pe2k;
            T fac1 = T(1) + pl;
            const T fac2 = r3 * (T(1) + pl * fac1);
            fac1 = T(0.5) * fac1;
            qsum = dp * (a2 + (a4 + a3 - a2) * fac1 - a4 * fac2);
            started = true;
          }
        }
      } else if (pe2k1 > pe1hi) {  // whole layer
        qsum = qsum + dp1l * a_0;
      } else {
        const T dp = pe2k1 - pe1lo;
        const T esl = dp * rdp1;
        const T fac1 = T(0.5) * esl;
        const T fac2 = T(1) - r23 * esl;
        qsum = qsum + dp * (a2 + fac1 * (a3 - a2 + a4 * fac2));
        v = qsum * rdpk;
        started = false;
        done = true;
      }
      if (!done) break;
      emit(v);
    }
    // ---- advance the generators to source layer l+1: the new interface value c(l+3) from x(l+3) = e(l+3) - f(l+3) x(l+2)
    if (l < km) {
      T c_n = T(0);
      if (l + 3 <= km + 1) {
        xr = n_e - n_f * xr;
        c_n = constrain(l + 3, xr, a_p1, a_p2, a_p3, n_a);
      }
      pe1lo = pe1hi;
      pe1hi = n_pe;
      rdp1 = n_rd;
      a_0 = a_p1;
      a_p1 = a_p2;
      a_p2 = a_p3;
      a_p3 = n_a;
      c_0 = c_p1;
      c_p1 = c_p2;
      c_p2 = c_n;
      g_m1 = g_0;
      g_0 = g_p1;
      g_p1 = g_p2;
      g_p2 = a_p2 - a_p1;
      f_m = f_0;
      f_0 = f_p;
      f_p = (l + 2 <= km - 1) ? layer_flags<T>(AK, a_p1, c_p1, c_p2, g_p1, g_p2) : 0;
    }
  }
  // ---- fillz non-local rescale for the flagged columns (fv_fill.F90:131-152): re-reads this thread's own output; the two
  //      sums are taken here, in the same order as the reference takes them, instead of in every column
  if (p.fill && zfix) {
    const T* dp2 = p.delp + (long)t * plane * km + col;  // written by k_remap_coef5
    T sum0 = T(0), sum1 = T(0);
    for (int kk = 2; kk <= km; ++kk) {
      const T m = qd[(kk - 1) * ipl] * dp2[(kk - 1) * ipl];
      sum0 = sum0 + m;
      sum1 = sum1 + f_max(T(0), m);
    }
    if (sum0 > T(0)) {
      const T fac = sum0 / sum1;
      for (int kk = 2; kk <= km; ++kk) {
        const T dp = dp2[(kk - 1) * ipl];
        const T x = qd[(kk - 1) * ipl];
        qd[(kk - 1) * ipl] = f_max(T(0), fac * (x * dp) / dp);
      }
    }
  }
}

#ifdef __CUDACC__
template <class T> __global__ void __launch_bounds__(128) k_remap_coef5(const Remap3Params<T> p) {
  const int cols = p.n * p.n;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  remap_coef5_column<T>(p, blockIdx.y, c % p.n + 1, c / p.n + 1);
}
template <class T, int AK, int KM, int MINB> __global__ void __launch_bounds__(128, MINB) k_remap5(const Remap3Params<T> p) {
  __shared__ __align__(16) Pair<T> s_ring[4 * 128];
  const int cols = p.n * p.n;
  // blockIdx.x = tracer: the nq CTAs of one column block are adjacent in the grid, run together and share the coefficients and pe
  // through L2
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  remap5_column<T, AK, KM>(p, s_ring + threadIdx.x, 128, blockIdx.z, c % p.n + 1, c / p.n + 1, p.iq0 + blockIdx.x);
}
#endif

}  // namespace fv3t
