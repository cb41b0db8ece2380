// fv3atm_b200: production ("fast") vertical tracer remap for sm_100a.
//
// Same algorithm and the same three-pass streaming schedule as fv3t_remap2.cuh (mapn_tracer / map1_q2 with
// scalar_profile iv = 0, cs_limiters, fillz: atmos_cubed_sphere/model/fv_mapz.F90:1386-1592, 1691-2096, 2501-2576,
// model/fv_fill.F90:86-153), re-balanced like the fast advection kernel (the fp64 remap is bound by instruction issue and
// by the latency of dependent FP64 chains, not by HBM -- DESIGN.md "Roofline"):
//   k_remap_coef3 (one thread per column, ONCE for all tracers): the cubic-spline matrix of scalar_profile depends only on
//       delp (fv_mapz.F90:1736-1750; the reference recomputes it per tracer because scalar_profile sits inside the iq loop,
//       :1418-1425): d4, 1/bet, gam per level, the top / bottom closure coefficients, 1/dp1 for the overlap weights
//       (:1434-1471), pe2 -> delp and 1/dp2 (:263-272, 1481).  Five fp64 divisions per level leave the per-tracer kernel.
//   k_remap3 (one thread per column and tracer, abs(kord) a template parameter): forward sweep, back-substitution +
//       interface constraints, limited parabola + overlap integration + fillz, all divisions by the shared quantities
//       replaced by multiplications with the stored reciprocals, FMA contraction on, ONE per-thread scratch column
//       (the strict kernel keeps two), no software prefetch instructions.
// Results agree with the FMA-free oracle to ~1e-15 normalised on smooth data (the 1e-12 bar is asserted in the tests); the
// strict kernel (fv3t_remap2.cuh, FV3T_STRICT=1) stays bit-identical.  Tracer sets with mixed kord use the strict kernel.
#pragma once
#include <type_traits>

#include "fv3t_advect4.cuh"
#include "fv3t_remap2.cuh"

namespace fv3t {

template <class T> struct Remap3Params {
  const T* qsrc;   // (isd:ied, jsd:jed, km, nq) tile-major
  T* qdst;         // same layout, a different buffer
  const T* pe;     // (is-1:ie+1, km+1, js-1:je+1) tile-major
  const T *ak, *bk;
  T* delp;         // (isd:ied, jsd:jed, km) tile-major
  Pair<T>* P1;     // (isd:ied, jsd:jed, km+1): {d4, 1/bet}; level 1 {ctop, 1/bet1}; level km+1 {cbot, a_bot}
  T* GAM;          // (isd:ied, jsd:jed, km+1): gam;   level km+1: 1/den
  T* RD1;          // (isd:ied, jsd:jed, km):   1/dp1
  T* R2;           // storage of 2 x (isd:ied, jsd:jed, km), read as Pair: {1/dp2(k), pe2(k+1)}
  T ptop;
  int n, km, nq, ntiles, fill;
  int iq0 = 0, nql = -1;  // this launch remaps tracers iq0 .. iq0+nql-1 of the nq resident ones (nql < 0: all)
};

// ---- tracer-independent column coefficients ---------------------------------------------------------------------------
template <class T> FV3T_HD void remap_coef_column(const Remap3Params<T>& p, int t, int i, int j) {
  const int n = p.n, km = p.km;
  const long nd = n + 6, plane = nd * nd;
  const long pe_ld1 = n + 2, pe_ld2 = pe_ld1 * (km + 1);
  const T* pe = p.pe + (long)t * pe_ld2 * (n + 2) + (long)i + (long)j * pe_ld2;
  const long col = (long)(j + 2) * nd + (i + 2);
  Pair<T>* P1 = p.P1 + (long)t * plane * (km + 1) + col;
  T* GAM = p.GAM + (long)t * plane * (km + 1) + col;
  T* RD1 = p.RD1 + (long)t * plane * km + col;
  Pair<T>* R2 = reinterpret_cast<Pair<T>*>(p.R2) + (long)t * plane * km + col;
  T* delp = p.delp + (long)t * plane * km + col;
  const int ipl = (int)plane, ipe = (int)pe_ld1;  // 32-bit offsets inside a column (see remap3_column)
  auto PE1 = [&](int k) -> T { return pe[(k - 1) * ipe]; };
  const T ps = PE1(km + 1);
  auto PE2 = [&](int k) -> T { return k == 1 ? p.ptop : (k == km + 1 ? ps : add_rn(p.ak[k - 1], mul_rn(p.bk[k - 1], ps))); };  // uncontracted: delp is caller-visible
  T pa = PE1(1), pb = PE1(2), pc = PE1(3);
  T dpm = pb - pa, dpc = pc - pb;
  const T grat = dpc / dpm;
  T bet = grat * (grat + T(0.5));
  P1[0] = Pair<T>{(grat + grat) * (grat + T(1)), T(1) / bet};
  T gprev = (T(1) + grat * (grat + T(1.5))) / bet;
  GAM[0] = gprev;
  RD1[0] = T(1) / dpm;
  T d4 = T(0);
  for (int k = 2; k <= km; ++k) {
    d4 = dpm / dpc;
    bet = T(2) + d4 + d4 - gprev;
    gprev = d4 / bet;
    P1[(k - 1) * ipl] = Pair<T>{d4, T(1) / bet};
    GAM[(k - 1) * ipl] = gprev;
    RD1[(k - 1) * ipl] = T(1) / dpc;
    if (k < km) {
      pb = pc;
      pc = PE1(k + 2);
      dpm = dpc;
      dpc = pc - pb;
    }
  }
  const T a_bot = T(1) + d4 * (d4 + T(1.5));
  const T cbot = T(2) * d4 * (d4 + T(1));
  const T den = d4 * (d4 + T(0.5)) - a_bot * gprev;
  P1[km * ipl] = Pair<T>{cbot, a_bot};
  GAM[km * ipl] = T(1) / den;
  T p2a = PE2(1);
  for (int k = 1; k <= km; ++k) {
    const T p2b = PE2(k + 1);
    const T dp2 = p2b - p2a;
    delp[(k - 1) * ipl] = dp2;
    R2[(k - 1) * ipl] = Pair<T>{T(1) / dp2, p2b};
    p2a = p2b;
  }
}

// ---- one column of one tracer; AK = abs(kord) -----------------------------------------------------------------------------
template <class T, int AK, bool MAPN, int KM>
FV3T_HD void remap3_column(const Remap3Params<T>& p, Pair<T>* ring, int rstride, int t, int i, int j, int iq) {
  const int n = p.n, km = p.km;
  const long nd = n + 6, plane = nd * nd;
  const T r3 = K<T>::r3(), r23 = K<T>::r23();
  const long pe_ld1 = n + 2, pe_ld2 = pe_ld1 * (km + 1);
  const T* pe = p.pe + (long)t * pe_ld2 * (n + 2) + (long)i + (long)j * pe_ld2;  // pe1(k) = pe[(k-1)*pe_ld1]
  const long col = (long)(j + 2) * nd + (i + 2);
  const long off = (((long)t * p.nq + iq) * km) * plane + col;
  // offsets inside a column fit 32 bits ((km + 1) * plane < 2^31, checked by the launcher): one IMAD.WIDE per access instead of
  // a 64 x 64-bit multiply
  const int ipl = (int)plane, ipe = (int)pe_ld1;
  const T* __restrict__ qs = p.qsrc + off;
  T* __restrict__ qd = p.qdst + off;
  const Pair<T>* P1 = p.P1 + (long)t * plane * (km + 1) + col;
  const T* GAM = p.GAM + (long)t * plane * (km + 1) + col;
  const T* RD1 = p.RD1 + (long)t * plane * km + col;
  const Pair<T>* R2 = reinterpret_cast<const Pair<T>*>(p.R2) + (long)t * plane * km + col;
  auto A1 = [&](int k) -> T { return qs[(k - 1) * ipl]; };
  auto PE1 = [&](int k) -> T { return pe[(k - 1) * ipe]; };

  T qv[KM + 2];

  // ---- pass 1: forward sweep (fv_mapz.F90:1736-1750) with the stored d4, 1/bet.  Loads are issued a chunk of CH levels
  //      ahead of the dependent recurrence: the column walks through HBM with a plane-sized stride, and nothing but
  //      memory-level parallelism hides that latency (profiles/r01_remap3_v0_ncu.txt: 8 of 12 stall cycles per issue)
  constexpr int CH = 8;
  // chunks whose levels are all interior run without the per-level bound tests (`full`), the first / last ones with them
  {
    T a1mm = A1(1), a1m = A1(2);
    const Pair<T> c1 = P1[0];
    T qk = (c1.a * a1mm + a1m) * c1.b;
    qv[1] = qk;
    auto chunk = [&](auto full, int k0) {
      constexpr bool FULL = decltype(full)::value;
      T an[CH];
      Pair<T> ck[CH];
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        const int k = k0 + u;
        ck[u] = (FULL || k <= km) ? P1[(k - 1) * ipl] : Pair<T>{T(0), T(0)};
        an[u] = (FULL || k + 1 <= km) ? A1(k + 1) : T(0);
      }
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        const int k = k0 + u;
        if (FULL || k <= km) {
          qk = (T(3) * (a1mm + ck[u].a * a1m) - qk) * ck[u].b;
          qv[k] = qk;
          if (FULL || k < km) {
            a1mm = a1m;
            a1m = an[u];
          }
        }
      }
    };
    int k0 = 2;
    for (; k0 + CH <= km; k0 += CH) chunk(std::true_type{}, k0);
    for (; k0 <= km; k0 += CH) chunk(std::false_type{}, k0);
    const Pair<T> cb = P1[km * ipl];
    const T rden = GAM[km * ipl];
    qv[km + 1] = (cb.a * a1m + a1mm - cb.b * qk) * rden;
  }

  // ---- pass 2: back-substitution (fv_mapz.F90:1757-1762) + interface constraints (:1783-1818); iv = 0
  {
    T r = qv[km + 1];
    T ap = T(0), a0 = A1(km), am = A1(km - 1), amm = A1(km - 2);
    auto chunk = [&](auto full, int k0) {  // full: 4 <= k < km for every level of the chunk
      constexpr bool FULL = decltype(full)::value;
      T qq[CH], gg[CH], an[CH];
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        const int k = k0 - u;
        qq[u] = (FULL || k >= 1) ? qv[k] : T(0);
        gg[u] = (FULL || k >= 1) ? GAM[(k - 1) * ipl] : T(0);
        an[u] = (AK <= 16 && (FULL || k - 3 >= 1)) ? A1(k - 3) : T(0);
      }
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        const int k = k0 - u;
        if (FULL || k >= 1) {
          r = qq[u] - gg[u] * r;
          T c = r;
          if (AK <= 16) {
            if (!FULL && (k == km || k == 2)) {
              c = f_min(c, f_max(am, a0));
              c = f_max(c, f_min(am, a0));
            } else if (FULL || k >= 3) {
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
            ap = a0;
            a0 = am;
            am = amm;
            amm = an[u];
          }
          qv[k] = c;
        }
      }
    };
    int k0 = km;
    chunk(std::false_type{}, k0);  // holds k = km
    k0 -= CH;
    for (; k0 - CH + 1 >= 4; k0 -= CH) chunk(std::true_type{}, k0);
    for (; k0 >= 1; k0 -= CH) chunk(std::false_type{}, k0);
  }

  // ---- pass 3: source-layer-major sweep (see fv3t_remap2.cuh)
  T a_0 = A1(1), a_p1 = A1(2), a_p2 = A1(3);
  T c_0 = qv[1], c_p1 = qv[2], c_p2 = qv[3];
  T g_m1 = T(0), g_0 = T(0), g_p1 = a_p1 - a_0, g_p2 = a_p2 - a_p1;
  int f_m = 0, f_0 = 0, f_p = layer_flags<T>(AK, a_p1, c_p1, c_p2, g_p1, g_p2);
  T qsum = T(0), xa = T(0), xb = T(0);
  bool zfix = false;
  int k = 1;
  // {1/dp2(k), pe2(k+1)} of the target layers reaches the thread through a four-deep cp.async ring (slot = k mod 4, stride
  // `rstride` Pairs between slots): requested three target layers ahead, completed with cp.async.wait_group.  As a plain load
  // 1/dp2 was the largest stall site (24 % of all stall samples on `qsum * rdpk`); rotated through registers two layers ahead it
  // still cost 14 % because the compiler copies the freshly loaded register at the loop join and that copy waits at once.
  // pe2 = ak + bk*ps rides along (k_remap_coef3 evaluates it once per column, uncontracted), so the loop evaluates none.
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

  auto finalize = [&](int kk, T x) { qd[(kk - 1) * ipl] = x; };  // the fillz sums are taken only where they are needed (below)
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
      finalize(k - 2, xa);
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
        finalize(km - 1, xa);
        finalize(km, xb);
      }
    }
    ++k;
    if (k <= km) {
      pe2k = pe2k1;
      dpk_m2 = dpk_m1;
      dpk_m1 = dpk;
      r2_request(k + 2);
      async_wait_pending<2>();  // at most the requests for k+1, k+2 are still in flight: slot k has landed
      const Pair<T> r2k = ring[(k & 3) * rstride];
      rdpk = r2k.a;
      pe2k1 = r2k.b;
      dpk = pe2k1 - pe2k;
    }
  };

  T pe1lo = PE1(1), pe1hi = PE1(2);
  T rdp1 = RD1[0];
  for (int l = 1; k <= km; ++l) {
    const bool have = l <= km;
    const T dp1l = pe1hi - pe1lo;
    // inputs of the NEXT source layer, requested before this layer's arithmetic so that their latency hides behind it
    T n_pe = T(0), n_rd = T(0), n_a = T(0), n_c = T(0);
    if (l < km) {
      n_pe = PE1(l + 2);
      n_rd = RD1[l * ipl];
      n_a = (l + 3 <= km) ? A1(l + 3) : T(0);
      n_c = (l + 3 <= km + 1) ? qv[l + 3] : T(0);
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
            if (MAPN) {
              T fac1 = pr + pl;
              const T fac2 = r3 * (pr * fac1 + pl * pl);
              fac1 = T(0.5) * fac1;
              v = a2 + (a4 + a3 - a2) * fac1 - a4 * fac2;
            } else {
              v = a2 + T(0.5) * (a4 + a3 - a2) * (pr + pl) - a4 * r3 * (pr * (pr + pl) + pl * pl);
            }
            done = true;
          } else {
            const T dp = pe1hi - pe2k;
            if (MAPN) {
              T fac1 = T(1) + pl;
              const T fac2 = r3 * (T(1) + pl * fac1);
              fac1 = T(0.5) * fac1;
              qsum = dp * (a2 + (a4 + a3 - a2) * fac1 - a4 * fac2);
            } else {
              qsum = dp * (a2 + T(0.5) * (a4 + a3 - a2) * (T(1) + pl) - a4 * (r3 * (T(1) + pl * (T(1) + pl))));
            }
            started = true;
          }
        }
      } else if (pe2k1 > pe1hi) {  // whole layer
        qsum = qsum + dp1l * a_0;
      } else {
        const T dp = pe2k1 - pe1lo;
        const T esl = dp * rdp1;
        if (MAPN) {
          const T fac1 = T(0.5) * esl;
          const T fac2 = T(1) - r23 * esl;
          qsum = qsum + dp * (a2 + fac1 * (a3 - a2 + a4 * fac2));
        } else {
          qsum = qsum + dp * (a2 + T(0.5) * esl * (a3 - a2 + a4 * (T(1) - r23 * esl)));
        }
        v = qsum * rdpk;
        started = false;
        done = true;
      }
      if (!done) break;
      emit(v);
    }
    // ---- advance the generators to source layer l+1
    if (l < km) {
      pe1lo = pe1hi;
      pe1hi = n_pe;
      rdp1 = n_rd;
      a_0 = a_p1;
      a_p1 = a_p2;
      a_p2 = n_a;
      c_0 = c_p1;
      c_p1 = c_p2;
      c_p2 = n_c;
      g_m1 = g_0;
      g_0 = g_p1;
      g_p1 = g_p2;
      g_p2 = a_p2 - a_p1;
      f_m = f_0;
      f_0 = f_p;
      f_p = (l + 2 <= km - 1) ? layer_flags<T>(AK, a_p1, c_p1, c_p2, g_p1, g_p2) : 0;
    }
  }
  // ---- fillz non-local rescale for the flagged columns (fv_fill.F90:131-152); re-reads this thread's own output
  //      the two sums are taken here, in the reference's order, instead of in every column
  if (p.fill && zfix) {
    const T* dp2 = p.delp + (long)t * plane * km + col;  // written by k_remap_coef3
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
template <class T> __global__ void __launch_bounds__(128) k_remap_coef3(const Remap3Params<T> p) {
  const int cols = p.n * p.n;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  remap_coef_column<T>(p, blockIdx.y, c % p.n + 1, c / p.n + 1);
}
template <class T, int AK, bool MAPN, int KM, int MINB> __global__ void __launch_bounds__(128, MINB) k_remap3(const Remap3Params<T> p) {
  __shared__ __align__(16) Pair<T> s_ring[4 * 128];
  const int cols = p.n * p.n;
  // blockIdx.x = tracer: the nq CTAs of one column block are adjacent in the grid, run together and share P1, GAM, RD1, R2, pe
  // through L2 (with the tracer as the slowest index every tracer streamed them from HBM again: 64 of 121 B per update)
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  remap3_column<T, AK, MAPN, KM>(p, s_ring + threadIdx.x, 128, blockIdx.z, c % p.n + 1, c / p.n + 1, p.iq0 + blockIdx.x);
}
#endif

}  // namespace fv3t
