// fv3atm_b200: lanes-over-levels vertical tracer remap for sm_100a (mapn_tracer with scalar_profile iv = 0, cs_limiters,
// fillz: atmos_cubed_sphere/model/fv_mapz.F90:1386-1499, 1691-2096, 2501-2576, model/fv_fill.F90:86-153).
//
// Same arithmetic per element as fv3t_remap3.cuh (interior_parabola, layer_flags, cs_limiters1 of fv3t_remap2.cuh /
// fv3t_remap.cuh; reciprocals of the shared quantities, FMA contraction), but a different decomposition.  k_remap3 gives one
// thread a whole column of one tracer: 1 KB of per-thread scratch in local memory, three passes over q, 508 instructions and
// 53 DRAM bytes per cell-update of which 17.8 are algorithmic (profiles/r01_remap3_c384_ncu.txt).  Here
//   * a WARP owns four adjacent columns of one tracer: lane = (column c = lane & 3, level group g = lane >> 2), every lane holds
//     LPL consecutive levels of its column.  A load / store instruction of the warp touches eight levels x four contiguous
//     columns (whole 32-byte sectors); q is read once and written once.
//   * the two recurrences of the cubic-spline solve (forward elimination, back-substitution) are affine in the carried value:
//     each lane runs its LPL levels with a zero carry, a three-step warp scan over the eight level groups (shuffles with
//     stride 4) delivers the true carry, and a second local pass applies it in the reference's own operation order.
//   * the column's a1 / a2 / a3 / a4 live in a per-warp shared-memory buffer: interface constraints, limiter windows and the
//     overlap integration reach neighbouring levels through it (no per-thread column arrays, no rolling-window moves).
//   * the warps of a CTA are the TRACERS of one column group: the tracer-independent per-column coefficients (spline matrix,
//     Lagrangian pressures, reciprocal layer thicknesses and the overlap plan = first source layer of every target layer,
//     k_remap_coef4) arrive once per column group by ONE bulk copy (cp.async.bulk, UBLKCP) into a double-buffered
//     shared-memory block, issued by a producer warp and completed on mbarriers.
//   * fillz is a no-op for a column without negative values (fv_fill.F90:86-153 only ever acts on q < 0): the remapped column is
//     stored at once and a per-column vote triggers the serial borrow sweep for the columns that need it.
#pragma once
#include "fv3t_advect5.cuh"
#include "fv3t_remap3.cuh"

namespace fv3t {

constexpr int R4_LVP = 152;  // level slots per coefficient field: km + 2 slots plus one pad slot per level group (see r4_cs)
enum : int { R4_D4 = 0, R4_RB = 1, R4_GAM = 2, R4_PE1 = 3, R4_RDP1 = 4, R4_RDP2 = 5, R4_NF = 6 };
constexpr int R4_MAXW = 9;   // tracer warps per CTA

// Slot of 0-based level s in a coefficient field / in the per-warp arrays: one pad slot after every level group, so that the
// eight lanes that hold the same column (level groups LPL levels apart) fall into different shared-memory banks.
FV3T_HD int r4_lpl(int km) { return km <= 63 ? 8 : 16; }
FV3T_HD int r4_cs(int s, int lpl) { return s + s / lpl; }

// one column group (4 columns) of coefficients: R4_NF fields [R4_LVP][4] of T, then the overlap plan [R4_LVP][4] of uint8
template <class T> struct R4Block {
  static constexpr int FIELD = R4_LVP * 4;
  static constexpr int BYTES = R4_NF * FIELD * (int)sizeof(T) + R4_LVP * 4;  // a multiple of 16
};
// per-warp buffer: a1, a2, a3, a4 [128 levels + one 4-element pad per level group][4 columns] + flags
template <class T, int LPL> struct R4Warp {
  static constexpr int NGRP = 8;
  static constexpr int ARR = (NGRP * LPL + NGRP) * 4;  // elements per array
  static constexpr int BYTES = 4 * ARR * (int)sizeof(T) + ARR;
  FV3T_HD static int idx(int s, int c) { return (s + s / LPL) * 4 + c; }  // s = 0-based level slot
};

template <class T> struct Remap4Params {
  const T* qsrc;   // (isd:ied, jsd:jed, km, nq) tile-major
  T* qdst;         // same layout, a different buffer
  const T* pe;     // (is-1:ie+1, km+1, js-1:je+1) tile-major
  const T *ak, *bk;
  T* delp;         // (isd:ied, jsd:jed, km) tile-major
  unsigned char* coef;  // [tile][row][group] blocks of R4Block<T>::BYTES
  unsigned char* neg;   // [tile][tracer][row][column]: the remapped column holds a negative value (fillz has work to do)
  T ptop;
  int n, km, nq, ntiles, fill;
  int iq0, nql;    // this launch remaps tracers iq0 .. iq0+nql-1
  int ntw;         // tracer warps per CTA
  int groups_per_cta;
};

FV3T_HD int r4_groups_per_row(int n) { return (n + 3) / 4; }

// ---- tracer-independent column coefficients: one thread per column ---------------------------------------------------------
template <class T> FV3T_HD void remap_coef4_column(const Remap4Params<T>& p, int t, int i, int j) {
  const int n = p.n, km = p.km;
  const long nd = n + 6, plane = nd * nd;
  const long pe_ld1 = n + 2, pe_ld2 = pe_ld1 * (km + 1);
  const T* pe = p.pe + (long)t * pe_ld2 * (n + 2) + (long)i + (long)j * pe_ld2;
  const long col = (long)(j + 2) * nd + (i + 2);
  T* delp = p.delp + (long)t * plane * km + col;
  const int NG = r4_groups_per_row(n);
  const long blk = ((long)t * n + (j - 1)) * NG + (i - 1) / 4;
  const int c = (i - 1) & 3;
  T* f = reinterpret_cast<T*>(p.coef + blk * R4Block<T>::BYTES) + c;
  unsigned char* l0 = p.coef + blk * R4Block<T>::BYTES + (long)R4_NF * R4Block<T>::FIELD * sizeof(T) + c;
  const int lpl = r4_lpl(km);
  auto F = [&](int field, int s) -> T& { return f[(field * R4_LVP + r4_cs(s, lpl)) * 4]; };
  auto L0 = [&](int s) -> unsigned char& { return l0[r4_cs(s, lpl) * 4]; };
  auto PE1 = [&](int k) -> T { return pe[(long)(k - 1) * pe_ld1]; };
  const T ps = PE1(km + 1);
  auto PE2 = [&](int k) -> T { return k == 1 ? p.ptop : (k == km + 1 ? ps : add_rn(p.ak[k - 1], mul_rn(p.bk[k - 1], ps))); };  // uncontracted: delp is caller-visible
  // spline matrix (fv_mapz.F90:1736-1755), as in remap_coef_column (fv3t_remap3.cuh)
  T pa = PE1(1), pb = PE1(2), pc = PE1(3);
  T dpm = pb - pa, dpc = pc - pb;
  const T grat = dpc / dpm;
  T bet = grat * (grat + T(0.5));
  F(R4_D4, 0) = (grat + grat) * (grat + T(1));  // ctop
  F(R4_RB, 0) = T(1) / bet;
  T gprev = (T(1) + grat * (grat + T(1.5))) / bet;
  F(R4_GAM, 0) = gprev;
  F(R4_PE1, 0) = pa;
  F(R4_PE1, 1) = pb;
  F(R4_RDP1, 0) = T(1) / dpm;
  T d4 = T(0);
  for (int k = 2; k <= km; ++k) {
    d4 = dpm / dpc;
    bet = T(2) + d4 + d4 - gprev;
    gprev = d4 / bet;
    F(R4_D4, k - 1) = d4;
    F(R4_RB, k - 1) = T(1) / bet;
    F(R4_GAM, k - 1) = gprev;
    F(R4_RDP1, k - 1) = T(1) / dpc;
    F(R4_PE1, k) = pc;
    if (k < km) {
      pb = pc;
      pc = PE1(k + 2);
      dpm = dpc;
      dpc = pc - pb;
    }
  }
  const T a_bot = T(1) + d4 * (d4 + T(1.5));
  F(R4_D4, km) = T(2) * d4 * (d4 + T(1));                      // cbot
  F(R4_D4, km + 1) = a_bot;
  F(R4_RB, km) = T(1) / (d4 * (d4 + T(0.5)) - a_bot * gprev);  // 1/den of the closure
  F(R4_GAM, km) = T(0);
  const int nslot = 8 * lpl;  // slots the remap kernel may touch: 0 .. 8*LPL - 1 (+ two look-ahead levels)
  for (int s = km + 1; s < nslot + 2; ++s) {
    F(R4_RB, s) = T(0);
    F(R4_GAM, s) = T(0);
    if (s > km + 1) F(R4_D4, s) = T(0);
    F(R4_PE1, s) = ps;
  }
  for (int s = km; s < nslot + 2; ++s) {
    F(R4_RDP1, s) = T(0);
    F(R4_RDP2, s) = T(0);
    L0(s) = (unsigned char)km;
  }
  // target grid: delp <- dp2 (fv_mapz.F90:263-272, 364-368), 1/dp2, and the overlap plan: the source layer in which the
  // reference's search (fv_mapz.F90:1428-1436) finds the top of target layer k
  T p2a = PE2(1);
  int lsrc = 1;
  for (int k = 1; k <= km; ++k) {
    const T p2b = PE2(k + 1);
    const T dp2 = p2b - p2a;
    delp[(long)(k - 1) * plane] = dp2;
    F(R4_RDP2, k - 1) = T(1) / dp2;
    while (lsrc < km && !(p2a >= PE1(lsrc) && p2a <= PE1(lsrc + 1))) ++lsrc;
    L0(k - 1) = (unsigned char)lsrc;
    // the search for the next target resumes in the layer where this one ends
    while (lsrc < km && p2b > PE1(lsrc + 1)) ++lsrc;
    p2a = p2b;
  }
}

#ifdef __CUDACC__
template <class T> __global__ void __launch_bounds__(128) k_remap_coef4(const Remap4Params<T> p) {
  const int cols = p.n * p.n;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  remap_coef4_column<T>(p, blockIdx.y, c % p.n + 1, c / p.n + 1);
}

__device__ __forceinline__ void r4_bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(a5_smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(a5_smem_u32(bar))
               : "memory");
}

template <class T> __device__ __forceinline__ T r4_shfl_up(T v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
template <class T> __device__ __forceinline__ T r4_shfl_down(T v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }

// one column group (columns i0 .. i0+3 of row j, tile t) of tracer iq, by one warp.  cf / l0s: this lane's column of the
// coefficient block in shared memory; A1s .. FLs: the warp's buffer.
template <class T, int AK, int LPL>
__device__ __forceinline__ void remap4_group(const Remap4Params<T>& p, const T* cf, const unsigned char* l0s, T* A1s, T* Q2s, T* Q3s, T* Q4s,
                                             unsigned char* FLs, const T* s_ak, const T* s_bk, int t, int j, int i0, int iq, int lane) {
  using W = R4Warp<T, LPL>;
  const int n = p.n, km = p.km;
  const long nd = n + 6, plane = nd * nd;
  const int c = lane & 3, g = lane >> 2, s0 = g * LPL;
  auto CF = [&](int field, int s) -> T { return cf[(field * R4_LVP + s + s / LPL) * 4]; };
  const bool colvalid = i0 + c <= n;
  const int ic = colvalid ? i0 + c : n;
  const long off = (((long)t * p.nq + iq) * km) * plane + (long)(j + 2) * nd + (ic + 2);
  const T* __restrict__ qs = p.qsrc + off;
  T* __restrict__ qd = p.qdst + off;
  const T r3 = K<T>::r3(), r23 = K<T>::r23();

  // ---- 1. this lane's LPL cell means; the whole column to the warp buffer
  T A[LPL];
#pragma unroll
  for (int e = 0; e < LPL; ++e) A[e] = (s0 + e < km) ? qs[(long)(s0 + e) * plane] : T(0);
#pragma unroll
  for (int e = 0; e < LPL; ++e) A1s[W::idx(s0 + e, 0)] = A[e];
  __syncwarp();
  auto a1s = [&](int s) -> T { return A1s[W::idx(s < 0 ? 0 : s, 0)]; };  // 0-based level slot of this lane's column

  // ---- 2. forward elimination (fv_mapz.F90:1736-1755): x(k) = (rhs(k) - mul(k) x(k-1)) rb(k) for the interface k = s+1
  T rhs[LPL], QP[LPL];
  {
    const T am1 = a1s(s0 - 1), akm = a1s(km - 1), akm1 = a1s(km - 2);
#pragma unroll
    for (int e = 0; e < LPL; ++e) {
      const int s = s0 + e;
      const T d4 = CF(R4_D4, s);
      const T prev = e == 0 ? am1 : A[e == 0 ? 0 : e - 1];
      T r = T(3) * (prev + d4 * A[e]);
      if (s == 0) r = d4 * A[0] + A[1];           // top closure: ctop a1(1) + a1(2)
      if (s == km) r = d4 * akm + akm1;           // bottom closure: cbot a1(km) + a1(km-1)
      if (s > km) r = T(0);
      rhs[e] = r;
    }
    const T abot = CF(R4_D4, km + 1);
    T rbv[LPL];
#pragma unroll
    for (int e = 0; e < LPL; ++e) rbv[e] = CF(R4_RB, s0 + e);  // zero beyond the closure
    T x = T(0), M = T(1);
#pragma unroll
    for (int e = 0; e < LPL; ++e) {
      const T mul = (s0 + e == km) ? abot : T(1);
      x = (rhs[e] - mul * x) * rbv[e];
      M = M * (-(mul * rbv[e]));
    }
    T E = x;
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
      const T Eu = r4_shfl_up(E, 4 * d), Mu = r4_shfl_up(M, 4 * d);
      if (g >= d) {
        E = E + M * Eu;
        M = M * Mu;
      }
    }
    T carry = r4_shfl_up(E, 4);
    if (g == 0) carry = T(0);
    x = carry;
#pragma unroll
    for (int e = 0; e < LPL; ++e) {
      const T mul = (s0 + e == km) ? abot : T(1);
      x = (rhs[e] - mul * x) * rbv[e];
      QP[e] = x;
    }
  }
  // ---- 3. back-substitution (fv_mapz.F90:1757-1762): q(k) = q'(k) - gam(k) q(k+1)
  {
    T gm[LPL];
#pragma unroll
    for (int e = 0; e < LPL; ++e) gm[e] = CF(R4_GAM, s0 + e);
    T y = T(0), G = T(1);
#pragma unroll
    for (int e = LPL - 1; e >= 0; --e) {
      y = QP[e] - gm[e] * y;
      G = G * (-gm[e]);
    }
    T E = y;
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
      const T Ed = r4_shfl_down(E, 4 * d), Gd = r4_shfl_down(G, 4 * d);
      if (g + d < 8) {
        E = E + G * Ed;
        G = G * Gd;
      }
    }
    T carry = r4_shfl_down(E, 4);
    if (g == 7) carry = T(0);
    y = carry;
#pragma unroll
    for (int e = LPL - 1; e >= 0; --e) {
      y = QP[e] - gm[e] * y;
      QP[e] = y;
    }
  }
  // ---- 4. interface constraints (fv_mapz.F90:1783-1818, iv = 0) and the interface values into the layers' a2 / a3 slots
  {
    const T b1 = a1s(s0 - 1), b2 = a1s(s0 - 2), n1 = a1s(s0 + LPL);
#pragma unroll
    for (int e = 0; e < LPL; ++e) {
      const int s = s0 + e, k = s + 1;
      T cv = QP[e];
      if (AK <= 16 && k >= 2 && k <= km) {
        const T am = e >= 1 ? A[e == 0 ? 0 : e - 1] : b1;                    // a1(k-1)
        const T a0 = A[e];                                                    // a1(k)
        if (k == km || k == 2) {
          cv = f_min(cv, f_max(am, a0));
          cv = f_max(cv, f_min(am, a0));
        } else {
          const T amm = e >= 2 ? A[e < 2 ? 0 : e - 2] : (e == 1 ? b1 : b2);  // a1(k-2)
          const T ap = e + 1 < LPL ? A[e + 1 < LPL ? e + 1 : 0] : n1;         // a1(k+1)
          const T gmm = am - amm, gpp = ap - a0;
          if (gmm * gpp > T(0)) {
            cv = f_min(cv, f_max(am, a0));
            cv = f_max(cv, f_min(am, a0));
          } else if (gmm > T(0)) {
            cv = f_max(cv, f_min(am, a0));
          } else {
            cv = f_min(cv, f_max(am, a0));
            cv = f_max(T(0), cv);
          }
        }
      }
      if (s < km) Q2s[W::idx(s, 0)] = cv;                   // a2 of layer k
      if (s >= 1 && s <= km) Q3s[W::idx(s - 1, 0)] = cv;    // a3 of layer k-1
    }
  }
  __syncwarp();
  // ---- 5. layer flags (fv_mapz.F90:1820-1846) for the layers 2 .. km-1
#pragma unroll 4
  for (int e = 0; e < LPL; ++e) {
    const int s = s0 + e, l = s + 1;
    int f = 0;
    if (l >= 2 && l <= km - 1) {
      const T a0 = a1s(s);
      f = layer_flags<T>(AK, a0, Q2s[W::idx(s, 0)], Q3s[W::idx(s, 0)], a0 - a1s(s - 1), a1s(s + 1) - a0);
    }
    FLs[W::idx(s, 0)] = (unsigned char)f;
  }
  __syncwarp();
  // ---- 6. limited parabolas (fv_mapz.F90:1848-2096 / cs_limiters), in place
#pragma unroll 2
  for (int e = 0; e < LPL; ++e) {
    const int s = s0 + e, l = s + 1;
    if (l <= km) {
      const T a0 = a1s(s);
      T a2 = Q2s[W::idx(s, 0)], a3 = Q3s[W::idx(s, 0)], a4;
      if (AK > 16) {
        a4 = T(3) * (T(2) * a0 - (a2 + a3));
      } else if (l >= 3 && l <= km - 2) {
        const T am2 = a1s(s - 2), am1 = a1s(s - 1), ap1 = a1s(s + 1), ap2 = a1s(s + 2);
        interior_parabola<T>(AK, a0, a2, a3, a4, am1 - am2, a0 - am1, ap1 - a0, ap2 - ap1, FLs[W::idx(s - 1, 0)], FLs[W::idx(s, 0)],
                             FLs[W::idx(s + 1, 0)], T(0));
      } else {
        if (l == 1) a2 = f_max(T(0), a2);
        if (l == km) a3 = f_max(T(0), a3);
        a4 = T(3) * (T(2) * a0 - (a2 + a3));
        cs_limiters1<T>(FLs[W::idx(s, 0)] & 1, a0, a2, a3, a4, (l == 1 || l == km) ? 1 : 2);
      }
      Q2s[W::idx(s, 0)] = a2;
      Q3s[W::idx(s, 0)] = a3;
      Q4s[W::idx(s, 0)] = a4;
    }
  }
  __syncwarp();
  // ---- 7. overlap integration onto the target layers (fv_mapz.F90:1428-1486), one target layer per step
  const T ps = CF(R4_PE1, km);
  auto PE2 = [&](int k) -> T { return k == 1 ? p.ptop : (k == km + 1 ? ps : add_rn(s_ak[k - 1], mul_rn(s_bk[k - 1], ps))); };
  bool neg = false;
  T p2a = PE2(s0 + 1 <= km + 1 ? s0 + 1 : km + 1);
#pragma unroll 2
  for (int e = 0; e < LPL; ++e) {
    const int s = s0 + e, k = s + 1;
    if (k <= km) {
      const T p2b = PE2(k + 1);
      const int l = l0s[W::idx(s, 0)];  // 1-based source layer holding the top of the target layer
      const T pe1l = CF(R4_PE1, l - 1), pe1h = CF(R4_PE1, l), rd = CF(R4_RDP1, l - 1);
      const int li = W::idx(l - 1, 0);
      T a2 = Q2s[li], a3 = Q3s[li], a4 = Q4s[li];
      const T pl = (p2a - pe1l) * rd;
      T v;
      if (p2b <= pe1h) {
        const T pr = (p2b - pe1l) * rd;
        T fac1 = pr + pl;
        const T fac2 = r3 * (pr * fac1 + pl * pl);
        fac1 = T(0.5) * fac1;
        v = a2 + (a4 + a3 - a2) * fac1 - a4 * fac2;
      } else {
        T fac1 = T(1) + pl;
        const T fac2 = r3 * (T(1) + pl * fac1);
        fac1 = T(0.5) * fac1;
        T qsum = (pe1h - p2a) * (a2 + (a4 + a3 - a2) * fac1 - a4 * fac2);
        T lo = pe1h;
        for (int m = l + 1; m <= km; ++m) {
          const T hi = CF(R4_PE1, m);
          const int mi = W::idx(m - 1, 0);
          if (p2b > hi) {
            qsum = qsum + (hi - lo) * A1s[mi];
            lo = hi;
          } else {
            const T dp = p2b - lo;
            const T esl = dp * CF(R4_RDP1, m - 1);
            a2 = Q2s[mi];
            a3 = Q3s[mi];
            a4 = Q4s[mi];
            qsum = qsum + dp * (a2 + T(0.5) * esl * (a3 - a2 + a4 * (T(1) - r23 * esl)));
            break;
          }
        }
        v = qsum * CF(R4_RDP2, s);
      }
      neg = neg || (v < T(0));
      if (colvalid) qd[(long)s * plane] = v;
      p2a = p2b;
    }
  }
  // ---- 8. fillz (fv_fill.F90:86-153) only ever acts on columns that hold a negative value: flag them for k_fillz4
  if (p.fill) {
    const unsigned bal = __ballot_sync(0xffffffffu, neg);
    if (g == 0 && colvalid) p.neg[(((long)t * p.nq + iq) * n + (j - 1)) * n + (ic - 1)] = ((bal >> c) & 0x11111111u) ? 1 : 0;
  }
  __syncwarp();  // the warp buffer is rewritten by the next column group
}

// CTA = p.ntw tracer warps + one producer warp; it walks p.groups_per_cta consecutive column groups
template <class T, int AK, int LPL> __global__ void __launch_bounds__(32 * (R4_MAXW + 1), 1) k_remap4(const __grid_constant__ Remap4Params<T> p) {
  extern __shared__ __align__(128) unsigned char smem4r[];
  using W = R4Warp<T, LPL>;
  constexpr int CB = (R4Block<T>::BYTES + 127) & ~127;
  constexpr int WB = (W::BYTES + 127) & ~127;
  constexpr int OFF_BAR = 2 * CB, OFF_AK = OFF_BAR + 128, OFF_W = (OFF_AK + 2 * 132 * (int)sizeof(T) + 127) & ~127;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem4r + OFF_BAR);
  uint64_t* empty = full + 2;
  T* s_ak = reinterpret_cast<T*>(smem4r + OFF_AK);
  T* s_bk = s_ak + 132;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NG = r4_groups_per_row(p.n);
  const long total = (long)p.ntiles * p.n * NG;
  const long first = (long)blockIdx.x * p.groups_per_cta;
  const long last = first + p.groups_per_cta < total ? first + p.groups_per_cta : total;
  const int iq_first = p.iq0 + blockIdx.y * p.ntw;
  const int ntw = min(p.ntw, p.iq0 + p.nql - iq_first);
  for (int k = threadIdx.x; k <= p.km; k += blockDim.x) {
    s_ak[k] = p.ak[k];
    s_bk[k] = p.bk[k];
  }
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      a5_mbar_init(&full[b], 1);
      a5_mbar_init(&empty[b], ntw);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (first >= last) return;
  if (warp == p.ntw) {  // producer warp
    if (lane == 0) {
      int b = 0, wrap = 0;
      for (long grp = first; grp < last; ++grp) {
        if (grp - first >= 2) a5_mbar_wait(&empty[b], (wrap - 1) & 1);
        a5_mbar_expect_tx(&full[b], R4Block<T>::BYTES);
        r4_bulk_g2s(smem4r + b * CB, p.coef + grp * R4Block<T>::BYTES, R4Block<T>::BYTES, &full[b]);
        if (++b == 2) {
          b = 0;
          ++wrap;
        }
      }
    }
    return;
  }
  if (warp >= ntw) return;
  // per-lane views: column c = lane & 3 of the warp's arrays and of the coefficient blocks
  T* A1s = reinterpret_cast<T*>(smem4r + OFF_W + warp * WB) + (lane & 3);
  T* Q2s = A1s + W::ARR;
  T* Q3s = Q2s + W::ARR;
  T* Q4s = Q3s + W::ARR;
  unsigned char* FLs = smem4r + OFF_W + warp * WB + 4 * W::ARR * (int)sizeof(T) + (lane & 3);
  const int iq = iq_first + warp;
  int b = 0, wrap = 0;
  for (long grp = first; grp < last; ++grp) {
    const int m = (int)(grp % NG);
    const long row = grp / NG;
    const int j = (int)(row % p.n) + 1, t = (int)(row / p.n);
    a5_mbar_wait(&full[b], wrap & 1);
    const T* cf = reinterpret_cast<const T*>(smem4r + b * CB) + (lane & 3);
    const unsigned char* l0s = smem4r + b * CB + R4_NF * R4Block<T>::FIELD * (int)sizeof(T) + (lane & 3);
    remap4_group<T, AK, LPL>(p, cf, l0s, A1s, Q2s, Q3s, Q4s, FLs, s_ak, s_bk, t, j, 1 + 4 * m, iq, lane);
    if (lane == 0) a5_mbar_arrive(&empty[b]);
    if (++b == 2) {
      b = 0;
      ++wrap;
    }
  }
}

// fillz (fv_fill.F90:86-153) for the columns k_remap4 flagged: one thread per column, in place on the remapped field; dp = the
// Eulerian delp written by k_remap_coef4.  grid: (column blocks, tiles, tracers of the launch)
template <class T> __global__ void __launch_bounds__(128) k_fillz4(const Remap4Params<T> p) {
  const int n = p.n, km = p.km;
  const int cidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (cidx >= n * n) return;
  const int t = blockIdx.y, iq = p.iq0 + blockIdx.z;
  if (!p.neg[((long)t * p.nq + iq) * n * n + cidx]) return;
  const int i = cidx % n + 1, j = cidx / n + 1;
  const long nd = n + 6, plane = nd * nd;
  const long col = (long)(j + 2) * nd + (i + 2);
  T* q = p.qdst + (((long)t * p.nq + iq) * km) * plane + col;
  const T* dpp = p.delp + (long)t * plane * km + col;
  auto DP = [&](int k) -> T { return dpp[(long)(k - 1) * plane]; };
  auto QV = [&](int k) -> T& { return q[(long)(k - 1) * plane]; };
  // streaming form of the reference sweep (the same state machine as the emit() of fv3t_remap3.cuh): at step k the window is
  // (xa, xb, xc) = layers k-2, k-1, k with thicknesses (da, db, dc); layer k-1 borrows from above, then from below
  T xa = QV(1), xb = QV(2);
  T da = DP(1), db = DP(2);
  if (xa < T(0)) {  // top layer: push the deficit down (fv_fill.F90:88-93)
    xb = xb + xa * da / db;
    xa = T(0);
  }
  bool zfix = false;
  T sum0 = T(0), sum1 = T(0);
  auto finalize = [&](int kk, T x, T dp) {
    QV(kk) = x;
    if (kk >= 2) {
      const T m = x * dp;
      sum0 = sum0 + m;
      sum1 = sum1 + f_max(T(0), m);
    }
  };
  for (int k = 3; k <= km; ++k) {
    T xc = QV(k);
    const T dc = DP(k);
    if (xb < T(0)) {
      zfix = true;
      if (xa > T(0)) {
        const T dq = f_min(xa * da, -xb * db);
        xa = xa - dq / da;
        xb = xb + dq / db;
      }
      if (xb < T(0) && xc > T(0)) {
        const T dq = f_min(xc * dc, -xb * db);
        xc = xc - dq / dc;
        xb = xb + dq / db;
      }
    }
    finalize(k - 2, xa, da);
    xa = xb;
    da = db;
    xb = xc;
    db = dc;
  }
  if (xb < T(0) && xa > T(0)) {  // bottom layer: borrow from the layer above (fv_fill.F90:117-128)
    zfix = true;
    const T qup = xa * da;
    const T qly = -xb * db;
    const T dup = f_min(qly, qup);
    xa = xa - dup / da;
    xb = xb + dup / db;
  }
  finalize(km - 1, xa, da);
  finalize(km, xb, db);
  if (zfix && sum0 > T(0)) {  // non-local rescale (fv_fill.F90:131-152)
    const T fac = sum0 / sum1;
    for (int k = 2; k <= km; ++k) {
      const T dp = DP(k);
      QV(k) = f_max(T(0), fac * (QV(k) * dp) / dp);
    }
  }
}
#endif

}  // namespace fv3t
