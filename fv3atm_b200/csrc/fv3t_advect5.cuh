// fv3atm_b200: multi-tracer marching advection CTA with TMA-staged level fields (sm_100a).
//
// Same operator as fv3t_advect3/4.cuh (one tracer_2d sub-step, atmos_cubed_sphere/model/fv_tracer2d.F90:503-556 -> fv_tp_2d,
// model/tp_core.F90:110-249 with xppm :332-704, yppm :707-1124, copy_corners :253-330, pert_ppm :1178-1236) and the same
// per-element arithmetic (fv3t_ppm.cuh, ystream_core).  What changes is the decomposition:
//   * one CTA = one 58-column strip of one (tile, level) for UP TO TEN TRACERS: warp 0 is a TMA producer, every following
//     pair of warps (64 threads, one column each) marches one tracer over the rows of the tile.  The tracer-independent level
//     fields ({cx,xfx}, {cy,yfx}, 1/ra_x, 1/ra_y, mfx, mfy, area, {dp1/dp2, rarea/2dp2}) are fetched ONCE per strip and
//     level -- not once per tracer -- by cp.async.bulk.tensor (UTMALDG) into a three-stage shared-memory ring of four-row
//     boxes, completion on mbarriers; k_advect4 issued eleven per-thread LDGSTS plus their address arithmetic per tracer-cell
//     and moved 2.1x the algorithmic bytes through DRAM (profiles/r01_advect4_c384_ncu.txt, r01_traffic_c768.json).
//   * the row loop is unrolled by four with phase-indexed windows (YWin): the rolling y-stream windows and the delayed inner
//     x flux rotate by renaming, not by register moves.
//   * rows / strips away from the tile edges run instantiations with the tile-edge formulas, the corner views of q and all
//     row clamps compiled out.
//   * the two-warp group of a tracer synchronises on its own named barrier; the groups of a CTA drift apart freely.
// The per-thread phase functions run unchanged on the host (tests/hostsim/advect5_hostsim.cu, test infrastructure).
#pragma once
#include <cuda.h>  // CUtensorMap (type only; the driver entry point is resolved at run time in fv3t_fast.cu)

#include "fv3t_advect4.cuh"

namespace fv3t {

constexpr int A5_R = 4;    // row steps per staged box
constexpr int A5_NS = 3;   // stages of the ring
constexpr int A5_GW = 64;  // threads (columns) per tracer group: 58 compute columns + 3 halo columns either side
constexpr int A5_W = A5_GW - 6;
constexpr int A5_XP = A5_GW + 8;  // pitch of an exchange row (4 pad slots either side)
constexpr int A5_XPAD = 4;
// exchange rows of a tracer group: 0..3 q of row r (slot = row step mod 4)  4,5 q_i of row o (step parity)  6 dm/al of row r
// 7 dm/al of row o  8 xfx*fx2  9 (fx+fx2)*mfx
constexpr int A5_XROWS = 10;
enum : int { A5X_Q = 0, A5X_QI = 4, A5X_DR = 6, A5X_DO = 7, A5X_SF1 = 8, A5X_SFT = 9 };
// private per-thread slots (A5_GW apart): 0..3 qy of the corner rows (step mod 4)  4..7 inner x flux of the last four rows
// 8,9 / 10,11 tile-edge values of the inner / outer y stream  12,13 inner y flux  14,15 q of the output row (step parity)
enum : int { A5V_QY = 0, A5V_FX2 = 4, A5V_EIN = 8, A5V_EOU = 10, A5V_FY2 = 12, A5V_QO = 14, A5V_N = 16 };
constexpr int A5_MAXTG = 9;   // 1 + 2*9 = 19 warps -> 96 registers per thread at one CTA per SM

// padded row pitch of the scratch planes: a multiple of 16 bytes in fp32 and fp64 (TMA global strides)
FV3T_HD int a5_pitch(int n) { return (n + 6 + 3) & ~3; }

// staged arrays of one box (A5_R rows x A5_GW columns each)
enum : int { A5_XR = 0, A5_XO = 1, A5_Y2 = 2, A5_CAB = 3, A5_NPAIR = 4 };                           // Pair<T>
enum : int { A5_RX = 0, A5_MFX = 1, A5_RY = 2, A5_MFY = 3, A5_AR = 4, A5_AO = 5, A5_RA = 6, A5_NSC = 7 };  // T
template <class T> struct A5Stage {
  // A TMA box must start on a 16-byte boundary of global memory (the hardware raises an illegal-instruction error otherwise).
  // Strip s starts at plane column 58 s: 464 s bytes for fp64 and for the fp32 pairs, but 232 s bytes for fp32 scalars -- their
  // boxes start at the column rounded down to a multiple of four and are four columns wider; the lanes read at that shift.
  static constexpr int SW = sizeof(T) == 4 ? A5_GW + 4 : A5_GW;  // columns of a scalar box
  FV3T_HD static int shift(int xs) { return sizeof(T) == 4 ? (xs & 3) : 0; }
  static constexpr int PAIR_BYTES = A5_NPAIR * A5_R * A5_GW * 2 * (int)sizeof(T);
  static constexpr int SBYTES = (A5_R * SW * (int)sizeof(T) + 127) & ~127;  // one scalar box (TMA destinations: 128-byte aligned)
  static constexpr int SFIELD = SBYTES / (int)sizeof(T);
  static constexpr int TX_BYTES = PAIR_BYTES + A5_NSC * A5_R * SW * (int)sizeof(T);  // bytes the ten boxes deliver
  static constexpr int BYTES = PAIR_BYTES + A5_NSC * SBYTES;
  static constexpr int PRIV = A5V_N;
  static constexpr int GROUP_ELEMS = A5_XROWS * A5_XP + PRIV * A5_GW;
  static constexpr int BAR_OFF = A5_NS * BYTES;
  static constexpr int GROUP_OFF = BAR_OFF + 128;
  static size_t smem_bytes(int tg) { return (size_t)GROUP_OFF + (size_t)tg * GROUP_ELEMS * sizeof(T); }
};

// Sub-tile contexts (a rank that owns PART of a cubed-sphere tile, fv_arrays.F90:1178-1186 with layout > 1x1): the resident
// "tile" is a square sub-domain; the tile-edge formulas of xppm / yppm apply only on the sides that lie on a tile edge and the
// corner views of copy_corners only at true tile corners (gridstruct%sw_corner .., fv_arrays.F90:181) -- everywhere else the
// halo holds the neighbouring sub-domain's cells (diagonal neighbours included) and the interior formulas apply.
// The PPM element functions select the edge formulas by index against 1 and npx: a missing west / south edge is expressed by
// SHIFTING the index they see (and the accessors back) by A5_SHIFT, a missing east / north edge by an unreachable npx.
struct A5Sub {
  int no_w = 0, no_e = 0, no_s = 0, no_n = 0;  // 1: that side of the resident square is NOT a tile edge
  int cmask = 15;                              // true tile corners: 1 SW, 2 SE, 4 NE, 8 NW
};
constexpr int A5_SHIFT = 1 << 20;
constexpr int A5_NOEDGE = 1 << 28;

struct alignas(64) Adv5Maps {
  CUtensorMap x2, y2, cab, rx, ry, mfx, mfy, area, rarea;
};

template <class T> struct Adv5Params {
  const T* qin;
  T* qout;
  // scratch of the resident level chunk, padded plane layout [level][row 0..nd-1][PP] (row = j+2, column = i+2)
  const Pair<T>*X2, *Y2, *CAB;
  const T *RX, *RY, *MFX, *MFY;
  const T *AREA, *RAREA;  // [tile][nd][PP]
  const T *dxa, *dya;   // Fortran layout (tile-edge formulas only)
  const int* ksplt;
  int n, npz, nq, ntiles, it;
  int lev0;             // global level (tile*npz + kz) of chunk level 0
  int tg;               // tracers per CTA (block = 32 + 64*tg threads)
  int iq0, nql;         // this launch advects tracers iq0 .. iq0+nql-1
  T lim_fac;
  static constexpr bool SUB = false;
};
// the parameters of the instantiations that serve sub-tile contexts: the flags of every resident sub-domain ride along.  A type of
// its own, so that the whole-tile instantiations stay exactly what they were (the lookups cost the exact-arithmetic hord-10 kernel 6 %)
template <class T> struct Adv5ParamsSub : Adv5Params<T> {
  A5Sub sub[6];
  static constexpr bool SUB = true;
};

// ---------------------------------------------------------------------------------------------------------------------
// k_prep5: the tracer-independent part of one sub-step for a chunk of levels (fv_tracer2d.F90:387-405 xfx/yfx, :449-486
// frac scaling, :510-526 dp2 / ra_x / ra_y, :547-553 dp1 <- dp2).  Reads the UNSCALED cx, cy, mfx, mfy (k_scale3 scales them
// in place when the call finishes).  mode_all = 0 (it > 1, scratch holds every level): only dp1 and CAB are refreshed.
// ---------------------------------------------------------------------------------------------------------------------
template <class T> struct Prep5Params {
  const T *cx, *cy, *mfx, *mfy;
  T* dp1;
  GridDev<T> g;
  Pair<T>*X2, *Y2, *CAB;
  T *RX, *RY, *MFX, *MFY;
  const int* ksplt;
  int n, npz, ntiles, lev0, nlev, it, mode_all;
  int exact;  // 1: scratch for the exact-arithmetic instantiations (ra_x, ra_y, {dp1, dp2}); 0: reciprocals and {dp1/dp2, rarea/2dp2}
  int mode_1l = 0;  // tracer_2d_1L: dp1 <- dp2 only between the level's own sub-steps (fv_tracer2d.F90:305)
};

template <class T> FV3T_HD void prep5_cell(const Prep5Params<T>& p, int levc, int e) {
  const int n = p.n, nd = n + 6, PP = a5_pitch(n);
  const long plane = (long)nd * nd;
  const long lev = p.lev0 + levc;
  const int t = (int)(lev / p.npz), kz = (int)(lev % p.npz);
  const int ks = p.ksplt[kz];
  if (p.it - 1 > ks) return;  // the level took no part in sub-step it-1: nothing changes any more
  if (p.mode_1l && p.it > ks) return;
  const int j = e / nd - 2, i = e % nd - 2;
  const T frac = T(1) / (T)ks;
  const T* area = p.g.area + (long)t * plane;
  const T* rarea = p.g.rarea + (long)t * plane;
  const T* mxp = p.mfx + lev * (long)(n + 1) * n;
  const T* myp = p.mfy + lev * (long)n * (n + 1);
  const long o = ((long)levc * nd + (j + 2)) * PP + (i + 2);
  const bool cell = i >= 1 && i <= n && j >= 1 && j <= n;
  Pair<T> ab{T(0), T(0)};
  if (cell) {
    const long ox = (long)(j - 1) * (n + 1) + (i - 1), oy = (long)(j - 1) * n + (i - 1);
    const T rar = rarea[e];
    const T m0 = mul_rn(mxp[ox], frac), m1 = mul_rn(mxp[ox + 1], frac), m2 = mul_rn(myp[oy], frac), m3 = mul_rn(myp[oy + n], frac);
    T d1 = p.dp1[lev * plane + e];
    if (p.it > 1) {  // dp1 <- dp2 of sub-step it-1 (fv_tracer2d.F90:547-553)
      d1 = dp2_of<T>(d1, m0, m1, m2, m3, rar);
      p.dp1[lev * plane + e] = d1;
    }
    if (p.it <= ks) {
      const T d2 = dp2_of<T>(d1, m0, m1, m2, m3, rar);
      if (p.exact) {
        ab.a = d1;
        ab.b = d2;
      } else {
        const T r2 = T(1) / d2;
        ab.a = d1 * r2;
        ab.b = T(0.5) * rar * r2;
      }
    }
  }
  if (p.it > ks) return;
  p.CAB[o] = ab;
  if (!p.mode_all) return;
  const T* dxa = p.g.dxa + (long)t * plane;
  const T* dya = p.g.dya + (long)t * plane;
  const T* dxg = p.g.dx + (long)t * nd * (nd + 1);
  const T* dyg = p.g.dy + (long)t * (nd + 1) * nd;
  const T* ssg = p.g.sin_sg + (long)t * plane * 5;
  const T* cxp = p.cx + lev * (long)(n + 1) * nd;
  const T* cyp = p.cy + lev * (long)nd * (n + 1);
  Pair<T> x2{T(0), T(0)}, y2{T(0), T(0)};
  T rx = T(0), ry = T(0), mx = T(0), my = T(0);
  if (i >= 1 && i <= n + 1) {
    T c;
    const T xf = xfx_of<T>(cxp, dxa, dyg, ssg, plane, nd, n, i, j, frac, c);
    x2.a = mul_rn(c, frac);
    x2.b = xf;
    if (i <= n) {
      T c1;
      const T xf1 = xfx_of<T>(cxp, dxa, dyg, ssg, plane, nd, n, i + 1, j, frac, c1);
      rx = add_rn(add_rn(area[e], xf), -xf1);  // ra_x (fv_tracer2d.F90:522-526)
      if (!p.exact) rx = T(1) / rx;
    }
    if (j >= 1 && j <= n) mx = mul_rn(mxp[(long)(j - 1) * (n + 1) + (i - 1)], frac);
  }
  if (j >= 1 && j <= n + 1) {
    T c;
    const T yf = yfx_of<T>(cyp, dya, dxg, ssg, plane, nd, n, i, j, frac, c);
    y2.a = mul_rn(c, frac);
    y2.b = yf;
    if (j <= n) {
      T c1;
      const T yf1 = yfx_of<T>(cyp, dya, dxg, ssg, plane, nd, n, i, j + 1, frac, c1);
      ry = add_rn(add_rn(area[e], yf), -yf1);  // ra_y (fv_tracer2d.F90:517-521)
      if (!p.exact) ry = T(1) / ry;
    }
    if (i >= 1 && i <= n) my = mul_rn(myp[(long)(j - 1) * n + (i - 1)], frac);
  }
  p.X2[o] = x2;
  p.Y2[o] = y2;
  p.RX[o] = rx;
  p.RY[o] = ry;
  p.MFX[o] = mx;
  p.MFY[o] = my;
}

#ifdef __CUDACC__
template <class T> __global__ void __launch_bounds__(256) k_prep5(const Prep5Params<T> p) {
  const int nd = p.n + 6;
  const int total = nd * nd;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) prep5_cell<T>(p, blockIdx.y, e);
}
// padded copy of a 2-D metric array: dst [tile][nd][PP] <- src [tile][nd][nd]
template <class T> __global__ void __launch_bounds__(256) k_pad_plane(T* __restrict__ dst, const T* __restrict__ src, int nd, int PP, int ntiles) {
  const long total = (long)ntiles * nd * PP;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int col = (int)(e % PP);
    const long row = e / PP;
    dst[e] = col < nd ? src[row * nd + col] : T(0);
  }
}
#endif

// ---------------------------------------------------------------------------------------------------------------------
// the marching group
// ---------------------------------------------------------------------------------------------------------------------
struct Adv5Cta {  // uniform over one tracer group
  int n, npx, nd, i0, nw, gb;
  int tile, tileoff;  // element offset of this tile in the Fortran-layout 2-D metric arrays
  long qoff;          // element offset of the (tile, tracer, level) plane of q
  bool xedge;         // the strip touches the west or east tile edge
};

// index shift / edge position the PPM element functions see in x and in y (A5Sub); EDGE = false: nothing looks at them
template <class T, bool EDGE, class P> FV3T_HD int a5_shift_x(const P& p, const Adv5Cta& c) {
  if constexpr (EDGE && P::SUB) return p.sub[c.tile].no_w ? A5_SHIFT : 0;
  return 0;
}
template <class T, bool EDGE, class P> FV3T_HD int a5_shift_y(const P& p, const Adv5Cta& c) {
  if constexpr (EDGE && P::SUB) return p.sub[c.tile].no_s ? A5_SHIFT : 0;
  return 0;
}
template <class T, bool EDGE, class P> FV3T_HD int a5_edge_x(const P& p, const Adv5Cta& c, int sh) {
  if constexpr (EDGE && P::SUB) return p.sub[c.tile].no_e ? A5_NOEDGE : c.npx + sh;
  return c.npx;
}
template <class T, bool EDGE, class P> FV3T_HD int a5_edge_y(const P& p, const Adv5Cta& c, int sh) {
  if constexpr (EDGE && P::SUB) return p.sub[c.tile].no_n ? A5_NOEDGE : c.npx + sh;
  return c.npx;
}
template <class P> FV3T_HD bool a5_corner_on(const P& p, const Adv5Cta& c, int bit) {
  if constexpr (P::SUB) return (p.sub[c.tile].cmask & bit) != 0;
  return true;
}

template <class T, class P> FV3T_HD bool adv5_make_cta(const P& p, int strip, int levc, int iq, Adv5Cta& c) {
  const int n = p.n, npz = p.npz;
  const int lev = p.lev0 + levc;
  const int t = lev / npz, kz = lev % npz;
  if (p.it > p.ksplt[kz]) return false;
  const int nd = n + 6;
  c.n = n;
  c.npx = n + 1;
  c.nd = nd;
  c.i0 = 1 + strip * A5_W;
  c.nw = (A5_W < n - c.i0 + 1) ? A5_W : n - c.i0 + 1;
  c.gb = c.i0 - 3;
  c.tile = t;
  c.tileoff = t * nd * nd;
  c.qoff = (((long)t * p.nq + iq) * npz + kz) * (long)nd * nd;
  // (the flags of a sub-tile context are read from the kernel parameters where the EDGE instantiations need them -- a5_shift_* --
  // and not carried in registers: the interior blocks, which never look at them, are the ones that set the register budget)
  // x-faces i0 .. i0+nw evaluate cells i0-1 .. i0+nw; the tile-edge formulas apply to cells <= 2 and >= npx-2
  bool w_edge = true, e_edge = true;
  if constexpr (P::SUB) {
    w_edge = !p.sub[t].no_w;
    e_edge = !p.sub[t].no_e;
  }
  c.xedge = (w_edge && c.i0 - 1 <= 2) || (e_edge && c.i0 + c.nw >= c.npx - 2);
  return true;
}

template <class T, int OI, int OO> struct Adv5State {
  YWin<T, OI> yin;
  YWin<T, OO> you;
  T Fy_prev, fys_prev;  // yfx*fy2 / (fy+fy2)*mfy at the previous y-face
  const T* qg;          // this thread's (clamped) column of q, row -2
  T* qo;
  T* smt;               // this thread's slot of exchange row 0
  T* qys;               // this thread's private slot 0 (A5V_*)
};

// per-thread view of the staged box of the current block of four row steps
template <class T> struct A5View {
  const Pair<T>* sp;
  const T* ss;
};
#define A5P(v, f, k) ((v).sp[((f) * A5_R + (k)) * A5_GW])
#define A5S(v, f, k) ((v).ss[(f) * A5Stage<T>::SFIELD + (k) * A5Stage<T>::SW])
#define A5XROW(s, k) ((s).smt + (k) * A5_XP)

// xs = plane column of the group's thread 0 (strip * A5_W)
template <class T> FV3T_HD A5View<T> a5_view(const void* stage, int gtid, int xs) {
  A5View<T> v;
  v.sp = reinterpret_cast<const Pair<T>*>(stage) + gtid;
  v.ss = reinterpret_cast<const T*>(reinterpret_cast<const unsigned char*>(stage) + A5Stage<T>::PAIR_BYTES) + gtid + A5Stage<T>::shift(xs);
  return v;
}

FV3T_HD Adv3Thr adv5_thread(const Adv5Cta& c, int tid) {
  Adv3Thr t;
  const int n = c.n;
  t.tid = tid;
  t.i = c.gb + tid;
  const int ic = t.i > n + 3 ? n + 3 : t.i;
  t.pix = ic + 2;
  t.imx = t.imy = 0;
  t.cell = (tid >= 3 && tid < c.nw + 3) ? 1 : 0;
  t.icor = (t.i < 1 || t.i > n) ? 1 : 0;
  keep(t.i);
  keep(t.pix);
  keep(t.cell);
  keep(t.icor);
  return t;
}

template <class T, int OI, int OO, class P>
FV3T_HD void adv5_init(const P& p, const Adv5Cta& c, const Adv3Thr& t, T* group_smem, Adv5State<T, OI, OO>& s) {
  s.yin.init();
  s.you.init();
  s.Fy_prev = s.fys_prev = T(0);
  s.qg = p.qin + c.qoff + t.pix;
  s.qo = p.qout + c.qoff + t.pix;
  keep_ptr(s.qg);
  keep_ptr(s.qo);
  s.smt = group_smem + A5_XPAD + t.tid;
  s.qys = group_smem + A5_XROWS * A5_XP + t.tid;
  for (int k = 0; k < A5Stage<T>::PRIV; ++k) s.qys[k * A5_GW] = T(0);
}

// asynchronous copy of q(i, r) into its exchange row (slot PH = row step mod 4); CORNER: rows outside 1..n also fetch the dir = 2
// view into the private slot (the x sweeps see the dir = 1 corner view of q, the y sweeps the dir = 2 view: copy_corners,
// tp_core.F90:265-328)
template <class T, int OI, int OO, int PH, bool CORNER, class P>
FV3T_HD void adv5_issue_q(const P& p, const Adv5Cta& c, Adv5State<T, OI, OO>& s, const Adv3Thr& t, int r) {
  const int n = c.n, nd = c.nd, npx = c.npx;
  if (!CORNER) {
    async_copy<sizeof(T)>(A5XROW(s, A5X_Q + PH), s.qg + (r + 2) * nd);
  } else {
    r = r > n + 3 ? n + 3 : r;
    const int i = t.i;
    int ox = (r + 2) * nd, oy = ox;
    const int cbit = r < 1 ? (i < 1 ? 1 : 2) : (i > n ? 4 : 8);
    if (t.icor && (r < 1 || r > n) && i <= n + 3 && a5_corner_on(p, c, cbit)) {
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
      ox = (s1j + 2) * nd + (s1i - i);
      oy = (s2j + 2) * nd + (s2i - i);
    }
    async_copy<sizeof(T)>(A5XROW(s, A5X_Q + PH), s.qg + ox);
    if (r < 1 || r > n) async_copy<sizeof(T)>(s.qys + (A5V_QY + PH) * A5_GW, s.qg + oy);
  }
  async_commit();
}

// The four phases of row step r (PH = step mod 4 selects the staged box row, the exchange-row slots and the window rotation):
//   phase 1  inner y sweep (flux at y-face c = r-2), q_i of row o = r-3                          reads q row PH (own slot)
//   phase 2  dm (ORD >= 7) / al (ORD < 7) of row r (inner x sweep) and of row o (outer x sweep on q_i)
//   phase 3  x-face fluxes: inner sweep of row r (-> xfx*fx2), outer sweep of row o (-> (fx+fx2)*mfx)
//   phase 4  q_j of row r, outer y sweep (flux at y-face c), flux-form update of row o
// Dependencies inside a step: 1 -> 2 -> 3 -> 4 through the exchange rows.  Phases 1 and 4 (y direction, register-resident
// windows) never touch what phases 2 and 3 (x direction, shared-memory rows) write in the same interval, which is what the
// software-pipelined schedule of adv5_block uses.
// EX selects the reference's own operation order (divisions by ra_x / ra_y / dp2, no shared reciprocals); instantiated in a
// translation unit built with -fmad=false it is bit-identical to the FMA-free oracle (and to the strict k_advect2).
template <class T, int OI, int OO, int PH, bool YE, bool EX = false, class P>
FV3T_HD void adv5_phase1(const P& p, const Adv5Cta& c, Adv5State<T, OI, OO>& s, const Adv3Thr& t, const A5View<T>& v, int r) {
  const int nd = c.nd;
  const int cc = r - 2;
  const Pair<T> y2 = A5P(v, A5_Y2, PH);  // zero outside the faces 1..n+1 (k_prep5, TMA zero fill)
  const T* dya = p.dya + c.tileoff + t.pix;
  const int shy = a5_shift_y<T, YE>(p, c);
  auto met_y = [&](int row) -> T { return dya[(row - shy + 2) * nd]; };
  T qy = A5XROW(s, A5X_Q + PH)[0];
  if (YE && (r < 1 || r > c.n)) qy = s.qys[(A5V_QY + PH) * A5_GW];
  const T q_o = s.yin.template q_cm1<PH>();
  const T fy2_c = s.yin.template push<PH, YE>(cc + shy, qy, y2.a, a5_edge_y<T, YE>(p, c, shy), p.lim_fac, met_y, s.qys + A5V_EIN * A5_GW, A5_GW);
  const T Fy_c = y2.b * fy2_c;
  T qi;  // only rows o = 1..n are consumed
  if (EX) {
    const T ray = A5S(v, A5_RY, PH);  // zero outside rows 1..n: keep the quotient finite there
    qi = (q_o * A5S(v, A5_AO, PH) + s.Fy_prev - Fy_c) / (ray != T(0) ? ray : T(1));
  } else {
    qi = (q_o * A5S(v, A5_AO, PH) + s.Fy_prev - Fy_c) * A5S(v, A5_RY, PH);
  }
  s.Fy_prev = Fy_c;
  s.qys[(A5V_FY2 + (PH & 1)) * A5_GW] = fy2_c;  // phase 4 of this step runs after phase 1 of the next one
  s.qys[(A5V_QO + (PH & 1)) * A5_GW] = q_o;
  A5XROW(s, A5X_QI + (PH & 1))[0] = qi;
}

template <class T, int OI, int OO, int PH, bool XE, class P>
FV3T_HD void adv5_phase2(const P& p, const Adv5Cta& c, Adv5State<T, OI, OO>& s, const Adv3Thr& t, int r) {
  const int nd = c.nd;
  const int rr = XE ? clampi(r, -2, c.n + 3) : r;
  const int o = XE ? clampi(r - 3, 1, c.n) : r - 3;
  const int shx = a5_shift_x<T, XE>(p, c), nxx = a5_edge_x<T, XE>(p, c, shx);
  const int i = t.i + shx;  // the column index the PPM functions see (A5Sub); the accessors take the same shifted index
  const T* dxa = p.dxa + c.tileoff + 2 - shx;
  auto dxa_r = [&](int gi) -> T { return dxa[(rr + 2) * nd + gi]; };
  auto dxa_o = [&](int gi) -> T { return dxa[(o + 2) * nd + gi]; };
  const T* sqa = A5XROW(s, A5X_Q + PH) - i;  // indexable by the (shifted) column
  const T* sqb = A5XROW(s, A5X_QI + (PH & 1)) - i;
  auto qa = [&](int gi) -> T { return sqa[gi]; };
  auto qb = [&](int gi) -> T { return sqb[gi]; };
  A5XROW(s, A5X_DR)[0] = ppm_pre<T, OI, XE>(i, nxx, qa, dxa_r);
  A5XROW(s, A5X_DO)[0] = ppm_pre<T, OO, XE>(i, nxx, qb, dxa_o);
}

template <class T, int OI, int OO, int PH, bool XE, bool EX = false, class P>
FV3T_HD void adv5_phase3(const P& p, const Adv5Cta& c, Adv5State<T, OI, OO>& s, const Adv3Thr& t, const A5View<T>& v, int r) {
  const int nd = c.nd;
  const int rr = XE ? clampi(r, -2, c.n + 3) : r;
  const int o = XE ? clampi(r - 3, 1, c.n) : r - 3;
  const int shx = a5_shift_x<T, XE>(p, c), nxx = a5_edge_x<T, XE>(p, c, shx);
  const int i = t.i + shx;
  const T* dxa = p.dxa + c.tileoff + 2 - shx;
  auto dxa_r = [&](int gi) -> T { return dxa[(rr + 2) * nd + gi]; };
  auto dxa_o = [&](int gi) -> T { return dxa[(o + 2) * nd + gi]; };
  const T *sqa = A5XROW(s, A5X_Q + PH) - i, *sda = A5XROW(s, A5X_DR) - i, *sqb = A5XROW(s, A5X_QI + (PH & 1)) - i, *sdb = A5XROW(s, A5X_DO) - i;
  auto qa = [&](int gi) -> T { return sqa[gi]; };
  auto aa = [&](int gi) -> T { return sda[gi]; };
  auto qb = [&](int gi) -> T { return sqb[gi]; };
  auto ab = [&](int gi) -> T { return sdb[gi]; };
  const Pair<T> x2r = A5P(v, A5_XR, PH);
  const T fx2 = xface_flux<T, OI, XE>(i, x2r.a, nxx, p.lim_fac, qa, aa, dxa_r);
  A5XROW(s, A5X_SF1)[0] = x2r.b * fx2;
  const T fxo = xface_flux<T, OO, XE>(i, A5P(v, A5_XO, PH).a, nxx, p.lim_fac, qb, ab, dxa_o);
  const T fx2o = s.qys[(A5V_FX2 + ((PH + 1) & 3)) * A5_GW];  // fx2 of row r-3, stored three row steps ago
  A5XROW(s, A5X_SFT)[0] = EX ? T(0.5) * (fxo + fx2o) * A5S(v, A5_MFX, PH) : (fxo + fx2o) * A5S(v, A5_MFX, PH);
  s.qys[(A5V_FX2 + PH) * A5_GW] = fx2;
}

template <class T, int OI, int OO, int PH, bool YE, bool EX = false, class P>
FV3T_HD void adv5_phase4(const P& p, const Adv5Cta& c, Adv5State<T, OI, OO>& s, const Adv3Thr& t, const A5View<T>& v, int r) {
  const int n = c.n, nd = c.nd;
  const int cc = r - 2, o = r - 3;
  const T* dya = p.dya + c.tileoff + t.pix;
  const int shy = a5_shift_y<T, YE>(p, c);
  auto met_y = [&](int row) -> T { return dya[(row - shy + 2) * nd]; };
  const T* sf1 = A5XROW(s, A5X_SF1);
  const T* sft = A5XROW(s, A5X_SFT);
  const T qx = A5XROW(s, A5X_Q + PH)[0];
  const T q_o = s.qys[(A5V_QO + (PH & 1)) * A5_GW];
  T qj;
  if (EX) {
    const T rax = A5S(v, A5_RX, PH);
    qj = (qx * A5S(v, A5_AR, PH) + sf1[0] - sf1[1]) / (rax != T(0) ? rax : T(1));
  } else {
    qj = (qx * A5S(v, A5_AR, PH) + sf1[0] - sf1[1]) * A5S(v, A5_RX, PH);
  }
  const T fyo_c = s.you.template push<PH, YE>(cc + shy, qj, A5P(v, A5_Y2, PH).a, a5_edge_y<T, YE>(p, c, shy), p.lim_fac, met_y, s.qys + A5V_EOU * A5_GW, A5_GW);
  const T fy2c = s.qys[(A5V_FY2 + (PH & 1)) * A5_GW];
  const T fys_c = EX ? T(0.5) * (fyo_c + fy2c) * A5S(v, A5_MFY, PH) : (fyo_c + fy2c) * A5S(v, A5_MFY, PH);  // mfy: zero outside faces 1..n+1
  const Pair<T> ab = A5P(v, A5_CAB, PH);
  T qnew;
  if (EX)  // {dp1, dp2}: (q dp1 + (fx(i) - fx(i+1) + fy(j) - fy(j+1)) rarea) / dp2, fv_tracer2d.F90:538-542
    qnew = (q_o * ab.a + (sft[0] - sft[1] + s.fys_prev - fys_c) * A5S(v, A5_RA, PH)) / (ab.b != T(0) ? ab.b : T(1));
  else
    qnew = q_o * ab.a + (sft[0] - sft[1] + s.fys_prev - fys_c) * ab.b;
  const bool o_ok = !YE || (o >= 1 && o <= n);
  if (o_ok && t.cell) s.qo[(o + 2) * nd] = qnew;
  s.fys_prev = fys_c;
}

// number of four-row blocks: rows -2 .. n+3 plus at least one padding step (the schedule runs phase 4 one step late)
FV3T_HD int a5_nblocks(int n) { return (n + 6) / A5_R + 1; }
// a block whose phases all see interior cells c = r-2 (3 <= c <= npx-3) and corner-free q rows: phase 4 runs for rows r0-1 ..
// r0+2, phases 2 and 3 for r0 .. r0+3, phase 1 for r0+1 .. r0+4, q is requested for r0+2 .. r0+5
FV3T_HD bool a5_block_interior(int r0, int n) { return r0 >= 6 && r0 + 5 <= n; }

// host restatement of what the producer warp's TMA boxes deliver for block b of (strip, level): test infrastructure
// (tests/hostsim) and documentation of the box coordinates in one place
template <class T> FV3T_HD void a5_box_rows(int b, int* row_of /*[A5_NPAIR + A5_NSC]*/) {
  const int r0 = -2 + A5_R * b;  // first row step of the block; plane row = Fortran row + 2
  row_of[A5_XR] = r0 + 2;        // row r
  row_of[A5_XO] = r0 - 1;        // row o = r-3
  row_of[A5_Y2] = r0;            // row c = r-2
  row_of[A5_CAB] = r0 - 1;
  row_of[A5_NPAIR + A5_RX] = r0 + 2;
  row_of[A5_NPAIR + A5_MFX] = r0 - 1;
  row_of[A5_NPAIR + A5_RY] = r0 - 1;
  row_of[A5_NPAIR + A5_MFY] = r0;
  row_of[A5_NPAIR + A5_AR] = r0 + 2;
  row_of[A5_NPAIR + A5_AO] = r0 - 1;
  row_of[A5_NPAIR + A5_RA] = r0 - 1;
}

#ifdef __CUDACC__
// ---- mbarrier / TMA primitives (PTX ISA 8.x, sm_90+) ------------------------------------------------------------------
__device__ __forceinline__ unsigned a5_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void a5_mbar_init(uint64_t* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a5_smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void a5_mbar_expect_tx(uint64_t* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a5_smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void a5_mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a5_smem_u32(b)) : "memory");
}
__device__ __forceinline__ void a5_mbar_wait(uint64_t* b, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "A5_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra A5_DONE_%=;\n"
      "bra A5_WAIT_%=;\n"
      "A5_DONE_%=:\n"
      "}\n" ::"r"(a5_smem_u32(b)),
      "r"(parity), "r"(0x989680u)  // suspend-time hint: let the hardware park the warp instead of spinning through try_wait
      : "memory");
}
__device__ __forceinline__ void a5_tma_3d(void* dst, const CUtensorMap* m, int x, int y, int z, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                   a5_smem_u32(dst)),
               "l"(m), "r"(a5_smem_u32(bar)), "r"(x), "r"(y), "r"(z)
               : "memory");
}
// named barrier of one tracer group (ids 1..A5_MAXTG; 0 is __syncthreads).  ONEG (one tracer per CTA): a constant id, so that
// ptxas reserves two hardware barriers instead of all sixteen and several such CTAs share an SM.
template <bool ONEG> __device__ __forceinline__ void a5_group_sync(int g) {
  if (ONEG)
    asm volatile("bar.sync 1, %0;" ::"n"(A5_GW) : "memory");
  else
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(A5_GW) : "memory");
}

// One block = four row steps, software-pipelined over two barrier intervals per step:
//     interval 1:  phase 2 of step s   |  phase 4 of step s-1        (x direction | y direction: independent instruction streams)
//     interval 2:  phase 3 of step s   |  phase 1 of step s+1
// Two barriers per row step instead of three and two independent dependency chains between them: k_advect5's first version
// (phases 1-2-3-4 in sequence, three barriers) spent 2.6 of 9.4 stall cycles per issued instruction waiting on fixed-latency
// FP64 dependencies and 0.8 at barriers (profiles/r02_advect5_v1_ncu.txt).  vp / v / vn: staged boxes of the previous, this and
// the next block (phase 4 of step r0-1 reads the previous box, phase 1 of step r0+4 the next one).
template <class T, int OI, int OO, bool YE, bool XE, bool ONEG, bool EX, class P>
__device__ __forceinline__ void adv5_block(const P& p, const Adv5Cta& c, Adv5State<T, OI, OO>& s, const Adv3Thr& t,
                                           const A5View<T>& vp, const A5View<T>& v, const unsigned char* next_stage, uint64_t* full_next,
                                           unsigned par_next, uint64_t* empty_prev, int r0, int g, int gtid) {
  // ---- step r0 (PH 0)
  adv5_issue_q<T, OI, OO, 2, YE>(p, c, s, t, r0 + 2);
  async_wait_pending<1>();  // q of row r0+1 has landed (r0 landed one step ago)
  adv5_phase2<T, OI, OO, 0, XE>(p, c, s, t, r0);
  if (!YE || r0 > -2) adv5_phase4<T, OI, OO, 3, YE, EX>(p, c, s, t, vp, r0 - 1);
  if (empty_prev) {  // the previous box is free for the producer
    __syncwarp();
    if ((gtid & 31) == 0) a5_mbar_arrive(empty_prev);
  }
  a5_group_sync<ONEG>(g);
  adv5_phase3<T, OI, OO, 0, XE, EX>(p, c, s, t, v, r0);
  adv5_phase1<T, OI, OO, 1, YE, EX>(p, c, s, t, v, r0 + 1);
  a5_group_sync<ONEG>(g);
  // ---- step r0+1 (PH 1)
  adv5_issue_q<T, OI, OO, 3, YE>(p, c, s, t, r0 + 3);
  async_wait_pending<1>();
  adv5_phase2<T, OI, OO, 1, XE>(p, c, s, t, r0 + 1);
  adv5_phase4<T, OI, OO, 0, YE, EX>(p, c, s, t, v, r0);
  a5_group_sync<ONEG>(g);
  adv5_phase3<T, OI, OO, 1, XE, EX>(p, c, s, t, v, r0 + 1);
  adv5_phase1<T, OI, OO, 2, YE, EX>(p, c, s, t, v, r0 + 2);
  a5_group_sync<ONEG>(g);
  // ---- step r0+2 (PH 2)
  adv5_issue_q<T, OI, OO, 0, YE>(p, c, s, t, r0 + 4);
  async_wait_pending<1>();
  adv5_phase2<T, OI, OO, 2, XE>(p, c, s, t, r0 + 2);
  adv5_phase4<T, OI, OO, 1, YE, EX>(p, c, s, t, v, r0 + 1);
  a5_group_sync<ONEG>(g);
  adv5_phase3<T, OI, OO, 2, XE, EX>(p, c, s, t, v, r0 + 2);
  adv5_phase1<T, OI, OO, 3, YE, EX>(p, c, s, t, v, r0 + 3);
  a5_group_sync<ONEG>(g);
  // ---- step r0+3 (PH 3)
  adv5_issue_q<T, OI, OO, 1, YE>(p, c, s, t, r0 + 5);
  async_wait_pending<1>();
  adv5_phase2<T, OI, OO, 3, XE>(p, c, s, t, r0 + 3);
  adv5_phase4<T, OI, OO, 2, YE, EX>(p, c, s, t, v, r0 + 2);
  a5_group_sync<ONEG>(g);
  adv5_phase3<T, OI, OO, 3, XE, EX>(p, c, s, t, v, r0 + 3);
  if (full_next) {
    a5_mbar_wait(full_next, par_next);
    const A5View<T> vn = a5_view<T>(next_stage, gtid, c.i0 - 1);
    adv5_phase1<T, OI, OO, 0, YE, EX>(p, c, s, t, vn, r0 + 4);
  }
  a5_group_sync<ONEG>(g);
}

template <class T, int OI, int OO, int NTHR, int MINB, bool EX = false, class P = Adv5Params<T>>
__global__ void __launch_bounds__(NTHR, MINB) k_advect5(const __grid_constant__ P p, const __grid_constant__ Adv5Maps maps) {
  extern __shared__ __align__(128) unsigned char smem5[];
  using S = A5Stage<T>;
  const int strip = blockIdx.x, levc = blockIdx.y;
  const int n = p.n;
  const int lev = p.lev0 + levc;
  if (p.it > p.ksplt[lev % p.npz]) return;
  const int iq_first = p.iq0 + blockIdx.z * p.tg;
  const int tg = min(p.tg, p.iq0 + p.nql - iq_first);  // tracers of this CTA
  uint64_t* full = reinterpret_cast<uint64_t*>(smem5 + S::BAR_OFF);
  uint64_t* empty = full + A5_NS;
  if (threadIdx.x == 0) {
    for (int k = 0; k < A5_NS; ++k) {
      a5_mbar_init(&full[k], 1);
      a5_mbar_init(&empty[k], 2 * tg);  // one arrival per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const int nblocks = a5_nblocks(n);
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    // ---- producer: one thread streams the level fields of the strip through the ring
    if (threadIdx.x == 0) {
      const int xs = strip * A5_W;  // plane column of group thread 0
      const int tile = lev / p.npz;
      int st = 0, wrap = 0;
      for (int b = 0; b < nblocks; ++b) {
        if (b >= A5_NS) a5_mbar_wait(&empty[st], (wrap - 1) & 1);
        a5_mbar_expect_tx(&full[st], S::TX_BYTES);
        unsigned char* base = smem5 + st * S::BYTES;
        constexpr int PB = A5_R * A5_GW * 2 * (int)sizeof(T), SB = S::SBYTES;
        const int xq = xs - S::shift(xs);  // 16-byte aligned start of the scalar boxes
        const int r0 = -2 + A5_R * b;
        a5_tma_3d(base + A5_XR * PB, &maps.x2, 2 * xs, r0 + 2, levc, &full[st]);
        a5_tma_3d(base + A5_XO * PB, &maps.x2, 2 * xs, r0 - 1, levc, &full[st]);
        a5_tma_3d(base + A5_Y2 * PB, &maps.y2, 2 * xs, r0, levc, &full[st]);
        a5_tma_3d(base + A5_CAB * PB, &maps.cab, 2 * xs, r0 - 1, levc, &full[st]);
        unsigned char* sb = base + S::PAIR_BYTES;
        a5_tma_3d(sb + A5_RX * SB, &maps.rx, xq, r0 + 2, levc, &full[st]);
        a5_tma_3d(sb + A5_MFX * SB, &maps.mfx, xq, r0 - 1, levc, &full[st]);
        a5_tma_3d(sb + A5_RY * SB, &maps.ry, xq, r0 - 1, levc, &full[st]);
        a5_tma_3d(sb + A5_MFY * SB, &maps.mfy, xq, r0, levc, &full[st]);
        a5_tma_3d(sb + A5_AR * SB, &maps.area, xq, r0 + 2, tile, &full[st]);
        a5_tma_3d(sb + A5_AO * SB, &maps.area, xq, r0 - 1, tile, &full[st]);
        a5_tma_3d(sb + A5_RA * SB, &maps.rarea, xq, r0 - 1, tile, &full[st]);
        if (++st == A5_NS) {
          st = 0;
          ++wrap;
        }
      }
    }
    return;
  }
  // ---- consumers: group g marches tracer iq_first + g
  const int ct = threadIdx.x - 32;
  const int g = ct >> 6, gtid = ct & (A5_GW - 1);
  if (g >= tg) return;
  Adv5Cta c;
  adv5_make_cta<T>(p, strip, levc, iq_first + g, c);
  const Adv3Thr t = adv5_thread(c, gtid);
  Adv5State<T, OI, OO> s;
  T* gsm = reinterpret_cast<T*>(smem5 + S::GROUP_OFF) + (size_t)g * S::GROUP_ELEMS;
  adv5_init<T, OI, OO>(p, c, t, gsm, s);
  constexpr bool ONEG = NTHR == 32 + A5_GW;
  adv5_issue_q<T, OI, OO, 0, true>(p, c, s, t, -2);
  adv5_issue_q<T, OI, OO, 1, true>(p, c, s, t, -1);
  async_wait_all();
  a5_mbar_wait(&full[0], 0);
  const int xs = strip * A5_W;
  A5View<T> v = a5_view<T>(smem5, gtid, xs), vp = v;
  adv5_phase1<T, OI, OO, 0, true, EX>(p, c, s, t, v, -2);
  a5_group_sync<ONEG>(g);
  int st = 0, wrap = 0;
  for (int b = 0; b < nblocks; ++b) {
    const int r0 = -2 + A5_R * b;
    int sn = st + 1, wn = wrap;
    if (sn == A5_NS) {
      sn = 0;
      ++wn;
    }
    const int sp = st == 0 ? A5_NS - 1 : st - 1;
    const unsigned char* next_stage = smem5 + sn * S::BYTES;
    uint64_t* full_next = b + 1 < nblocks ? &full[sn] : nullptr;
    uint64_t* empty_prev = b > 0 ? &empty[sp] : nullptr;
    if (a5_block_interior(r0, n)) {
      if (c.xedge)
        adv5_block<T, OI, OO, false, true, ONEG, EX>(p, c, s, t, vp, v, next_stage, full_next, wn & 1, empty_prev, r0, g, gtid);
      else
        adv5_block<T, OI, OO, false, false, ONEG, EX>(p, c, s, t, vp, v, next_stage, full_next, wn & 1, empty_prev, r0, g, gtid);
    } else {
      adv5_block<T, OI, OO, true, true, ONEG, EX>(p, c, s, t, vp, v, next_stage, full_next, wn & 1, empty_prev, r0, g, gtid);
    }
    vp = v;
    v = a5_view<T>(next_stage, gtid, xs);
    st = sn;
    wrap = wn;
  }
  async_wait_all();
}
#endif

}  // namespace fv3t
