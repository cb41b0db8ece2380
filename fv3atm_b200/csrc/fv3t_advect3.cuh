// fv3atm_b200: production ("fast") horizontal tracer advection for sm_100a.
//
// Same operator and the same marching decomposition as fv3t_advect2.cuh (tracer_2d sub-step, atmos_cubed_sphere/
// model/fv_tracer2d.F90:503-556 -> fv_tp_2d, model/tp_core.F90:110-249 with xppm :332-704, yppm :707-1124,
// copy_corners :253-330, pert_ppm :1178-1236), re-balanced for the B200 instruction budget: in fp64 the path is bound
// by the FP64 pipe and by instruction issue, not by HBM (DESIGN.md "Roofline"), so everything that does not depend on
// the tracer is computed ONCE per (tile, level) by k_prep3 and only read here:
//     X2  = {cx, xfx}      at x-faces      Y2 = {cy, yfx}     at y-faces        (16-byte pairs, one LDG.128 each)
//     rrx = 1/ra_x, rry = 1/ra_y           (fv_tracer2d.F90:518-526: the two divisions of q_i / q_j become multiplies)
//     cab = {dp1/dp2, 0.5*rarea/dp2}       (:510-515, 538-542: the flux-form update becomes q*a + sum*b)
// all in one "plane layout" (isd:ied, jsd:jed) so that a single per-thread index addresses every array of a row.
// Built with FMA contraction ON.  These are <= 1-ulp-per-operation re-associations of the reference formulas: the
// results agree with the FMA-free oracle to ~1e-15 normalised (tests assert the 1e-12 bar of BASELINE.json), while the
// strict kernels (fv3t_advect2.cuh, -fmad=false, selected with FV3T_STRICT=1) remain bit-identical to it.
// Loads of row r+1 are issued one row step ahead of their use (software prefetch into registers): the strict kernel
// spent 4.6 of every 9.7 issue-slot cycles waiting on the long scoreboard (profiles/r01_advect2_c384_ncu.txt).
//
// The CTA body is written as four barrier-separated phase functions over an explicit per-thread state so that the
// identical code runs on the host, thread by thread, in tests/hostsim/ (test infrastructure; never linked here).
#pragma once
#include "fv3t_advect2.cuh"

namespace fv3t {

template <class T> struct alignas(2 * sizeof(T)) Pair {
  T a, b;
};

// products / sums that must not be contracted (dp2: the caller-visible dp1 post-state stays bit-exact)
FV3T_HD double mul_rn(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
FV3T_HD float mul_rn(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
FV3T_HD double add_rn(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
FV3T_HD float add_rn(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}

template <class T> struct Adv3Params {
  const T* qin;
  T* qout;
  const Pair<T>* X2;   // per (tile, level) plane: {cx, xfx} of x-face (i, j), i = 1..n+1, j = jsd..jed
  const Pair<T>* Y2;   // {cy, yfx} of y-face (i, j), i = isd..ied, j = 1..n+1
  const T* rrx;        // 1 / ra_x(i, j), i = 1..n, j = jsd..jed
  const T* rry;        // 1 / ra_y(i, j), i = isd..ied, j = 1..n
  const Pair<T>* cab;  // {dp1/dp2, 0.5*rarea/dp2} on the compute domain
  const T *mfx, *mfy;  // scaled mass fluxes, Fortran extents
  const T *area, *dxa, *dya;
  const int* ksplt;
  int n, npz, nq, ntiles, it, W;
  T lim_fac;
  int iq0 = 0, nql = -1;  // this launch advects tracers iq0 .. iq0+nql-1 of the nq resident ones (nql < 0: all)
};

// ---------------------------------------------------------------------------------------------------------------------
// k_prep3: the tracer-independent part of one tracer_2d call, per (tile, level)  (fv_tracer2d.F90:387-405 xfx/yfx,
// :449-486 frac scaling, :510-526 dp2 / ra_x / ra_y).  Reads the UNSCALED cx, cy, mfx, mfy; nothing is scaled in place
// here (k_scale3 does that afterwards when nsplt /= 1), so neighbouring threads never race.
// ---------------------------------------------------------------------------------------------------------------------
template <class T> struct Prep3Params {
  const T *cx, *cy, *mfx, *mfy, *dp1;
  GridDev<T> g;
  Pair<T>*X2, *Y2, *cab;
  T *rrx, *rry;
  const int* ksplt;
  int n, npz, ntiles;
};

template <class T> FV3T_HD T xfx_of(const T* cxp, const T* dxa, const T* dyg, const T* ssg, long plane, int nd, int n, int i, int j, T frac, T& c_out) {
  const T c = cxp[(long)(j + 2) * (n + 1) + (i - 1)];
  const long oc = (long)(j + 2) * nd + (i + 2), ody = (long)(j + 2) * (nd + 1) + (i + 2);
  T xf;
  if (c > T(0))
    xf = mul_rn(mul_rn(mul_rn(c, dxa[oc - 1]), dyg[ody]), ssg[2 * plane + oc - 1]);
  else
    xf = mul_rn(mul_rn(mul_rn(c, dxa[oc]), dyg[ody]), ssg[0 * plane + oc]);
  c_out = c;
  return mul_rn(xf, frac);
}
template <class T> FV3T_HD T yfx_of(const T* cyp, const T* dya, const T* dxg, const T* ssg, long plane, int nd, int n, int i, int j, T frac, T& c_out) {
  const T c = cyp[(long)(j - 1) * nd + (i + 2)];
  const long oc = (long)(j + 2) * nd + (i + 2);
  T yf;
  if (c > T(0))
    yf = mul_rn(mul_rn(mul_rn(c, dya[oc - nd]), dxg[oc]), ssg[3 * plane + oc - nd]);
  else
    yf = mul_rn(mul_rn(mul_rn(c, dya[oc]), dxg[oc]), ssg[1 * plane + oc]);
  c_out = c;
  return mul_rn(yf, frac);
}

// dp2 of cell (i,j) from dp1 and the scaled mass fluxes, in the reference's operation order
template <class T> FV3T_HD T dp2_of(T dp1v, T mx0, T mx1, T my0, T my1, T rar) {
  return add_rn(dp1v, mul_rn(add_rn(add_rn(add_rn(mx0, -mx1), my0), -my1), rar));
}

template <class T> FV3T_HD void prep3_cell(const Prep3Params<T>& p, int t, int kz, int e) {
  const int n = p.n, nd = n + 6;
  const long plane = (long)nd * nd;
  const long lev = (long)t * p.npz + kz;
  const int j = e / nd - 2, i = e % nd - 2;
  const T frac = T(1) / (T)p.ksplt[kz];
  const T* area = p.g.area + (long)t * plane;
  const T* rarea = p.g.rarea + (long)t * plane;
  const T* dxa = p.g.dxa + (long)t * plane;
  const T* dya = p.g.dya + (long)t * plane;
  const T* dxg = p.g.dx + (long)t * nd * (nd + 1);
  const T* dyg = p.g.dy + (long)t * (nd + 1) * nd;
  const T* ssg = p.g.sin_sg + (long)t * plane * 5;
  const T* cxp = p.cx + lev * (long)(n + 1) * nd;
  const T* cyp = p.cy + lev * (long)nd * (n + 1);
  const T* mxp = p.mfx + lev * (long)(n + 1) * n;
  const T* myp = p.mfy + lev * (long)n * (n + 1);
  const long o = lev * plane + e;
  Pair<T> x2{T(0), T(0)}, y2{T(0), T(0)}, ab{T(0), T(0)};
  T rx = T(0), ry = T(0);
  if (i >= 1 && i <= n + 1) {
    T c;
    const T xf = xfx_of<T>(cxp, dxa, dyg, ssg, plane, nd, n, i, j, frac, c);
    x2.a = mul_rn(c, frac);
    x2.b = xf;
    if (i <= n) {
      T c1;
      const T xf1 = xfx_of<T>(cxp, dxa, dyg, ssg, plane, nd, n, i + 1, j, frac, c1);
      rx = T(1) / add_rn(add_rn(area[e], xf), -xf1);
    }
  }
  if (j >= 1 && j <= n + 1) {
    T c;
    const T yf = yfx_of<T>(cyp, dya, dxg, ssg, plane, nd, n, i, j, frac, c);
    y2.a = mul_rn(c, frac);
    y2.b = yf;
    if (j <= n) {
      T c1;
      const T yf1 = yfx_of<T>(cyp, dya, dxg, ssg, plane, nd, n, i, j + 1, frac, c1);
      ry = T(1) / add_rn(add_rn(area[e], yf), -yf1);
    }
  }
  if (i >= 1 && i <= n && j >= 1 && j <= n) {
    const long ox = (long)(j - 1) * (n + 1) + (i - 1), oy = (long)(j - 1) * n + (i - 1);
    const T rar = rarea[e];
    const T d1 = p.dp1[o];
    const T d2 = dp2_of<T>(d1, mul_rn(mxp[ox], frac), mul_rn(mxp[ox + 1], frac), mul_rn(myp[oy], frac), mul_rn(myp[oy + n], frac), rar);
    const T r2 = T(1) / d2;
    ab.a = d1 * r2;
    ab.b = T(0.5) * rar * r2;
  }
  p.X2[o] = x2;
  p.Y2[o] = y2;
  p.rrx[o] = rx;
  p.rry[o] = ry;
  p.cab[o] = ab;
}

#ifdef __CUDACC__
template <class T> __global__ void __launch_bounds__(256) k_prep3(const Prep3Params<T> p) {
  const int nd = p.n + 6;
  const int lev = blockIdx.y;
  const int t = lev / p.npz, kz = lev % p.npz;
  const int total = nd * nd;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) prep3_cell<T>(p, t, kz, e);
}

// in-place 1/ksplt scaling of cx, cy, mfx, mfy: the post-state the caller sees (fv_tracer2d.F90:463-481); nsplt /= 1 only
template <class T>
__global__ void __launch_bounds__(256) k_scale3(T* __restrict__ cx, T* __restrict__ cy, T* __restrict__ mfx, T* __restrict__ mfy,
                                                const int* __restrict__ ksplt, int n, int npz) {
  const long ncx = (long)(n + 1) * (n + 6), nmf = (long)(n + 1) * n;
  const int lev = blockIdx.y;
  const T frac = T(1) / (T)ksplt[lev % npz];
  T* a = cx + lev * ncx;
  T* b = cy + lev * ncx;
  T* c = mfx + lev * nmf;
  T* d = mfy + lev * nmf;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < ncx; e += (long)gridDim.x * blockDim.x) {
    a[e] = mul_rn(a[e], frac);
    b[e] = mul_rn(b[e], frac);
    if (e < nmf) {
      c[e] = mul_rn(c[e], frac);
      d[e] = mul_rn(d[e], frac);
    }
  }
}
#endif

// Between sub-steps (it >= 2): dp1 <- dp2 for the levels that took part in sub-step it-1 (fv_tracer2d.F90:547-553),
// then cab of sub-step `it` for the levels still active.  mfx, mfy are the scaled arrays here.
template <class T> struct Cab3Params {
  T* dp1;
  const T *mfx, *mfy, *rarea;
  Pair<T>* cab;
  const int* ksplt;
  int n, npz, it;
  int mode_1l = 0;  // tracer_2d_1L: dp1 <- dp2 only between the level's own sub-steps (fv_tracer2d.F90:305)
};
template <class T> FV3T_HD void cab3_cell(const Cab3Params<T>& p, int t, int kz, int e) {
  const int n = p.n, nd = n + 6;
  const long plane = (long)nd * nd;
  const int j = e / n + 1, i = e % n + 1;
  const int ks = p.ksplt[kz];
  if (p.it - 1 > ks) return;
  if (p.mode_1l && p.it > ks) return;
  const long lev = (long)t * p.npz + kz;
  const long oc = (long)(j + 2) * nd + (i + 2);
  const long ox = (long)(j - 1) * (n + 1) + (i - 1), oy = (long)(j - 1) * n + (i - 1);
  const T* mx = p.mfx + lev * (long)(n + 1) * n;
  const T* my = p.mfy + lev * (long)n * (n + 1);
  const T rar = p.rarea[(long)t * plane + oc];
  const T d1 = dp2_of<T>(p.dp1[lev * plane + oc], mx[ox], mx[ox + 1], my[oy], my[oy + n], rar);
  p.dp1[lev * plane + oc] = d1;
  if (p.it > ks) return;
  const T d2 = dp2_of<T>(d1, mx[ox], mx[ox + 1], my[oy], my[oy + n], rar);
  const T r2 = T(1) / d2;
  p.cab[lev * plane + oc] = Pair<T>{d1 * r2, T(0.5) * rar * r2};
}
#ifdef __CUDACC__
template <class T> __global__ void __launch_bounds__(256) k_cab3(const Cab3Params<T> p) {
  const int lev = blockIdx.y;
  const int t = lev / p.npz, kz = lev % p.npz;
  const int total = p.n * p.n;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) cab3_cell<T>(p, t, kz, e);
}
#endif

// ---------------------------------------------------------------------------------------------------------------------
// The marching CTA: one strip of W columns (+3 halo columns either side) of one (tile, level, tracer).
//
// Addressing is arranged for few instructions per access: every per-level array is indexed on the kernel-parameter
// pointer itself with ONE 32-bit running index L = lev*plane + (r+2)*nd + (i+2) (fits 31 bits up to C768 L127), the 2-D
// metric arrays with A = tile*plane + ..., so an access is one IMAD.WIDE off the constant bank + the load; only q (whose
// tracer offset exceeds 32 bits) keeps a per-thread 64-bit pointer.  Shared-memory rows have a fixed pitch so that
// neighbour offsets are immediates off one per-thread pointer.  `keep()` makes these few values opaque: without it the
// compiler re-derives them from blockIdx / threadIdx in every row step (25 % of all executed instructions in the first
// version, profiles/r01_advect3_v0_lines.txt).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int SMP = 264;   // shared-memory row pitch: the largest block size + 4 pad slots either side
constexpr int SMPAD = 4;

FV3T_HD void keep(int& x) {
#ifdef __CUDA_ARCH__
  asm volatile("" : "+r"(x));
#else
  (void)x;
#endif
}
template <class P> FV3T_HD void keep_ptr(P*& x) {
#ifdef __CUDA_ARCH__
  asm volatile("" : "+l"(x));
#else
  (void)x;
#endif
}

template <class T, int G> struct Adv3Cta {  // CTA-uniform
  int n, npx, nd, i0, nw, gb;
  int levoff, tileoff, mxoff, myoff;  // element offsets of this (tile, level) in the per-level / 2-D / mfx / mfy arrays
  long qoff[G];                       // element offset of the (tile, tracer, level) plane of q, per tracer of the group
  bool live[G];                       // tracer iq0 + g exists (the last group of an nq not divisible by G is padded)
};

template <class T, int G> FV3T_HD bool adv3_make_cta(const Adv3Params<T>& p, int iqg, int strip, int lev, Adv3Cta<T, G>& c) {
  const int n = p.n, npz = p.npz;
  const int t = lev / npz, kz = lev % npz;
  if (p.it > p.ksplt[kz]) return false;
  const int nd = n + 6;
  const int plane = nd * nd;
  c.n = n;
  c.npx = n + 1;
  c.nd = nd;
  c.i0 = 1 + strip * p.W;
  c.nw = (p.W < n - c.i0 + 1) ? p.W : n - c.i0 + 1;
  c.gb = c.i0 - 3;
  c.levoff = lev * plane;
  c.tileoff = t * plane;
  c.mxoff = lev * (n + 1) * n;
  c.myoff = lev * n * (n + 1);
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int iq = p.iq0 + iqg * G + g;
    c.live[g] = iq < p.iq0 + (p.nql < 0 ? p.nq : p.nql);
    c.qoff[g] = (((long)t * p.nq + (c.live[g] ? iq : p.nq - 1)) * npz + kz) * (long)plane;
  }
  return true;
}

// Per-thread constants.  Every thread of the CTA executes every phase unconditionally -- threads outside their role
// compute on clamped (always valid) addresses and their results are never stored -- because role-dependent early exits
// made the compiler merge the whole carried state with register moves at every join (a quarter of all executed
// instructions in the second version, profiles/r01_advect3_block_sweep.txt).
struct Adv3Thr {
  int tid, i;
  int cell;  // this thread owns a compute column of the strip: its flux-form update is stored
  int icor;  // the column lies in the west / east halo (corner views of q differ between the x and y sweeps)
  int pix;   // clamped in-plane column offset (i + 2)
  int imx;   // clamped column offset in an mfx row, i - 1 in 0..n
  int imy;   // clamped column offset in an mfy row, i - 1 in 0..n-1
};

template <class T, int G> FV3T_HD Adv3Thr adv3_thread(const Adv3Cta<T, G>& c, int tid) {
  Adv3Thr t;
  const int n = c.n;
  t.tid = tid;
  t.i = c.gb + tid;
  const int ic = t.i > n + 3 ? n + 3 : t.i;
  t.pix = ic + 2;
  t.imx = (ic < 1 ? 1 : (ic > n + 1 ? n + 1 : ic)) - 1;
  t.imy = (ic < 1 ? 1 : (ic > n ? n : ic)) - 1;
  t.cell = (tid >= 3 && tid < c.nw + 3) ? 1 : 0;
  t.icor = (t.i < 1 || t.i > n) ? 1 : 0;
  keep(t.i);
  keep(t.pix);
  keep(t.imx);
  keep(t.imy);
  keep(t.cell);
  keep(t.icor);
  return t;
}

// G tracers of one (tile, level, strip) march in lock-step in one thread: the level fields, their addresses, the row
// bookkeeping, the upwind selections and the three barriers per row step are paid once per G tracers, and G independent
// dependency chains hide the FP64 latency.
template <class T, int OI, int OO, int G> struct Adv3State {
  struct PerTracer {
    YStream<T, OI> yin;
    YStream<T, OO> you;
    T fx2_a, fx2_b, fx2_c;  // inner x flux of this thread's face at rows r-1, r-2, r-3
    T Fy_prev, fys_prev;    // yfx*fy2 / (fy+fy2)*mfy at the previous y-face
    T fy2_c;                // inner y flux of the current row step (crosses its barriers)
    T in_qx, in_qy;         // phase-1 input
    const T* qg;            // this thread's (clamped) column of q, row -2
    T* qo;
  } tr[G];
  T cyv;
  // tracer-independent inputs of the phases, each (re)loaded for the NEXT row step at the end of the phase that has just
  // consumed it, so that every global load has a whole row step of latency tolerance and needs neither a second register
  // nor a move
  T in_area_o, in_rry;            // phase 1
  Pair<T> in_y2;
  Pair<T> in_x2r;                 // phase 3
  T in_cxo, in_mfx;
  T in_area_r, in_rrx, in_mfy;    // phase 4
  Pair<T> in_ab;
  T* smt;  // shared memory: this thread's slot of row 0 (rows are SMP apart; tracer g uses rows 6g .. 6g+5)
};

template <class T, int OI, int OO, int G>
FV3T_HD void adv3_init(const Adv3Params<T>& p, const Adv3Cta<T, G>& c, const Adv3Thr& t, T* smem, Adv3State<T, OI, OO, G>& s) {
#pragma unroll
  for (int g = 0; g < G; ++g) {
    auto& a = s.tr[g];
    a.yin.init();
    a.you.init();
    a.fx2_a = a.fx2_b = a.fx2_c = T(0);
    a.Fy_prev = a.fys_prev = T(0);
    a.fy2_c = T(0);
    a.in_qx = a.in_qy = T(0);
    a.qg = p.qin + c.qoff[g] + t.pix;
    a.qo = p.qout + c.qoff[g] + t.pix;
    keep_ptr(a.qg);
    keep_ptr(a.qo);
  }
  s.cyv = T(0);
  s.in_area_o = s.in_rry = T(0);
  s.in_y2 = s.in_x2r = s.in_ab = Pair<T>{T(0), T(0)};
  s.in_cxo = s.in_mfx = s.in_area_r = s.in_rrx = s.in_mfy = T(0);
  s.smt = smem + SMPAD + t.tid;  // not made opaque: the compiler must keep seeing a shared-memory address (LDS/STS)
}

// shared-memory rows (each SMP elements) of tracer g: 0 q of row r, 1 dm/al of row r, 2 q_i of row o, 3 dm/al of row o,
// 4 xfx*fx2 of row r, 5 (fx+fx2)*mfx of row o
#define FV3T_SROW(s, g, k) ((s).smt + ((g) * 6 + (k)) * SMP)

FV3T_HD int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

// Keeps the re-loads of a phase's inputs BEHIND the last use of their previous values.  Without it the scheduler hoists
// the loads into the phase, has to give them fresh registers, and copies them into the loop-carried registers at the loop
// back-edge -- a register move that waits for the load right away and so defeats the one-row-step latency tolerance
// (12 % of all stall samples sat on that move, profiles/r01_advect3_c384_ncu.txt).
FV3T_HD void load_fence() {
#ifdef __CUDA_ARCH__
  asm volatile("" ::: "memory");
#endif
}

// loads for phase 1 of row step r (rows beyond the tile are clamped: their values are never used)
template <class T, int OI, int OO, int G>
FV3T_HD void adv3_fetch1(const Adv3Params<T>& p, const Adv3Cta<T, G>& c, Adv3State<T, OI, OO, G>& s, const Adv3Thr& t, int r) {
  const int n = c.n, nd = c.nd, npx = c.npx;
  r = r > n + 3 ? n + 3 : r;
  const int i = t.i;
  const int cc = clampi(r - 2, 1, n + 1), o = clampi(r - 3, 1, n);
  // the x sweeps see the dir = 1 corner view of q, the y sweeps the dir = 2 view (copy_corners, tp_core.F90:265-328)
  int ox = (r + 2) * nd, oy = ox;
  if (t.icor && (r < 1 || r > n) && i <= n + 3) {
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
#pragma unroll
  for (int g = 0; g < G; ++g) {
    s.tr[g].in_qx = s.tr[g].qg[ox];
    s.tr[g].in_qy = s.tr[g].qg[oy];
  }
  s.in_y2 = p.Y2[c.levoff + (cc + 2) * nd + t.pix];
  s.in_area_o = p.area[c.tileoff + (o + 2) * nd + t.pix];
  s.in_rry = p.rry[c.levoff + (o + 2) * nd + t.pix];
}

// loads for phase 3 of row step r
template <class T, int OI, int OO, int G>
FV3T_HD void adv3_fetch3(const Adv3Params<T>& p, const Adv3Cta<T, G>& c, Adv3State<T, OI, OO, G>& s, const Adv3Thr& t, int r) {
  const int n = c.n, nd = c.nd;
  r = r > n + 3 ? n + 3 : r;
  const int o = clampi(r - 3, 1, n);
  s.in_x2r = p.X2[c.levoff + (r + 2) * nd + t.pix];
  s.in_cxo = p.X2[c.levoff + (o + 2) * nd + t.pix].a;
  s.in_mfx = p.mfx[c.mxoff + (o - 1) * (n + 1) + t.imx];
}

// loads for phase 4 of row step r
template <class T, int OI, int OO, int G>
FV3T_HD void adv3_fetch4(const Adv3Params<T>& p, const Adv3Cta<T, G>& c, Adv3State<T, OI, OO, G>& s, const Adv3Thr& t, int r) {
  const int n = c.n, nd = c.nd;
  r = r > n + 3 ? n + 3 : r;
  const int cc = clampi(r - 2, 1, n + 1), o = clampi(r - 3, 1, n);
  s.in_area_r = p.area[c.tileoff + (r + 2) * nd + t.pix];
  s.in_rrx = p.rrx[c.levoff + (r + 2) * nd + t.pix];
  s.in_mfy = p.mfy[c.myoff + (cc - 1) * n + t.imy];
  s.in_ab = p.cab[c.levoff + (o + 2) * nd + t.pix];
}

// phase 1: inner y sweep (flux at y-face c = r-2), q_i of row o = r-3, rows of q / q_i to shared memory
template <class T, int OI, int OO, int G>
FV3T_HD void adv3_phase1(const Adv3Params<T>& p, const Adv3Cta<T, G>& c, Adv3State<T, OI, OO, G>& s, const Adv3Thr& t, int r) {
  const int n = c.n, nd = c.nd;
  const int cc = r - 2;
  const bool c_ok = cc >= 1 && cc <= n + 1;
  const T cyv = c_ok ? s.in_y2.a : T(0), yfv = c_ok ? s.in_y2.b : T(0);
  const T* dya = p.dya + c.tileoff + t.pix;
  auto met_y = [&](int row) -> T { return dya[(row + 2) * nd]; };
#pragma unroll
  for (int g = 0; g < G; ++g) {
    auto& a = s.tr[g];
    const T qx = a.in_qx, qy = a.in_qy;
    T q_o;
    const T fy2_c = a.yin.push(cc, qy, cyv, c.npx, p.lim_fac, met_y, q_o);
    const T Fy_c = yfv * fy2_c;
    const T qi = (q_o * s.in_area_o + a.Fy_prev - Fy_c) * s.in_rry;  // only rows o = 1..n are consumed
    a.Fy_prev = Fy_c;
    a.fy2_c = fy2_c;
    FV3T_SROW(s, g, 0)[0] = qx;
    FV3T_SROW(s, g, 2)[0] = qi;
  }
  s.cyv = cyv;
  load_fence();
  adv3_fetch1<T, OI, OO, G>(p, c, s, t, r + 1);
}

// phase 2: dm (ORD >= 7) or al (ORD < 7) of row r (inner x sweep) and of row o (outer x sweep on q_i)
template <class T, int OI, int OO, int G>
FV3T_HD void adv3_phase2(const Adv3Params<T>& p, const Adv3Cta<T, G>& c, Adv3State<T, OI, OO, G>& s, const Adv3Thr& t, int r) {
  const int nd = c.nd;
  const int o = clampi(r - 3, 1, c.n);
  const int i = t.i;
  const T* dxa = p.dxa + c.tileoff + 2;
  auto dxa_r = [&](int gi) -> T { return dxa[(r + 2) * nd + gi]; };
  auto dxa_o = [&](int gi) -> T { return dxa[(o + 2) * nd + gi]; };
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const T* sqa = FV3T_SROW(s, g, 0) - i;  // indexable by the global column
    const T* sqb = FV3T_SROW(s, g, 2) - i;
    auto qa = [&](int gi) -> T { return sqa[gi]; };
    auto qb = [&](int gi) -> T { return sqb[gi]; };
    FV3T_SROW(s, g, 1)[0] = ppm_pre<T, OI>(i, c.npx, qa, dxa_r);
    FV3T_SROW(s, g, 3)[0] = ppm_pre<T, OO>(i, c.npx, qb, dxa_o);
  }
}

// phase 3: x-face fluxes: inner sweep of row r (-> xfx*fx2), outer sweep of row o (-> (fx+fx2)*mfx)
template <class T, int OI, int OO, int G>
FV3T_HD void adv3_phase3(const Adv3Params<T>& p, const Adv3Cta<T, G>& c, Adv3State<T, OI, OO, G>& s, const Adv3Thr& t, int r) {
  const int nd = c.nd;
  const int o = clampi(r - 3, 1, c.n);
  const int i = t.i;
  const T* dxa = p.dxa + c.tileoff + 2;
  auto dxa_r = [&](int gi) -> T { return dxa[(r + 2) * nd + gi]; };
  auto dxa_o = [&](int gi) -> T { return dxa[(o + 2) * nd + gi]; };
#pragma unroll
  for (int g = 0; g < G; ++g) {
    auto& a = s.tr[g];
    const T *sqa = FV3T_SROW(s, g, 0) - i, *sda = FV3T_SROW(s, g, 1) - i, *sqb = FV3T_SROW(s, g, 2) - i, *sdb = FV3T_SROW(s, g, 3) - i;
    auto qa = [&](int gi) -> T { return sqa[gi]; };
    auto aa = [&](int gi) -> T { return sda[gi]; };
    auto qb = [&](int gi) -> T { return sqb[gi]; };
    auto ab = [&](int gi) -> T { return sdb[gi]; };
    const T fx2 = xface_flux<T, OI>(i, s.in_x2r.a, c.npx, p.lim_fac, qa, aa, dxa_r);
    FV3T_SROW(s, g, 4)[0] = s.in_x2r.b * fx2;
    const T fxo = xface_flux<T, OO>(i, s.in_cxo, c.npx, p.lim_fac, qb, ab, dxa_o);
    FV3T_SROW(s, g, 5)[0] = (fxo + a.fx2_c) * s.in_mfx;
    a.fx2_c = a.fx2_b;
    a.fx2_b = a.fx2_a;
    a.fx2_a = fx2;
  }
  load_fence();
  adv3_fetch3<T, OI, OO, G>(p, c, s, t, r + 1);
}

// phase 4: q_j of row r, outer y sweep (flux at y-face c), flux-form update of row o
template <class T, int OI, int OO, int G>
FV3T_HD void adv3_phase4(const Adv3Params<T>& p, const Adv3Cta<T, G>& c, Adv3State<T, OI, OO, G>& s, const Adv3Thr& t, int r) {
  const int n = c.n, nd = c.nd;
  const int cc = r - 2, o = r - 3;
  const bool o_ok = o >= 1 && o <= n;
  const bool c_ok = cc >= 1 && cc <= n + 1;
  const T* dya = p.dya + c.tileoff + t.pix;
  auto met_y = [&](int row) -> T { return dya[(row + 2) * nd]; };
#pragma unroll
  for (int g = 0; g < G; ++g) {
    auto& a = s.tr[g];
    const T* sf1 = FV3T_SROW(s, g, 4);
    const T* sft = FV3T_SROW(s, g, 5);
    // q(i, r) comes back from its shared-memory row (intact until phase 1 of the next row step) and q(i, o) from the inner
    // y stream's window (after the push of cell c its qm2 is q(c-1) = q(o)): carrying them in registers made the compiler
    // copy the freshly prefetched q at the loop back-edge, a move that waited for the load at once (12 % of all stalls)
    const T qx = FV3T_SROW(s, g, 0)[0];
    const T q_o = a.yin.qm2;
    const T qj = (qx * s.in_area_r + sf1[0] - sf1[1]) * s.in_rrx;
    T dummy;
    const T fyo_c = a.you.push(cc, qj, s.cyv, c.npx, p.lim_fac, met_y, dummy);
    const T fys_c = c_ok ? (fyo_c + a.fy2_c) * s.in_mfy : T(0);
    const T qnew = q_o * s.in_ab.a + (sft[0] - sft[1] + a.fys_prev - fys_c) * s.in_ab.b;
    if (o_ok && t.cell && c.live[g]) a.qo[(o + 2) * nd] = qnew;
    a.fys_prev = fys_c;
  }
  load_fence();
  adv3_fetch4<T, OI, OO, G>(p, c, s, t, r + 1);
}

#ifdef __CUDACC__
template <class T, int OI, int OO, int G, int MINB>
__global__ void __launch_bounds__(256, MINB) k_advect3(const __grid_constant__ Adv3Params<T> p) {
  __shared__ __align__(16) T smem3[6 * G * SMP];
  Adv3Cta<T, G> c;
  if (!adv3_make_cta<T, G>(p, blockIdx.x, blockIdx.y, blockIdx.z, c)) return;
  const Adv3Thr t = adv3_thread<T, G>(c, threadIdx.x);
  Adv3State<T, OI, OO, G> s;
  adv3_init<T, OI, OO, G>(p, c, t, smem3, s);
  adv3_fetch1<T, OI, OO, G>(p, c, s, t, -2);
  adv3_fetch3<T, OI, OO, G>(p, c, s, t, -2);
  adv3_fetch4<T, OI, OO, G>(p, c, s, t, -2);
  for (int r = -2; r <= c.n + 3; ++r) {
    adv3_phase1<T, OI, OO, G>(p, c, s, t, r);
    __syncthreads();
    adv3_phase2<T, OI, OO, G>(p, c, s, t, r);
    __syncthreads();
    adv3_phase3<T, OI, OO, G>(p, c, s, t, r);
    __syncthreads();
    adv3_phase4<T, OI, OO, G>(p, c, s, t, r);
  }
}
#endif

}  // namespace fv3t
