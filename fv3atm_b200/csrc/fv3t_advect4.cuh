// fv3atm_b200: the marching advection CTA of fv3t_advect3.cuh with its inputs staged through an asynchronous-copy ring.
//
// Same arithmetic, same phases, same k_prep3 scratch.  What changes is how the eleven per-row inputs of a thread (q in its
// two corner views, {cy,yfx}, area, 1/ra_y, {cx,xfx}, cx of row o, mfx, 1/ra_x, mfy, {dp1/dp2, rarea/2dp2}) reach it:
// k_advect3 prefetches them into registers one row step ahead, but the hardware tracks outstanding loads with SIX counting
// scoreboard slots per warp that ptxas has to share between the three load groups of a row step, so the first use of any
// prefetched value also waited for the loads issued just before it -- 12-15 % of all stall samples sat on such an
// instruction whichever one was consumed early (profiles/r01_advect3_c384_ncu.txt, r01_advect3_block_sweep.txt).
// Here every thread issues cp.async (LDGSTS) copies of ALL inputs of row r+1 into its private slots of a two-deep
// shared-memory ring at the top of row step r and waits with cp.async.wait_group at the top of row step r+1: completion is
// tracked per copy group, not per register scoreboard, every input gets a whole row step of latency, and ~30 registers
// of prefetch state disappear.  The q row of the x sweeps is copied straight into its (double-buffered) exchange row.
#pragma once
#include <cstring>

#include "fv3t_advect3.cuh"

namespace fv3t {

template <int BYTES> FV3T_HD void async_copy(void* smem_dst, const void* gsrc) {
#ifdef __CUDA_ARCH__
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(sa), "l"(gsrc), "n"(BYTES) : "memory");
#else
  std::memcpy(smem_dst, gsrc, BYTES);
#endif
}
FV3T_HD void async_commit() {
#ifdef __CUDA_ARCH__
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
// wait until at most N of the most recently committed copy groups of this thread are still in flight
template <int N> FV3T_HD void async_wait_pending() {
#ifdef __CUDA_ARCH__
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}
FV3T_HD void async_wait_all() {
#ifdef __CUDA_ARCH__
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}

// shared-memory layout in elements of T (NTC = compile-time block size, PITCH = NTC + 8):
//   exchange rows (PITCH each): 0,1 q of row r (double-buffered by row parity)  2 dm/al of row r  3 q_i of row o
//                               4 dm/al of row o  5 xfx*fx2 of row r  6 (fx+fx2)*mfx of row o
//   input ring, slot d = row parity: 8 scalar arrays [NTC] then 3 pair arrays [NTC]  (14*NTC elements per slot)
template <int NTC> struct Adv4Layout {
  static constexpr int PITCH = NTC + 8;
  static constexpr int PAD = 4;
  static constexpr int XROWS = 7;
  static constexpr int NSC = 8;   // qy, area_o, rry, cxo, mfx, area_r, rrx, mfy
  static constexpr int NPR = 3;   // y2, x2r, ab
  static constexpr int SLOT = (NSC + 2 * NPR) * NTC;
  static constexpr int RING0 = XROWS * PITCH;
  static constexpr int TOTAL = RING0 + 2 * SLOT;
};
enum : int { I_QY = 0, I_AREA_O = 1, I_RRY = 2, I_CXO = 3, I_MFX = 4, I_AREA_R = 5, I_RRX = 6, I_MFY = 7 };
enum : int { P_Y2 = 0, P_X2R = 1, P_AB = 2 };

template <class T, int OI, int OO, int NTC> struct Adv4State {
  YStream<T, OI> yin;
  YStream<T, OO> you;
  T fx2_a, fx2_b, fx2_c;  // inner x flux of this thread's face at rows r-1, r-2, r-3
  T Fy_prev, fys_prev;    // yfx*fy2 / (fy+fy2)*mfy at the previous y-face
  T fy2_c, cyv;           // inner y flux / Courant number of the current row step (cross its barriers)
  const T* qg;            // this thread's (clamped) column of q, row -2
  T* qo;
  T* smt;                 // shared memory: this thread's slot of exchange row 0
  T* rsc;                 // this thread's slot of scalar array 0 of ring slot 0
  Pair<T>* rpr;           // this thread's slot of pair array 0 of ring slot 0
};

#define FV3T_XROW(s, k) ((s).smt + (k) * Adv4Layout<NTC>::PITCH)
#define FV3T_INS(s, d, f) ((s).rsc[((d) * Adv4Layout<NTC>::SLOT) + (f) * NTC])
#define FV3T_INP(s, d, f) ((s).rpr[((d) * (Adv4Layout<NTC>::SLOT / 2)) + (f) * NTC])

template <class T, int OI, int OO, int NTC>
FV3T_HD void adv4_init(const Adv3Params<T>& p, const Adv3Cta<T, 1>& c, const Adv3Thr& t, T* smem, Adv4State<T, OI, OO, NTC>& s) {
  using L = Adv4Layout<NTC>;
  s.yin.init();
  s.you.init();
  s.fx2_a = s.fx2_b = s.fx2_c = T(0);
  s.Fy_prev = s.fys_prev = T(0);
  s.fy2_c = s.cyv = T(0);
  s.qg = p.qin + c.qoff[0] + t.pix;
  s.qo = p.qout + c.qoff[0] + t.pix;
  keep_ptr(s.qg);
  keep_ptr(s.qo);
  s.smt = smem + L::PAD + t.tid;
  s.rsc = smem + L::RING0 + t.tid;
  s.rpr = reinterpret_cast<Pair<T>*>(smem + L::RING0 + L::NSC * NTC) + t.tid;
}

// issue the asynchronous copies of every input of row step r into ring slot d (and q into exchange row d)
template <class T, int OI, int OO, int NTC>
FV3T_HD void adv4_issue(const Adv3Params<T>& p, const Adv3Cta<T, 1>& c, Adv4State<T, OI, OO, NTC>& s, const Adv3Thr& t, int r, int d) {
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
  constexpr int W1 = sizeof(T), W2 = 2 * sizeof(T);
  const int L_r = c.levoff + (r + 2) * nd + t.pix, L_c = c.levoff + (cc + 2) * nd + t.pix, L_o = c.levoff + (o + 2) * nd + t.pix;
  const int A_r = c.tileoff + (r + 2) * nd + t.pix, A_o = c.tileoff + (o + 2) * nd + t.pix;
  async_copy<W1>(FV3T_XROW(s, d), s.qg + ox);
  async_copy<W1>(&FV3T_INS(s, d, I_QY), s.qg + oy);
  async_copy<W2>(&FV3T_INP(s, d, P_Y2), p.Y2 + L_c);
  async_copy<W1>(&FV3T_INS(s, d, I_AREA_O), p.area + A_o);
  async_copy<W1>(&FV3T_INS(s, d, I_RRY), p.rry + L_o);
  async_copy<W2>(&FV3T_INP(s, d, P_X2R), p.X2 + L_r);
  async_copy<W1>(&FV3T_INS(s, d, I_CXO), &p.X2[L_o].a);
  async_copy<W1>(&FV3T_INS(s, d, I_MFX), p.mfx + c.mxoff + (o - 1) * (n + 1) + t.imx);
  async_copy<W1>(&FV3T_INS(s, d, I_AREA_R), p.area + A_r);
  async_copy<W1>(&FV3T_INS(s, d, I_RRX), p.rrx + L_r);
  async_copy<W1>(&FV3T_INS(s, d, I_MFY), p.mfy + c.myoff + (cc - 1) * n + t.imy);
  async_copy<W2>(&FV3T_INP(s, d, P_AB), p.cab + L_o);
  async_commit();
}

// phase 1: inner y sweep (flux at y-face c = r-2), q_i of row o = r-3 to shared memory
template <class T, int OI, int OO, int NTC>
FV3T_HD void adv4_phase1(const Adv3Params<T>& p, const Adv3Cta<T, 1>& c, Adv4State<T, OI, OO, NTC>& s, const Adv3Thr& t, int r, int d) {
  const int n = c.n, nd = c.nd;
  const int cc = r - 2;
  const bool c_ok = cc >= 1 && cc <= n + 1;
  const Pair<T> y2 = FV3T_INP(s, d, P_Y2);
  const T cyv = c_ok ? y2.a : T(0), yfv = c_ok ? y2.b : T(0);
  const T* dya = p.dya + c.tileoff + t.pix;
  auto met_y = [&](int row) -> T { return dya[(row + 2) * nd]; };
  const T qy = FV3T_INS(s, d, I_QY);
  T q_o;
  const T fy2_c = s.yin.push(cc, qy, cyv, c.npx, p.lim_fac, met_y, q_o);
  const T Fy_c = yfv * fy2_c;
  const T qi = (q_o * FV3T_INS(s, d, I_AREA_O) + s.Fy_prev - Fy_c) * FV3T_INS(s, d, I_RRY);  // only rows o = 1..n are consumed
  s.Fy_prev = Fy_c;
  s.fy2_c = fy2_c;
  s.cyv = cyv;
  FV3T_XROW(s, 3)[0] = qi;
}

// phase 2: dm (ORD >= 7) or al (ORD < 7) of row r (inner x sweep) and of row o (outer x sweep on q_i)
template <class T, int OI, int OO, int NTC>
FV3T_HD void adv4_phase2(const Adv3Params<T>& p, const Adv3Cta<T, 1>& c, Adv4State<T, OI, OO, NTC>& s, const Adv3Thr& t, int r, int d) {
  const int nd = c.nd;
  const int o = clampi(r - 3, 1, c.n);
  const int i = t.i;
  const T* dxa = p.dxa + c.tileoff + 2;
  auto dxa_r = [&](int gi) -> T { return dxa[(r + 2) * nd + gi]; };
  auto dxa_o = [&](int gi) -> T { return dxa[(o + 2) * nd + gi]; };
  const T* sqa = FV3T_XROW(s, d) - i;  // indexable by the global column
  const T* sqb = FV3T_XROW(s, 3) - i;
  auto qa = [&](int gi) -> T { return sqa[gi]; };
  auto qb = [&](int gi) -> T { return sqb[gi]; };
  FV3T_XROW(s, 2)[0] = ppm_pre<T, OI>(i, c.npx, qa, dxa_r);
  FV3T_XROW(s, 4)[0] = ppm_pre<T, OO>(i, c.npx, qb, dxa_o);
}

// phase 3: x-face fluxes: inner sweep of row r (-> xfx*fx2), outer sweep of row o (-> (fx+fx2)*mfx)
template <class T, int OI, int OO, int NTC>
FV3T_HD void adv4_phase3(const Adv3Params<T>& p, const Adv3Cta<T, 1>& c, Adv4State<T, OI, OO, NTC>& s, const Adv3Thr& t, int r, int d) {
  const int nd = c.nd;
  const int o = clampi(r - 3, 1, c.n);
  const int i = t.i;
  const T* dxa = p.dxa + c.tileoff + 2;
  auto dxa_r = [&](int gi) -> T { return dxa[(r + 2) * nd + gi]; };
  auto dxa_o = [&](int gi) -> T { return dxa[(o + 2) * nd + gi]; };
  const T *sqa = FV3T_XROW(s, d) - i, *sda = FV3T_XROW(s, 2) - i, *sqb = FV3T_XROW(s, 3) - i, *sdb = FV3T_XROW(s, 4) - i;
  auto qa = [&](int gi) -> T { return sqa[gi]; };
  auto aa = [&](int gi) -> T { return sda[gi]; };
  auto qb = [&](int gi) -> T { return sqb[gi]; };
  auto ab = [&](int gi) -> T { return sdb[gi]; };
  const Pair<T> x2r = FV3T_INP(s, d, P_X2R);
  const T fx2 = xface_flux<T, OI>(i, x2r.a, c.npx, p.lim_fac, qa, aa, dxa_r);
  FV3T_XROW(s, 5)[0] = x2r.b * fx2;
  const T fxo = xface_flux<T, OO>(i, FV3T_INS(s, d, I_CXO), c.npx, p.lim_fac, qb, ab, dxa_o);
  FV3T_XROW(s, 6)[0] = (fxo + s.fx2_c) * FV3T_INS(s, d, I_MFX);
  s.fx2_c = s.fx2_b;
  s.fx2_b = s.fx2_a;
  s.fx2_a = fx2;
}

// phase 4: q_j of row r, outer y sweep (flux at y-face c), flux-form update of row o
template <class T, int OI, int OO, int NTC>
FV3T_HD void adv4_phase4(const Adv3Params<T>& p, const Adv3Cta<T, 1>& c, Adv4State<T, OI, OO, NTC>& s, const Adv3Thr& t, int r, int d) {
  const int n = c.n, nd = c.nd;
  const int cc = r - 2, o = r - 3;
  const bool o_ok = o >= 1 && o <= n;
  const bool c_ok = cc >= 1 && cc <= n + 1;
  const T* dya = p.dya + c.tileoff + t.pix;
  auto met_y = [&](int row) -> T { return dya[(row + 2) * nd]; };
  const T* sf1 = FV3T_XROW(s, 5);
  const T* sft = FV3T_XROW(s, 6);
  const T qx = FV3T_XROW(s, d)[0];
  const T q_o = s.yin.qm2;  // after the push of cell c the inner stream's qm2 is q(c-1) = q(o)
  const T qj = (qx * FV3T_INS(s, d, I_AREA_R) + sf1[0] - sf1[1]) * FV3T_INS(s, d, I_RRX);
  T dummy;
  const T fyo_c = s.you.push(cc, qj, s.cyv, c.npx, p.lim_fac, met_y, dummy);
  const T fys_c = c_ok ? (fyo_c + s.fy2_c) * FV3T_INS(s, d, I_MFY) : T(0);
  const Pair<T> ab = FV3T_INP(s, d, P_AB);
  const T qnew = q_o * ab.a + (sft[0] - sft[1] + s.fys_prev - fys_c) * ab.b;
  if (o_ok && t.cell) s.qo[(o + 2) * nd] = qnew;
  s.fys_prev = fys_c;
}

#ifdef __CUDACC__
template <class T, int OI, int OO, int NTC, int MINB>
__global__ void __launch_bounds__(NTC, MINB) k_advect4(const __grid_constant__ Adv3Params<T> p) {
  extern __shared__ __align__(16) unsigned char smem4_raw[];
  T* smem4 = reinterpret_cast<T*>(smem4_raw);
  Adv3Cta<T, 1> c;
  if (!adv3_make_cta<T, 1>(p, blockIdx.x, blockIdx.y, blockIdx.z, c)) return;
  const Adv3Thr t = adv3_thread<T, 1>(c, threadIdx.x);
  Adv4State<T, OI, OO, NTC> s;
  adv4_init<T, OI, OO, NTC>(p, c, t, smem4, s);
  adv4_issue<T, OI, OO, NTC>(p, c, s, t, -2, 0);
  for (int r = -2; r <= c.n + 3; ++r) {
    const int d = (r + 2) & 1;
    async_wait_all();                                       // the inputs of row r (issued one row step ago) have landed
    adv4_issue<T, OI, OO, NTC>(p, c, s, t, r + 1, d ^ 1);    // request everything row r+1 needs
    adv4_phase1<T, OI, OO, NTC>(p, c, s, t, r, d);
    __syncthreads();
    adv4_phase2<T, OI, OO, NTC>(p, c, s, t, r, d);
    __syncthreads();
    adv4_phase3<T, OI, OO, NTC>(p, c, s, t, r, d);
    __syncthreads();
    adv4_phase4<T, OI, OO, NTC>(p, c, s, t, r, d);
  }
}
#endif

}  // namespace fv3t
