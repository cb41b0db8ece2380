// TEST INFRASTRUCTURE ONLY: runs the product's multi-tracer advection path (fv3atm_b200/csrc/fv3t_advect5.cuh: prep5_cell and the
// four phase functions of a marching tracer group) on the CPU, one simulated thread after the other between barriers, with the
// TMA boxes of the producer warp restated as plain copies (zero fill outside the tensor), so that the kernel logic can be
// compared with the oracle where no GPU exists.  Not a fallback: nothing in fv3atm_b200/ links this.  Host orchestration
// mirrors Impl<T>::substep / prepare of fv3t_api.cu (k_advect5 path).
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../fv3atm_b200/csrc/fv3t_advect5.cuh"

using namespace fv3t;

// what the ten cp.async.bulk.tensor boxes of block b deliver (k_advect5, producer warp)
template <class T> static void fill_stage(const Adv5Params<T>& p, int strip, int levc, int b, unsigned char* stage) {
  const int n = p.n, nd = n + 6, PP = a5_pitch(n);
  const int lev = p.lev0 + levc, tile = lev / p.npz;
  const int xs = strip * A5_W;
  int rows[A5_NPAIR + A5_NSC];
  a5_box_rows<T>(b, rows);
  const Pair<T>* psrc[A5_NPAIR] = {p.X2, p.X2, p.Y2, p.CAB};
  Pair<T>* pd = reinterpret_cast<Pair<T>*>(stage);
  for (int f = 0; f < A5_NPAIR; ++f)
    for (int k = 0; k < A5_R; ++k)
      for (int x = 0; x < A5_GW; ++x) {
        const int row = rows[f] + k, col = xs + x;
        Pair<T> v{T(0), T(0)};
        if (row >= 0 && row < nd && col >= 0 && col < PP) v = psrc[f][((long)levc * nd + row) * PP + col];
        pd[(f * A5_R + k) * A5_GW + x] = v;
      }
  const T* ssrc[A5_NSC] = {p.RX, p.MFX, p.RY, p.MFY, p.AREA, p.AREA, p.RAREA};
  T* sd = reinterpret_cast<T*>(stage + A5Stage<T>::PAIR_BYTES);
  const int xq = xs - A5Stage<T>::shift(xs);
  for (int f = 0; f < A5_NSC; ++f)
    for (int k = 0; k < A5_R; ++k)
      for (int x = 0; x < A5Stage<T>::SW; ++x) {
        const int row = rows[A5_NPAIR + f] + k, col = xq + x;
        const long plane_idx = (f >= A5_AR) ? tile : levc;
        T v = T(0);
        if (row >= 0 && row < nd && col >= 0 && col < PP) v = ssrc[f][(plane_idx * nd + row) * PP + col];
        sd[f * A5Stage<T>::SFIELD + k * A5Stage<T>::SW + x] = v;
      }
}

// one block of four row steps in the product's two-interval schedule (adv5_block): interval 1 = phase 2 of step s + phase 4 of
// step s-1, interval 2 = phase 3 of step s + phase 1 of step s+1; a barrier (= end of a loop over the threads) after each
template <class T, int OI, int OO, bool EX, class P> struct Sim {
  const P& p;
  const Adv5Cta& c;
  std::vector<Adv5State<T, OI, OO>>& st;
  const std::vector<Adv3Thr>& th;

  template <int PH, int PHQ, bool YE, bool XE, int PH4, int PH1>
  void step(const unsigned char* stage4, const unsigned char* stage, const unsigned char* stage1, int r, bool do4, bool do1) {
    for (int tid = 0; tid < A5_GW; ++tid) {
      adv5_issue_q<T, OI, OO, PHQ, YE>(p, c, st[tid], th[tid], r + 2);
      adv5_phase2<T, OI, OO, PH, XE>(p, c, st[tid], th[tid], r);
      if (do4) adv5_phase4<T, OI, OO, PH4, YE, EX>(p, c, st[tid], th[tid], a5_view<T>(stage4, tid, c.i0 - 1), r - 1);
    }
    for (int tid = 0; tid < A5_GW; ++tid) {
      adv5_phase3<T, OI, OO, PH, XE, EX>(p, c, st[tid], th[tid], a5_view<T>(stage, tid, c.i0 - 1), r);
      if (do1) adv5_phase1<T, OI, OO, PH1, YE, EX>(p, c, st[tid], th[tid], a5_view<T>(stage1, tid, c.i0 - 1), r + 1);
    }
  }
  template <bool YE, bool XE> void block(const unsigned char* sp, const unsigned char* sc, const unsigned char* sn, int r0) {
    step<0, 2, YE, XE, 3, 1>(sp, sc, sc, r0, !YE || r0 > -2, true);
    step<1, 3, YE, XE, 0, 2>(sc, sc, sc, r0 + 1, true, true);
    step<2, 0, YE, XE, 1, 3>(sc, sc, sc, r0 + 2, true, true);
    step<3, 1, YE, XE, 2, 0>(sc, sc, sn, r0 + 3, true, sn != nullptr);
  }
};

template <class T, int OI, int OO, bool EX, class P> static void run_substep(const P& p) {
  const int n = p.n;
  const int strips = (n + A5_W - 1) / A5_W;
  const int nblocks = a5_nblocks(n);
  std::vector<unsigned char> stage(A5_NS * A5Stage<T>::BYTES + 16);
  unsigned char* sg = stage.data() + ((16 - ((uintptr_t)stage.data() & 15)) & 15);
  std::vector<T> gsm(A5Stage<T>::GROUP_ELEMS + 2);
  T* gs = gsm.data() + (((uintptr_t)gsm.data() & 15) ? 1 : 0);
  std::vector<Adv5State<T, OI, OO>> st(A5_GW);
  std::vector<Adv3Thr> th(A5_GW);
  for (int levc = 0; levc < p.ntiles * p.npz; ++levc)
    for (int strip = 0; strip < strips; ++strip)
      for (int iq = p.iq0; iq < p.iq0 + p.nql; ++iq) {
        Adv5Cta c;
        if (!adv5_make_cta<T>(p, strip, levc, iq, c)) continue;
        Sim<T, OI, OO, EX, P> sim{p, c, st, th};
        fill_stage<T>(p, strip, levc, 0, sg);
        for (int tid = 0; tid < A5_GW; ++tid) {
          th[tid] = adv5_thread(c, tid);
          adv5_init<T, OI, OO>(p, c, th[tid], gs, st[tid]);
          adv5_issue_q<T, OI, OO, 0, true>(p, c, st[tid], th[tid], -2);
          adv5_issue_q<T, OI, OO, 1, true>(p, c, st[tid], th[tid], -1);
          adv5_phase1<T, OI, OO, 0, true, EX>(p, c, st[tid], th[tid], a5_view<T>(sg, tid, c.i0 - 1), -2);
        }
        for (int b = 0; b < nblocks; ++b) {
          const int r0 = -2 + A5_R * b;
          unsigned char* sc = sg + (b % A5_NS) * A5Stage<T>::BYTES;
          unsigned char* sp = sg + ((b + A5_NS - 1) % A5_NS) * A5Stage<T>::BYTES;
          unsigned char* sn = nullptr;
          if (b + 1 < nblocks) {  // the producer is (at least) one box ahead
            sn = sg + ((b + 1) % A5_NS) * A5Stage<T>::BYTES;
            fill_stage<T>(p, strip, levc, b + 1, sn);
          }
          if (a5_block_interior(r0, n)) {
            if (c.xedge)
              sim.template block<false, true>(sp, sc, sn, r0);
            else
              sim.template block<false, false>(sp, sc, sn, r0);
          } else {
            sim.template block<true, true>(sp, sc, sn, r0);
          }
        }
      }
}

// the instantiations the product builds: fast for fv3t::fast_hord_ok, exact for every scheme.  With the sub-tile parameter type
// (the instantiations of sub-tile contexts) only the schemes the sub-mosaic tests run, to bound the build time of this harness.
template <class T, bool EX, class P> static int dispatch(const P& p, int hord) {
  switch (hord) {
    case 8: run_substep<T, 8, 8, EX>(p); return 0;
  }
  if constexpr (!P::SUB) {
    switch (hord) {
      case 11: run_substep<T, 11, 11, EX>(p); return 0;
      case 2: run_substep<T, 2, 2, EX>(p); return 0;
    }
  }
  if constexpr (EX) {
    switch (hord) {
      case 10: run_substep<T, 8, 10, true>(p); return 0;
      case 13: run_substep<T, 13, 13, true>(p); return 0;
      case 5: run_substep<T, 5, 5, true>(p); return 0;
      case -5: run_substep<T, -5, -5, true>(p); return 0;
    }
    if constexpr (!P::SUB) {
      switch (hord) {
        case 9: run_substep<T, 9, 9, true>(p); return 0;
        case 12: run_substep<T, 12, 12, true>(p); return 0;
        case 7: run_substep<T, 7, 7, true>(p); return 0;
        case 6: run_substep<T, 6, 6, true>(p); return 0;
        case 1: run_substep<T, 1, 1, true>(p); return 0;
        case 3: run_substep<T, 3, 3, true>(p); return 0;
        case 4: run_substep<T, 4, 4, true>(p); return 0;
      }
    }
  }
  return 1;
}

// q [6, nq, npz, nd, nd] in/out; dp1 [6, npz, nd, nd] in/out; cx, cy, mfx, mfy in/out (scaled when nsplt != 1)
template <class T>
static int tracer_2d_sim(int n, int npz, int nq, T* q, T* dp1, T* mfx, T* mfy, T* cx, T* cy, const T* area, const T* rarea,
                         const T* dx, const T* dy, const T* dxa, const T* dya, const T* sin_sg, const int64_t* halo_dst,
                         const int64_t* halo_src, int64_t halo_len, int hord, T lim_fac, int nsplt, const int* ksplt, int exact,
                         int nt = 6, const int* sub = nullptr) {
  const int nd = n + 6, PP = a5_pitch(n);
  const long plane = (long)nd * nd;
  const size_t nlev = (size_t)nt * npz, pe = nlev * nd * PP;
  std::vector<Pair<T>> X2(pe, Pair<T>{T(0), T(0)}), Y2(pe, Pair<T>{T(0), T(0)}), CAB(pe, Pair<T>{T(0), T(0)});
  std::vector<T> RX(pe, T(0)), RY(pe, T(0)), MX(pe, T(0)), MY(pe, T(0)), AREA((size_t)nt * nd * PP, T(0)), RAREA((size_t)nt * nd * PP, T(0));
  for (int t = 0; t < nt; ++t)
    for (int r = 0; r < nd; ++r)
      for (int x = 0; x < nd; ++x) {
        AREA[((size_t)t * nd + r) * PP + x] = area[(size_t)t * plane + (size_t)r * nd + x];
        RAREA[((size_t)t * nd + r) * PP + x] = rarea[(size_t)t * plane + (size_t)r * nd + x];
      }
  std::vector<T> qb((size_t)nt * nq * npz * plane);
  const long tile_stride = plane * npz * nq;
  for (int it = 1; it <= nsplt; ++it) {
    Prep5Params<T> pp{cx, cy, mfx, mfy, dp1, GridDev<T>{area, rarea, dx, dy, dxa, dya, sin_sg}, X2.data(), Y2.data(), CAB.data(), RX.data(),
                      RY.data(), MX.data(), MY.data(), ksplt, n, npz, nt, 0, (int)nlev, it, it == 1 ? 1 : 0, exact};
    for (int levc = 0; levc < (int)nlev; ++levc)
      for (int e = 0; e < (int)plane; ++e) prep5_cell<T>(pp, levc, e);
    // edge-halo fill of every plane (complete_group_halo_update, fv_tracer2d.F90:499)
    for (long pl = 0; pl < (long)nq * npz; ++pl)
      for (int64_t e = 0; e < halo_len; ++e) {
        const int64_t d = halo_dst[e], s = halo_src[e];
        q[(d / plane) * tile_stride + pl * plane + d % plane] = q[(s / plane) * tile_stride + pl * plane + s % plane];
      }
    // one resident "tile" (whole tile or sub-domain) at a time
    for (int t = 0; t < nt; ++t) {
      const size_t so = (size_t)t * npz * nd * PP;
      Adv5ParamsSub<T> p{};
      p.qin = q + (size_t)t * tile_stride;
      p.qout = qb.data() + (size_t)t * tile_stride;
      p.X2 = X2.data() + so;
      p.Y2 = Y2.data() + so;
      p.CAB = CAB.data() + so;
      p.RX = RX.data() + so;
      p.RY = RY.data() + so;
      p.MFX = MX.data() + so;
      p.MFY = MY.data() + so;
      p.AREA = AREA.data() + (size_t)t * nd * PP;
      p.RAREA = RAREA.data() + (size_t)t * nd * PP;
      p.dxa = dxa + (size_t)t * plane;
      p.dya = dya + (size_t)t * plane;
      p.ksplt = ksplt;
      p.n = n;
      p.npz = npz;
      p.nq = nq;
      p.ntiles = 1;
      p.it = it;
      p.lev0 = 0;
      p.tg = 1;
      p.iq0 = 0;
      p.nql = nq;
      p.lim_fac = lim_fac;
      int rc;
      if (sub) {  // the parameter type, and with it the instantiations, of a sub-tile context
        p.sub[0] = A5Sub{sub[5 * t], sub[5 * t + 1], sub[5 * t + 2], sub[5 * t + 3], sub[5 * t + 4]};
        rc = exact ? dispatch<T, true>(p, hord) : dispatch<T, false>(p, hord);
      } else {
        const Adv5Params<T>& pw = p;
        rc = exact ? dispatch<T, true>(pw, hord) : dispatch<T, false>(pw, hord);
      }
      if (rc) return 1;
    }
    for (int t = 0; t < nt; ++t)
      for (int iq = 0; iq < nq; ++iq)
        for (int kz = 0; kz < npz; ++kz) {
          if (it > ksplt[kz]) continue;
          const long o = (((long)t * nq + iq) * npz + kz) * plane;
          for (int j = 1; j <= n; ++j)
            std::memcpy(q + o + (long)(j + 2) * nd + 3, qb.data() + o + (long)(j + 2) * nd + 3, sizeof(T) * n);
        }
  }
  // the in-place 1/ksplt scaling of cx, cy, mfx, mfy (fv_tracer2d.F90:463-481) is applied last (Impl<T>::finish)
  if (nsplt != 1) {
    const long ncx = (long)(n + 1) * nd, nmf = (long)(n + 1) * n;
    for (size_t lev = 0; lev < nlev; ++lev) {
      const T frac = T(1) / (T)ksplt[lev % npz];
      for (long e = 0; e < ncx; ++e) {
        cx[lev * ncx + e] *= frac;
        cy[lev * ncx + e] *= frac;
      }
      for (long e = 0; e < nmf; ++e) {
        mfx[lev * nmf + e] *= frac;
        mfy[lev * nmf + e] *= frac;
      }
    }
  }
  return 0;
}

#define API_SUB(T, S)                                                                                                          \
  extern "C" int hostsim5_tracer_2d_sub_##S(int nt, int n, int npz, int nq, T* q, T* dp1, T* mfx, T* mfy, T* cx, T* cy,        \
                                            const T* area, const T* rarea, const T* dx, const T* dy, const T* dxa,             \
                                            const T* dya, const T* sin_sg, const int64_t* halo_dst, const int64_t* halo_src,   \
                                            int64_t halo_len, int hord, T lim_fac, int nsplt, const int* ksplt, int exact,     \
                                            const int* sub) {                                                                  \
    return tracer_2d_sim<T>(n, npz, nq, q, dp1, mfx, mfy, cx, cy, area, rarea, dx, dy, dxa, dya, sin_sg, halo_dst, halo_src,  \
                            halo_len, hord, lim_fac, nsplt, ksplt, exact, nt, sub);                                            \
  }
#define API(T, S)                                                                                                              \
  extern "C" int hostsim5_tracer_2d_##S(int n, int npz, int nq, T* q, T* dp1, T* mfx, T* mfy, T* cx, T* cy, const T* area,     \
                                        const T* rarea, const T* dx, const T* dy, const T* dxa, const T* dya, const T* sin_sg, \
                                        const int64_t* halo_dst, const int64_t* halo_src, int64_t halo_len, int hord,          \
                                        T lim_fac, int nsplt, const int* ksplt, int exact) {                                   \
    return tracer_2d_sim<T>(n, npz, nq, q, dp1, mfx, mfy, cx, cy, area, rarea, dx, dy, dxa, dya, sin_sg, halo_dst, halo_src,  \
                            halo_len, hord, lim_fac, nsplt, ksplt, exact);                                                            \
  }
API(double, f64)
API(float, f32)
API_SUB(double, f64)
API_SUB(float, f32)
