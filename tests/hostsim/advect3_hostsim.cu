// TEST INFRASTRUCTURE ONLY: runs the product's fast advection path (fv3atm_b200/csrc/fv3t_advect3.cuh: prep3_cell,
// cab3_cell and the four phase functions of the marching CTA) on the CPU, one simulated thread after the other between
// barriers, so that the kernel logic can be compared with the oracle where no GPU exists.  Not a fallback: nothing in
// fv3atm_b200/ links this.  Host orchestration mirrors Impl<T>::tracer_2d_resident of fv3t_api.cu (fast mode).
#include <cstdint>
#include <cstring>
#include <cstdint>
#include <vector>

#include "../../fv3atm_b200/csrc/fv3t_advect4.cuh"

using namespace fv3t;

static int g_group = 1;  // tracers per simulated thread (set by the test through hostsim_set_group)

template <class T, int OI, int OO, int G>
static void run_substep_g(const Adv3Params<T>& p, int NT) {
  const int n = p.n;
  const int strips = (n + p.W - 1) / p.W;
  std::vector<T> smem(6 * G * (size_t)SMP);
  std::vector<Adv3State<T, OI, OO, G>> st(NT);
  std::vector<Adv3Thr> th(NT);
  for (int lev = 0; lev < p.ntiles * p.npz; ++lev)
    for (int strip = 0; strip < strips; ++strip)
      for (int iqg = 0; iqg < (p.nq + G - 1) / G; ++iqg) {
        Adv3Cta<T, G> c;
        if (!adv3_make_cta<T, G>(p, iqg, strip, lev, c)) continue;
        for (int tid = 0; tid < NT; ++tid) {
          th[tid] = adv3_thread<T, G>(c, tid);
          adv3_init<T, OI, OO, G>(p, c, th[tid], smem.data(), st[tid]);
          adv3_fetch1<T, OI, OO, G>(p, c, st[tid], th[tid], -2);
          adv3_fetch3<T, OI, OO, G>(p, c, st[tid], th[tid], -2);
          adv3_fetch4<T, OI, OO, G>(p, c, st[tid], th[tid], -2);
        }
        for (int r = -2; r <= n + 3; ++r) {
          for (int tid = 0; tid < NT; ++tid) adv3_phase1<T, OI, OO, G>(p, c, st[tid], th[tid], r);
          for (int tid = 0; tid < NT; ++tid) adv3_phase2<T, OI, OO, G>(p, c, st[tid], th[tid], r);
          for (int tid = 0; tid < NT; ++tid) adv3_phase3<T, OI, OO, G>(p, c, st[tid], th[tid], r);
          for (int tid = 0; tid < NT; ++tid) adv3_phase4<T, OI, OO, G>(p, c, st[tid], th[tid], r);
        }
      }
}

// the asynchronous-copy ring variant (k_advect4): copies complete immediately on the host
template <class T, int OI, int OO>
static void run_substep_ring(const Adv3Params<T>& p, int NT) {
  constexpr int NTC = 256;
  const int n = p.n;
  const int strips = (n + p.W - 1) / p.W;
  std::vector<T> smem(Adv4Layout<NTC>::TOTAL + 2);
  T* sm = smem.data() + (((uintptr_t)smem.data() & 15) ? 1 : 0);  // 16-byte aligned like the device's shared memory
  std::vector<Adv4State<T, OI, OO, NTC>> st(NT);
  std::vector<Adv3Thr> th(NT);
  for (int lev = 0; lev < p.ntiles * p.npz; ++lev)
    for (int strip = 0; strip < strips; ++strip)
      for (int iq = 0; iq < p.nq; ++iq) {
        Adv3Cta<T, 1> c;
        if (!adv3_make_cta<T, 1>(p, iq, strip, lev, c)) continue;
        for (int tid = 0; tid < NT; ++tid) {
          th[tid] = adv3_thread<T, 1>(c, tid);
          adv4_init<T, OI, OO, NTC>(p, c, th[tid], sm, st[tid]);
          adv4_issue<T, OI, OO, NTC>(p, c, st[tid], th[tid], -2, 0);
        }
        for (int r = -2; r <= n + 3; ++r) {
          const int d = (r + 2) & 1;
          for (int tid = 0; tid < NT; ++tid) {
            adv4_issue<T, OI, OO, NTC>(p, c, st[tid], th[tid], r + 1, d ^ 1);
            adv4_phase1<T, OI, OO, NTC>(p, c, st[tid], th[tid], r, d);
          }
          for (int tid = 0; tid < NT; ++tid) adv4_phase2<T, OI, OO, NTC>(p, c, st[tid], th[tid], r, d);
          for (int tid = 0; tid < NT; ++tid) adv4_phase3<T, OI, OO, NTC>(p, c, st[tid], th[tid], r, d);
          for (int tid = 0; tid < NT; ++tid) adv4_phase4<T, OI, OO, NTC>(p, c, st[tid], th[tid], r, d);
        }
      }
}

template <class T, int OI, int OO>
static void run_substep(const Adv3Params<T>& p, int NT) {
  if (g_group == 4)
    run_substep_ring<T, OI, OO>(p, NT);
  else if (g_group == 2)
    run_substep_g<T, OI, OO, 2>(p, NT);
  else if (g_group == 3)
    run_substep_g<T, OI, OO, 3>(p, NT);
  else
    run_substep_g<T, OI, OO, 1>(p, NT);
}

extern "C" void hostsim_set_group(int g) { g_group = g; }

template <class T> static int dispatch(const Adv3Params<T>& p, int hord, int NT) {
  switch (hord) {
    case 8: run_substep<T, 8, 8>(p, NT); break;
    case 10: run_substep<T, 8, 10>(p, NT); break;
    case 9: run_substep<T, 9, 9>(p, NT); break;
    case 7: run_substep<T, 7, 7>(p, NT); break;
    case 11: run_substep<T, 11, 11>(p, NT); break;
    case 12: run_substep<T, 12, 12>(p, NT); break;
    case 13: run_substep<T, 13, 13>(p, NT); break;
    case 5: run_substep<T, 5, 5>(p, NT); break;
    case -5: run_substep<T, -5, -5>(p, NT); break;
    case 6: run_substep<T, 6, 6>(p, NT); break;
    case 1: run_substep<T, 1, 1>(p, NT); break;
    case 2: run_substep<T, 2, 2>(p, NT); break;
    case 3: run_substep<T, 3, 3>(p, NT); break;
    case 4: run_substep<T, 4, 4>(p, NT); break;
    default: return 1;
  }
  return 0;
}

// q [6, nq, npz, nd, nd] in/out; dp1 [6, npz, nd, nd] in/out; cx, cy, mfx, mfy in/out (scaled when nsplt != 1)
template <class T>
static int tracer_2d_sim(int n, int npz, int nq, T* q, T* dp1, T* mfx, T* mfy, T* cx, T* cy, const T* area, const T* rarea,
                         const T* dx, const T* dy, const T* dxa, const T* dya, const T* sin_sg, const int64_t* halo_dst,
                         const int64_t* halo_src, int64_t halo_len, int hord, T lim_fac, int nsplt, const int* ksplt, int NT) {
  const int nt = 6, nd = n + 6;
  const long plane = (long)nd * nd;
  const size_t nlev = (size_t)nt * npz;
  std::vector<Pair<T>> X2(nlev * plane), Y2(nlev * plane), cab(nlev * plane);
  std::vector<T> rrx(nlev * plane), rry(nlev * plane);
  std::vector<T> qb((size_t)nt * nq * npz * plane);
  Prep3Params<T> pp{cx, cy, mfx, mfy, dp1, GridDev<T>{area, rarea, dx, dy, dxa, dya, sin_sg}, X2.data(), Y2.data(), cab.data(),
                    rrx.data(), rry.data(), ksplt, n, npz, nt};
  for (int t = 0; t < nt; ++t)
    for (int kz = 0; kz < npz; ++kz)
      for (int e = 0; e < (int)plane; ++e) prep3_cell<T>(pp, t, kz, e);
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
  const long tile_stride = plane * npz * nq;
  for (int it = 1; it <= nsplt; ++it) {
    if (it > 1) {
      Cab3Params<T> cp{dp1, mfx, mfy, rarea, cab.data(), ksplt, n, npz, it};
      for (int t = 0; t < nt; ++t)
        for (int kz = 0; kz < npz; ++kz)
          for (int e = 0; e < n * n; ++e) cab3_cell<T>(cp, t, kz, e);
    }
    // edge-halo fill of every plane (complete_group_halo_update, fv_tracer2d.F90:499)
    for (long pl = 0; pl < (long)nq * npz; ++pl)
      for (int64_t e = 0; e < halo_len; ++e) {
        const int64_t d = halo_dst[e], s = halo_src[e];
        q[(d / plane) * tile_stride + pl * plane + d % plane] = q[(s / plane) * tile_stride + pl * plane + s % plane];
      }
    Adv3Params<T> p;
    p.qin = q;
    p.qout = qb.data();
    p.X2 = X2.data();
    p.Y2 = Y2.data();
    p.rrx = rrx.data();
    p.rry = rry.data();
    p.cab = cab.data();
    p.mfx = mfx;
    p.mfy = mfy;
    p.area = area;
    p.dxa = dxa;
    p.dya = dya;
    p.ksplt = ksplt;
    p.n = n;
    p.npz = npz;
    p.nq = nq;
    p.ntiles = nt;
    p.it = it;
    p.W = NT - 6;
    p.lim_fac = lim_fac;
    if (dispatch<T>(p, hord, NT)) return 1;
    // copy the compute domain of the active levels back (the product ping-pongs two buffers instead)
    for (int t = 0; t < nt; ++t)
      for (int iq = 0; iq < nq; ++iq)
        for (int kz = 0; kz < npz; ++kz) {
          if (it > ksplt[kz]) continue;
          const long o = (((long)t * nq + iq) * npz + kz) * plane;
          for (int j = 1; j <= n; ++j)
            std::memcpy(q + o + (long)(j + 2) * nd + 3, qb.data() + o + (long)(j + 2) * nd + 3, sizeof(T) * n);
        }
  }
  // the dp1 post-state: the reference advances dp1 after every sub-step but the last (fv_tracer2d.F90:547-553); the
  // product does it lazily at the start of the next sub-step (cab3_cell), so nothing is left to do here
  return 0;
}

#define API(T, S)                                                                                                              \
  extern "C" int hostsim_tracer_2d_##S(int n, int npz, int nq, T* q, T* dp1, T* mfx, T* mfy, T* cx, T* cy, const T* area,      \
                                       const T* rarea, const T* dx, const T* dy, const T* dxa, const T* dya, const T* sin_sg, \
                                       const int64_t* halo_dst, const int64_t* halo_src, int64_t halo_len, int hord,          \
                                       T lim_fac, int nsplt, const int* ksplt, int NT) {                                       \
    return tracer_2d_sim<T>(n, npz, nq, q, dp1, mfx, mfy, cx, cy, area, rarea, dx, dy, dxa, dya, sin_sg, halo_dst, halo_src,  \
                            halo_len, hord, lim_fac, nsplt, ksplt, NT);                                                        \
  }
API(double, f64)
API(float, f32)
