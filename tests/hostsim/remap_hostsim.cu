// TEST INFRASTRUCTURE ONLY: runs the product's per-column device routine (fv3t::remap_column, a __host__ __device__
// template in fv3atm_b200/csrc/fv3t_remap2.cuh) on the CPU, column by column, so that its re-scheduled arithmetic can
// be compared bit-for-bit with the oracle where no GPU exists.  Not a fallback: nothing in fv3atm_b200/ links this.
#include "../../fv3atm_b200/csrc/fv3t_remap2.cuh"

template <class T, int G>
static void run(int ntiles, int n, int km, int nq, const T* pe, const T* ak, const T* bk, T ptop, const T* qsrc, T* qdst, T* delp,
                const int* kord, int fill) {
  fv3t::Remap2Params<T> p;
  p.qsrc = qsrc;
  p.qdst = qdst;
  p.pe = pe;
  p.ak = ak;
  p.bk = bk;
  p.delp = delp;
  p.kord = kord;
  p.ptop = ptop;
  p.n = n;
  p.km = km;
  p.nq = nq;
  p.ntiles = ntiles;
  p.fill = fill;
  p.j_first = 0;
  p.j_count = n;
  const bool mapn = nq > 5;
  for (int t = 0; t < ntiles; ++t)
    for (int j = 1; j <= n; ++j)
      for (int i = 1; i <= n; ++i)
        for (int iq0 = 0; iq0 < nq; iq0 += G) {
          if (mapn)
            fv3t::remap_column<T, G, true, 128>(p, t, i, j, iq0);
          else
            fv3t::remap_column<T, G, false, 128>(p, t, i, j, iq0);
        }
}

#define API(T, S)                                                                                                             \
  extern "C" void hostsim_remap_##S(int G, int ntiles, int n, int km, int nq, const T* pe, const T* ak, const T* bk, T ptop, \
                                    const T* qsrc, T* qdst, T* delp, const int* kord, int fill) {                             \
    if (G == 1)                                                                                                               \
      run<T, 1>(ntiles, n, km, nq, pe, ak, bk, ptop, qsrc, qdst, delp, kord, fill);                                           \
    else if (G == 2)                                                                                                          \
      run<T, 2>(ntiles, n, km, nq, pe, ak, bk, ptop, qsrc, qdst, delp, kord, fill);                                           \
    else                                                                                                                      \
      run<T, 3>(ntiles, n, km, nq, pe, ak, bk, ptop, qsrc, qdst, delp, kord, fill);                                           \
  }
API(double, f64)
API(float, f32)
