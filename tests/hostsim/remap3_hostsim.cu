// TEST INFRASTRUCTURE ONLY: runs the product's fast remap routines (fv3t::remap_coef_column, fv3t::remap3_column of
// fv3atm_b200/csrc/fv3t_remap3.cuh) on the CPU, column by column.  Not a fallback: nothing in fv3atm_b200/ links this.
#include <vector>

#include "../../fv3atm_b200/csrc/fv3t_remap5.cuh"

using namespace fv3t;

static int g_variant = 3;  // 3: remap3_column (three walks), 5: remap5_column (two walks, bottom-up elimination)
extern "C" void hostsim_remap_variant(int v) { g_variant = v; }

template <class T, int AK>
static void run_cols(const Remap3Params<T>& p) {
  for (int t = 0; t < p.ntiles; ++t)
    for (int j = 1; j <= p.n; ++j)
      for (int i = 1; i <= p.n; ++i)
        for (int iq = 0; iq < p.nq; ++iq) {
          if (g_variant == 5) {
            Pair<T> ring[4];
            remap5_column<T, AK, 128>(p, ring, 1, t, i, j, iq);
          } else {
            Pair<T> ring[4];
            remap3_column<T, AK, true, 128>(p, ring, 1, t, i, j, iq);
          }
        }
}

template <class T>
static int run(int ntiles, int n, int km, int nq, const T* pe, const T* ak, const T* bk, T ptop, const T* qsrc, T* qdst, T* delp,
               int akord, int fill) {
  const long plane = (long)(n + 6) * (n + 6);
  std::vector<Pair<T>> P1((size_t)ntiles * plane * (km + 1));
  std::vector<T> GAM((size_t)ntiles * plane * (km + 1)), RD1((size_t)ntiles * plane * km), R2((size_t)ntiles * plane * km * 2);
  Remap3Params<T> p{qsrc, qdst, pe, ak, bk, delp, P1.data(), GAM.data(), RD1.data(), R2.data(), ptop, n, km, nq, ntiles, fill};
  for (int t = 0; t < ntiles; ++t)
    for (int j = 1; j <= n; ++j)
      for (int i = 1; i <= n; ++i) {
        if (g_variant == 5)
          remap_coef5_column<T>(p, t, i, j);
        else
          remap_coef_column<T>(p, t, i, j);
      }
  switch (akord) {
    case 8: run_cols<T, 8>(p); break;
    case 9: run_cols<T, 9>(p); break;
    case 10: run_cols<T, 10>(p); break;
    case 11: run_cols<T, 11>(p); break;
    case 12: run_cols<T, 12>(p); break;
    case 13: run_cols<T, 13>(p); break;
    case 14: run_cols<T, 14>(p); break;
    case 15: run_cols<T, 15>(p); break;
    case 16: run_cols<T, 16>(p); break;
    case 17: run_cols<T, 17>(p); break;
    default: return 1;
  }
  return 0;
}

extern "C" int hostsim_remap3_f64(int ntiles, int n, int km, int nq, const double* pe, const double* ak, const double* bk,
                                  double ptop, const double* qsrc, double* qdst, double* delp, int akord, int fill) {
  return run<double>(ntiles, n, km, nq, pe, ak, bk, ptop, qsrc, qdst, delp, akord, fill);
}
extern "C" int hostsim_remap3_f32(int ntiles, int n, int km, int nq, const float* pe, const float* ak, const float* bk, float ptop,
                                  const float* qsrc, float* qdst, float* delp, int akord, int fill) {
  return run<float>(ntiles, n, km, nq, pe, ak, bk, ptop, qsrc, qdst, delp, akord, fill);
}
