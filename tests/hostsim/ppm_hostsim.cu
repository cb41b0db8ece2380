// TEST INFRASTRUCTURE ONLY: evaluates the product's 1-D PPM element functions (fv3atm_b200/csrc/fv3t_ppm.cuh: ppm_pre, ppm_blbr,
// ppm_flux through xface_flux of fv3t_advect2.cuh -- the functions every advection kernel is built from) on the CPU for one
// periodic line, so that they can be pinned DIRECTLY against the golden vectors generated from the reference's own notebook
// (tests/golden/), independently of the oracle.  Built with contraction off: the comparison is bit-for-bit.
#include <vector>

#include "../../fv3atm_b200/csrc/fv3t_advect2.cuh"

using namespace fv3t;

// q_ext: cells 1-ng .. nx+ng (ng = 3, periodic halos filled by the caller); c: faces 1 .. nx+1; flux: faces 1 .. nx+1
template <class T, int ORD> static void line(int nx, const T* q_ext, const T* c, T* flux, T lim_fac) {
  constexpr int S = 1000;          // index shift: every evaluated index lies far from the tile-edge branches
  const int npx = 1000000;
  std::vector<T> a(nx + 8, T(0));  // dm (ORD >= 7) or al (ORD < 7) of cells -1 .. nx+2
  auto q = [&](int g) -> T { return q_ext[(g - S) - 1 + 3]; };
  auto dxa = [&](int) -> T { return T(1); };
  for (int i = (ORD >= 7 ? -1 : 0); i <= nx + 2; ++i) a[i + 1] = ppm_pre<T, ORD>(i + S, npx, q, dxa);
  auto av = [&](int g) -> T { return a[(g - S) + 1]; };
  for (int i = 1; i <= nx + 1; ++i) flux[i - 1] = xface_flux<T, ORD>(i + S, c[i - 1], npx, lim_fac, q, av, dxa);
}

extern "C" int hostsim_ppm_line_f64(int ord, int nx, const double* q_ext, const double* c, double* flux, double lim_fac) {
  switch (ord) {
    case 8: line<double, 8>(nx, q_ext, c, flux, lim_fac); break;
    case 10: line<double, 10>(nx, q_ext, c, flux, lim_fac); break;
    case 5: line<double, 5>(nx, q_ext, c, flux, lim_fac); break;
    case -5: line<double, -5>(nx, q_ext, c, flux, lim_fac); break;
    case 6: line<double, 6>(nx, q_ext, c, flux, lim_fac); break;
    case 9: line<double, 9>(nx, q_ext, c, flux, lim_fac); break;
    case 13: line<double, 13>(nx, q_ext, c, flux, lim_fac); break;
    case 7: line<double, 7>(nx, q_ext, c, flux, lim_fac); break;
    default: return 1;
  }
  return 0;
}
