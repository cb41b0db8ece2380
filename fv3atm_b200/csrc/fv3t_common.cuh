// fv3atm_b200: shared device helpers for the tracer-transport kernels (sm_100a).
//
// Arithmetic contract: every expression in the kernels keeps the operation order of the reference
// Fortran (atmos_cubed_sphere/model/tp_core.F90, fv_tracer2d.F90, fv_mapz.F90, fv_fill.F90).  Built with
// -fmad=false and IEEE division the fp64/fp32 results are bit-identical to an FMA-free CPU evaluation.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace fv3t {

#define FV3T_HD __host__ __device__ __forceinline__

constexpr int NG = 3;  // halo width (fv_mp_mod.F90:104)

// Fortran SIGN(a,b): |a| with the sign bit of b (tp_core.F90:565)
FV3T_HD double f_sign(double a, double b) { return copysign(a, b); }
FV3T_HD float f_sign(float a, float b) { return copysignf(a, b); }
FV3T_HD double f_abs(double a) { return fabs(a); }
FV3T_HD float f_abs(float a) { return fabsf(a); }
// Fortran MAX/MIN on non-NaN data
template <class T> FV3T_HD T f_max(T a, T b) { return a > b ? a : b; }
template <class T> FV3T_HD T f_min(T a, T b) { return a < b ? a : b; }
template <class T> FV3T_HD T f_max(T a, T b, T c) { return f_max(f_max(a, b), c); }
template <class T> FV3T_HD T f_min(T a, T b, T c) { return f_min(f_min(a, b), c); }
template <class T> FV3T_HD T f_max(T a, T b, T c, T d) { return f_max(f_max(f_max(a, b), c), d); }
template <class T> FV3T_HD T f_min(T a, T b, T c, T d) { return f_min(f_min(f_min(a, b), c), d); }

// sign(min(|x|, |y|), x) (tp_core.F90:575-576) without materialising the absolute values: the magnitude comes from whichever
// operand is smaller in modulus, the sign bit is overwritten anyway.  Bit-identical to f_sign(f_min(f_abs(x), f_abs(y)), x).
template <class T> FV3T_HD T sign_min_abs(T x, T y) { return f_sign((f_abs(x) < f_abs(y)) ? x : y, x); }
// sign(min(|xt|, max(m, 0)), xt) (the monotone slope of tp_core.F90:563-567 once m = min(hi - q0, q0 - lo) is known), same idea
template <class T> FV3T_HD T dm_limit(T xt, T m) {
  const T t = (f_abs(xt) < m) ? xt : m;
  return f_sign((m > T(0)) ? t : T(0), xt);
}

// L1 prefetch of a global line that a later iteration of a column sweep will read (no register is tied up)
FV3T_HD void prefetch_l1(const void* p) {
#ifdef __CUDA_ARCH__
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

// default-real literal constants of tp_core.F90:62-98 and fv_mapz.F90:110, folded at compile time in T
template <class T> struct K {
  static FV3T_HD constexpr T r3() { return T(1) / T(3); }
  static FV3T_HD constexpr T r23() { return T(2) / T(3); }
  static FV3T_HD constexpr T r12() { return T(1) / T(12); }
  static FV3T_HD constexpr T near_zero() { return T(1.0e-25); }
  static FV3T_HD constexpr T ppm_fac() { return T(1.5); }
  static FV3T_HD constexpr T s11() { return T(11) / T(14); }
  static FV3T_HD constexpr T s14() { return T(4) / T(7); }
  static FV3T_HD constexpr T s15() { return T(3) / T(14); }
  static FV3T_HD constexpr T c1() { return T(-2) / T(14); }
  static FV3T_HD constexpr T c2() { return T(11) / T(14); }
  static FV3T_HD constexpr T c3() { return T(5) / T(14); }
  static FV3T_HD constexpr T p1() { return T(7) / T(12); }
  static FV3T_HD constexpr T p2() { return T(-1) / T(12); }
};

}  // namespace fv3t
