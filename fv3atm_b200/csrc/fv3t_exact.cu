// fv3atm_b200: translation unit of the exact-arithmetic instantiations of k_advect5 -- compiled with -fmad=false, IEEE division
// (see fv3t_advect5_launch.cuh).  Every hord_tr of xppm / yppm (tp_core.F90:332-1124): the schemes whose limiters are discontinuous
// functions of their inputs (1, 3-7, -5: bl*br < 0; 9, 12, 13: the positive-definite constraint that zeroes both edge
// perturbations; 10: near_zero / |3(bl+br)| > |bl-br|) flip individual cells under ANY re-association, so they run the marching
// multi-tracer kernel in the reference's own operation order and reproduce the oracle bit for bit.
#include "fv3t_fast.h"
#include "fv3t_advect5_launch.cuh"

namespace fv3t {

template <class T, class P> static cudaError_t exact_dispatch(P p, const Adv5Maps& m, int hord, int nlev, cudaStream_t stream) {
  switch (hord) {
    case 8: return launch5_ord<T, 8, 8, true>(p, m, nlev, stream);
    case 10: return launch5_ord<T, 8, 10, true>(p, m, nlev, stream);  // ord_in = 8 when hord == 10 (tp_core.F90:157-161)
#ifndef FV3T_A5_DEV
    case 9: return launch5_ord<T, 9, 9, true>(p, m, nlev, stream);
    case 7: return launch5_ord<T, 7, 7, true>(p, m, nlev, stream);
    case 11: return launch5_ord<T, 11, 11, true>(p, m, nlev, stream);
    case 12: return launch5_ord<T, 12, 12, true>(p, m, nlev, stream);
    case 13: return launch5_ord<T, 13, 13, true>(p, m, nlev, stream);
    case 5: return launch5_ord<T, 5, 5, true>(p, m, nlev, stream);
    case -5: return launch5_ord<T, -5, -5, true>(p, m, nlev, stream);
    case 6: return launch5_ord<T, 6, 6, true>(p, m, nlev, stream);
    case 1: return launch5_ord<T, 1, 1, true>(p, m, nlev, stream);
    case 2: return launch5_ord<T, 2, 2, true>(p, m, nlev, stream);
    case 3: return launch5_ord<T, 3, 3, true>(p, m, nlev, stream);
    case 4: return launch5_ord<T, 4, 4, true>(p, m, nlev, stream);
#endif
    default: return cudaErrorInvalidValue;
  }
}
template <class T> cudaError_t exact_advect5(Adv5Params<T> p, const Adv5Maps& m, int hord, int nlev, cudaStream_t stream) {
  return exact_dispatch<T>(p, m, hord, nlev, stream);
}
template <class T> cudaError_t exact_advect5_sub(Adv5ParamsSub<T> p, const Adv5Maps& m, int hord, int nlev, cudaStream_t stream) {
  return exact_dispatch<T>(p, m, hord, nlev, stream);
}

// one object per precision and per parameter type (whole tiles / sub-tile contexts), built in parallel: FV3T_INST_F64 / _F32 and
// FV3T_INST_WHOLE / _SUB select; with none of them defined everything is instantiated
#if defined(FV3T_INST_WHOLE) || !defined(FV3T_INST_SUB)
#define FV3T_EXACT_WHOLE(T) template cudaError_t exact_advect5<T>(Adv5Params<T>, const Adv5Maps&, int, int, cudaStream_t);
#else
#define FV3T_EXACT_WHOLE(T)
#endif
#if defined(FV3T_INST_SUB) || !defined(FV3T_INST_WHOLE)
#define FV3T_EXACT_SUB(T) template cudaError_t exact_advect5_sub<T>(Adv5ParamsSub<T>, const Adv5Maps&, int, int, cudaStream_t);
#else
#define FV3T_EXACT_SUB(T)
#endif
#if defined(FV3T_INST_F64) || !defined(FV3T_INST_F32)
FV3T_EXACT_WHOLE(double)
FV3T_EXACT_SUB(double)
#endif
#if defined(FV3T_INST_F32) || !defined(FV3T_INST_F64)
FV3T_EXACT_WHOLE(float)
FV3T_EXACT_SUB(float)
#endif

}  // namespace fv3t
