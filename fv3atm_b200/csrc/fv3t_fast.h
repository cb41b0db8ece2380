// fv3atm_b200: launchers of the production ("fast") kernels.  They live in their own translation unit (fv3t_fast.cu,
// built with FMA contraction on) so that the strict kernels of fv3t_api.cu keep -fmad=false and stay bit-identical
// to the FMA-free oracle.  Each launcher enqueues on `stream` and returns cudaGetLastError().
#pragma once
#include <cuda_runtime.h>

#include "fv3t_advect3.cuh"
#include "fv3t_advect4.cuh"
#include "fv3t_advect5.cuh"
#include "fv3t_remap3.cuh"
#include "fv3t_remap5.cuh"
#include "fv3t_remap4.cuh"

namespace fv3t {

// hord values whose limiter is a continuous function of its inputs: only these may use the fast path
// (9, 12, 13 carry the positive-definite constraint that zeroes BOTH edge perturbations when the parabola's minimum crosses
// zero -- a jump: measured 3.8e-3 on the sparse random tracer in fp32 at C384 with four sub-steps -- so they are not in the set)
inline bool fast_hord_ok(int hord) { return hord == 8 || hord == 11 || hord == 2; }

// abs(kord) values whose limiter decisions depend only on the INPUT cell means (not on computed interface values): only
// these may use the fast remap (kord 10, 11, 15, 16 compare computed quantities that are exactly equal on flat data)
inline bool fast_kord_ok(int akord) { return akord <= 9 || akord == 12 || akord == 13 || akord == 14 || akord >= 17; }

template <class T> cudaError_t fast_remap_coef3(const Remap3Params<T>& p, cudaStream_t stream);
// all tracers share abs(kord) = akord (8 stands for every value <= 8, 17 for every value >= 17); mapn_tracer form (nq > 5)
template <class T> cudaError_t fast_remap3(const Remap3Params<T>& p, int akord, cudaStream_t stream);
template <class T> cudaError_t fast_prep3(const Prep3Params<T>& p, cudaStream_t stream);
template <class T> cudaError_t fast_scale3(T* cx, T* cy, T* mfx, T* mfy, const int* ksplt, int n, int npz, int ntiles, cudaStream_t stream);
template <class T> cudaError_t fast_cab3(const Cab3Params<T>& p, int ntiles, cudaStream_t stream);
// NT = threads per CTA (strip width NT-6); p.W is set by the launcher
template <class T> cudaError_t fast_advect3(Adv3Params<T> p, int hord, int NT, cudaStream_t stream);

// ---- multi-tracer TMA-staged advection (fv3t_advect5.cuh) ----
// every scheme of xppm / yppm has an exact-arithmetic instantiation of k_advect5 (fv3t_exact.cu); fast_hord_ok also a fast one
inline bool adv5_hord_ok(int hord) { return (hord >= 1 && hord <= 13) || hord == -5; }
template <class T> cudaError_t fast_prep5(const Prep5Params<T>& p, cudaStream_t stream);
template <class T> cudaError_t fast_pad_plane(T* dst, const T* src, int nd, int PP, int ntiles, cudaStream_t stream);
// tensor maps of the scratch planes (nlev levels resident) and of the padded area array; returns cudaErrorNotSupported
// when the driver lacks cuTensorMapEncodeTiled
template <class T> cudaError_t fast_advect5_maps(Adv5Maps* m, const Adv5Params<T>& p, int nlev);
// p.tg is chosen by the launcher; nlev = levels of the resident chunk (grid.y)
template <class T> cudaError_t fast_advect5(Adv5Params<T> p, const Adv5Maps& m, int hord, int nlev, cudaStream_t stream);
// the reference's operation order, bit-identical to the FMA-free oracle (fv3t_exact.cu, -fmad=false); needs k_prep5 with exact = 1
template <class T> cudaError_t exact_advect5(Adv5Params<T> p, const Adv5Maps& m, int hord, int nlev, cudaStream_t stream);
// the same two for sub-tile contexts (fv3t_dims.sub_layout): the edge / corner flags of the resident sub-domains ride in p.sub
template <class T> cudaError_t fast_advect5_sub(Adv5ParamsSub<T> p, const Adv5Maps& m, int hord, int nlev, cudaStream_t stream);
template <class T> cudaError_t exact_advect5_sub(Adv5ParamsSub<T> p, const Adv5Maps& m, int hord, int nlev, cudaStream_t stream);

// ---- lanes-over-levels remap (fv3t_remap4.cuh): km <= 127, mapn_tracer form, uniform abs(kord) in fast_kord_ok ----
template <class T> size_t remap4_coef_bytes(int n, int ntiles);
template <class T> cudaError_t fast_remap_coef4(const Remap4Params<T>& p, cudaStream_t stream);
template <class T> cudaError_t fast_remap4(Remap4Params<T> p, int akord, cudaStream_t stream);

}  // namespace fv3t
