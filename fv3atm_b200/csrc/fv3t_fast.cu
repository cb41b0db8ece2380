// fv3atm_b200: translation unit of the production ("fast") kernels -- compiled WITHOUT -fmad=false (see fv3t_fast.h).
#include "fv3t_fast.h"

#include <cstdlib>
#include <cudaTypedefs.h>

#include "fv3t_advect5_launch.cuh"

namespace fv3t {

template <class T> cudaError_t fast_prep3(const Prep3Params<T>& p, cudaStream_t stream) {
  dim3 grid(32, p.ntiles * p.npz);
  k_prep3<T><<<grid, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

template <class T>
cudaError_t fast_scale3(T* cx, T* cy, T* mfx, T* mfy, const int* ksplt, int n, int npz, int ntiles, cudaStream_t stream) {
  dim3 grid(16, ntiles * npz);
  k_scale3<T><<<grid, 256, 0, stream>>>(cx, cy, mfx, mfy, ksplt, n, npz);
  return cudaGetLastError();
}

template <class T> cudaError_t fast_cab3(const Cab3Params<T>& p, int ntiles, cudaStream_t stream) {
  dim3 grid(16, ntiles * p.npz);
  k_cab3<T><<<grid, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

template <class T, int OI, int OO, int NTC, int MINB> static cudaError_t launch4(const Adv3Params<T>& p, int NT, dim3 grid, cudaStream_t stream) {
  constexpr size_t smem = (size_t)Adv4Layout<NTC>::TOTAL * sizeof(T);
  if (smem > 48 * 1024) {  // per device and cheap: set on every launch (a process may hold contexts on several GPUs)
    cudaError_t e = cudaFuncSetAttribute(k_advect4<T, OI, OO, NTC, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k_advect4<T, OI, OO, NTC, MINB><<<grid, NT, smem, stream>>>(p);
  return cudaGetLastError();
}

template <class T, int OI, int OO> static cudaError_t launch3(Adv3Params<T>& p, int NT, cudaStream_t stream) {
  p.W = NT - 6;
  const int strips = (p.n + p.W - 1) / p.W;
  const int nql = p.nql < 0 ? p.nq - p.iq0 : p.nql;
  // FV3T_ADV_RING=0 selects the register-prefetch kernel (k_advect3); default: the asynchronous-copy ring (k_advect4)
  static const int ring = getenv("FV3T_ADV_RING") ? atoi(getenv("FV3T_ADV_RING")) : 1;
  // tracers per thread of k_advect3 (tuning knob).  B200, C768: G = 1 131 ms, G = 2 177 ms, G = 3 211 ms (255 registers -> 8 warps/SM):
  // sharing the level fields does not pay for the lost occupancy (profiles/r01_advect3_block_sweep.txt)
  static const int grp = getenv("FV3T_ADV_G") ? atoi(getenv("FV3T_ADV_G")) : 1;
  if (ring) {
    dim3 grid(nql, strips, p.ntiles * p.npz);
    if (NT <= 64) return launch4<T, OI, OO, 64, 8>(p, NT, grid, stream);
    return launch4<T, OI, OO, 256, 2>(p, NT, grid, stream);
  }
  if (grp == 2) {
    dim3 grid((nql + 1) / 2, strips, p.ntiles * p.npz);
    k_advect3<T, OI, OO, 2, 1><<<grid, NT, 0, stream>>>(p);
  } else {
    dim3 grid(nql, strips, p.ntiles * p.npz);
    k_advect3<T, OI, OO, 1, 2><<<grid, NT, 0, stream>>>(p);
  }
  return cudaGetLastError();
}

template <class T> cudaError_t fast_advect3(Adv3Params<T> p, int hord, int NT, cudaStream_t stream) {
  switch (hord) {
    case 8: return launch3<T, 8, 8>(p, NT, stream);
    case 10: return launch3<T, 8, 10>(p, NT, stream);  // ord_in = 8 when hord == 10 (tp_core.F90:157-161)
    case 9: return launch3<T, 9, 9>(p, NT, stream);
    case 11: return launch3<T, 11, 11>(p, NT, stream);
    case 12: return launch3<T, 12, 12>(p, NT, stream);
    case 13: return launch3<T, 13, 13>(p, NT, stream);
    case 2: return launch3<T, 2, 2>(p, NT, stream);
    default: return cudaErrorInvalidValue;
  }
}


// ---- k_advect5 -------------------------------------------------------------------------------------------------------------
template <class T> cudaError_t fast_prep5(const Prep5Params<T>& p, cudaStream_t stream) {
  dim3 grid(32, p.nlev);
  k_prep5<T><<<grid, 256, 0, stream>>>(p);
  return cudaGetLastError();
}
template <class T> cudaError_t fast_pad_plane(T* dst, const T* src, int nd, int PP, int ntiles, cudaStream_t stream) {
  k_pad_plane<T><<<296, 256, 0, stream>>>(dst, src, nd, PP, ntiles);
  return cudaGetLastError();
}

typedef CUresult (*fv3t_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static fv3t_encode_fn tensor_map_encoder() {
  static fv3t_encode_fn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (fv3t_encode_fn)f;
  }
  return fn;
}
// [planes][nd rows][PP * per columns] of T, box = (A5_GW * per) x A5_R x 1, out-of-bounds elements read as zero
template <class T> static cudaError_t encode_plane_map(CUtensorMap* m, const void* base, int PP, int nd, int planes, int per) {
  fv3t_encode_fn enc = tensor_map_encoder();
  if (!enc) return cudaErrorNotSupported;
  const cuuint64_t gdim[3] = {(cuuint64_t)PP * per, (cuuint64_t)nd, (cuuint64_t)planes};
  const cuuint64_t gstr[2] = {(cuuint64_t)PP * per * sizeof(T), (cuuint64_t)PP * per * sizeof(T) * nd};
  const cuuint32_t box[3] = {(cuuint32_t)(per == 2 ? A5_GW * 2 : A5Stage<T>::SW), (cuuint32_t)A5_R, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(m, sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), gdim, gstr,
                         box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}
template <class T> cudaError_t fast_advect5_maps(Adv5Maps* m, const Adv5Params<T>& p, int nlev) {
  const int nd = p.n + 6, PP = a5_pitch(p.n);
  cudaError_t e;
  if ((e = encode_plane_map<T>(&m->x2, p.X2, PP, nd, nlev, 2)) != cudaSuccess) return e;
  if ((e = encode_plane_map<T>(&m->y2, p.Y2, PP, nd, nlev, 2)) != cudaSuccess) return e;
  if ((e = encode_plane_map<T>(&m->cab, p.CAB, PP, nd, nlev, 2)) != cudaSuccess) return e;
  if ((e = encode_plane_map<T>(&m->rx, p.RX, PP, nd, nlev, 1)) != cudaSuccess) return e;
  if ((e = encode_plane_map<T>(&m->ry, p.RY, PP, nd, nlev, 1)) != cudaSuccess) return e;
  if ((e = encode_plane_map<T>(&m->mfx, p.MFX, PP, nd, nlev, 1)) != cudaSuccess) return e;
  if ((e = encode_plane_map<T>(&m->mfy, p.MFY, PP, nd, nlev, 1)) != cudaSuccess) return e;
  if ((e = encode_plane_map<T>(&m->area, p.AREA, PP, nd, p.ntiles, 1)) != cudaSuccess) return e;
  return encode_plane_map<T>(&m->rarea, p.RAREA, PP, nd, p.ntiles, 1);
}

template <class T, class P> static cudaError_t fast_dispatch(P p, const Adv5Maps& m, int hord, int nlev, cudaStream_t stream) {
  switch (hord) {
    case 8: return launch5_ord<T, 8, 8, false>(p, m, nlev, stream);
    case 11: return launch5_ord<T, 11, 11, false>(p, m, nlev, stream);
    case 2: return launch5_ord<T, 2, 2, false>(p, m, nlev, stream);
    default: return cudaErrorInvalidValue;
  }
}
template <class T> cudaError_t fast_advect5(Adv5Params<T> p, const Adv5Maps& m, int hord, int nlev, cudaStream_t stream) {
  return fast_dispatch<T>(p, m, hord, nlev, stream);
}
template <class T> cudaError_t fast_advect5_sub(Adv5ParamsSub<T> p, const Adv5Maps& m, int hord, int nlev, cudaStream_t stream) {
  return fast_dispatch<T>(p, m, hord, nlev, stream);
}

// ---- k_remap4 ---------------------------------------------------------------------------------------------------------------
template <class T> size_t remap4_coef_bytes(int n, int ntiles) { return (size_t)ntiles * n * r4_groups_per_row(n) * R4Block<T>::BYTES; }
template <class T> cudaError_t fast_remap_coef4(const Remap4Params<T>& p, cudaStream_t stream) {
  dim3 grid((p.n * p.n + 127) / 128, p.ntiles);
  k_remap_coef4<T><<<grid, 128, 0, stream>>>(p);
  return cudaGetLastError();
}
template <class T, int AK, int LPL> static cudaError_t launch_remap4(Remap4Params<T>& p, cudaStream_t stream) {
  static const int cap_env = getenv("FV3T_REMAP_TW") ? atoi(getenv("FV3T_REMAP_TW")) : R4_MAXW;  // tracer warps per CTA (tuning knob)
  const int cap = cap_env < 1 ? 1 : (cap_env > R4_MAXW ? R4_MAXW : cap_env);
  const int chunks = (p.nql + cap - 1) / cap;
  p.ntw = (p.nql + chunks - 1) / chunks;
  const long total = (long)p.ntiles * p.n * r4_groups_per_row(p.n);
  static const int waves = getenv("FV3T_REMAP_WAVES") ? atoi(getenv("FV3T_REMAP_WAVES")) : 8;
  long ctas = 148L * (waves < 1 ? 1 : waves);
  if (ctas > total) ctas = total;
  p.groups_per_cta = (int)((total + ctas - 1) / ctas);
  ctas = (total + p.groups_per_cta - 1) / p.groups_per_cta;
  constexpr int CB = (R4Block<T>::BYTES + 127) & ~127;
  constexpr int WB = (R4Warp<T, LPL>::BYTES + 127) & ~127;
  const size_t smem = (((size_t)2 * CB + 128 + 2 * 132 * sizeof(T) + 127) & ~(size_t)127) + (size_t)p.ntw * WB;
  cudaError_t e = cudaFuncSetAttribute(k_remap4<T, AK, LPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid((unsigned)ctas, chunks);
  k_remap4<T, AK, LPL><<<grid, 32 * (p.ntw + 1), smem, stream>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess || !p.fill) return e;
  dim3 gf((p.n * p.n + 127) / 128, p.ntiles, p.nql);
  k_fillz4<T><<<gf, 128, 0, stream>>>(p);
  return cudaGetLastError();
}
template <class T, int AK> static cudaError_t launch_remap4_lpl(Remap4Params<T>& p, cudaStream_t stream) {
  if (p.km <= 63) return launch_remap4<T, AK, 8>(p, stream);
  return launch_remap4<T, AK, 16>(p, stream);
}
// Experimental (FV3T_REMAP4=1): instantiated for abs(kord) = 9 only -- on B200 the lanes-over-levels kernel is SLOWER than k_remap3
// (160 ms vs 89 ms at C768 L127 x9 fp64, profiles/r02_remap4_v2_ncu.txt): its 4.4 KB of shared memory per column-tracer in
// flight leaves ten warps per SM, too few for its dependent shuffle / shared-memory chains.
template <class T> cudaError_t fast_remap4(Remap4Params<T> p, int akord, cudaStream_t stream) {
  if (p.km > 127 || akord != 9) return cudaErrorInvalidValue;
  return launch_remap4_lpl<T, 9>(p, stream);
}

template <class T> static bool remap3_offsets_fit(const Remap3Params<T>& p);
// FV3T_REMAP5=1 selects the two-walk kernel pair of fv3t_remap5.cuh (parity-green, but SLOWER on B200: 84.9 ms against 76.6 ms at
// C768 L127 x9 fp64 -- see the header); the default is the three-walk pair k_remap_coef3 / k_remap3.  Both launchers consult the
// same flag: the coefficient arrays mean different things to the two pairs.
static bool remap_two_walk() {
  static const bool v = getenv("FV3T_REMAP5") && atoi(getenv("FV3T_REMAP5")) != 0;
  return v;
}
template <class T> cudaError_t fast_remap_coef3(const Remap3Params<T>& p, cudaStream_t stream) {
  if (!remap3_offsets_fit(p)) return cudaErrorInvalidValue;
  dim3 grid((p.n * p.n + 127) / 128, p.ntiles);
  if (remap_two_walk() && p.km >= 4)
    k_remap_coef5<T><<<grid, 128, 0, stream>>>(p);
  else
    k_remap_coef3<T><<<grid, 128, 0, stream>>>(p);
  return cudaGetLastError();
}

// the remap kernels index inside a column with 32-bit offsets
template <class T> static bool remap3_offsets_fit(const Remap3Params<T>& p) {
  return (long)(p.km + 1) * (p.n + 6) * (p.n + 6) < (1L << 31) && (long)(p.km + 1) * (p.n + 2) < (1L << 31);
}

template <class T, int AK> static cudaError_t launch_remap3(const Remap3Params<T>& p, cudaStream_t stream) {
  if (!remap3_offsets_fit(p)) return cudaErrorInvalidValue;
  dim3 grid(p.nql < 0 ? p.nq - p.iq0 : p.nql, (p.n * p.n + 127) / 128, p.ntiles);
  // CTAs per SM: fp64 4 (128 registers, no spills; 5 spills and is slower), fp32 6 (80 registers); FV3T_REMAP_MINB overrides
  static const int minb = getenv("FV3T_REMAP_MINB") ? atoi(getenv("FV3T_REMAP_MINB")) : (sizeof(T) == 4 ? 6 : 4);
  if (remap_two_walk() && p.km >= 4) {
    if (minb == 5)
      k_remap5<T, AK, 128, 5><<<grid, 128, 0, stream>>>(p);
    else
      k_remap5<T, AK, 128, 4><<<grid, 128, 0, stream>>>(p);
    return cudaGetLastError();
  }
  if (minb == 5)
    k_remap3<T, AK, true, 128, 5><<<grid, 128, 0, stream>>>(p);
  else if (minb == 6)
    k_remap3<T, AK, true, 128, 6><<<grid, 128, 0, stream>>>(p);
  else
    k_remap3<T, AK, true, 128, 4><<<grid, 128, 0, stream>>>(p);
  return cudaGetLastError();
}

template <class T> cudaError_t fast_remap3(const Remap3Params<T>& p, int akord, cudaStream_t stream) {
  if (akord <= 8) return launch_remap3<T, 8>(p, stream);
  if (akord >= 17) return launch_remap3<T, 17>(p, stream);
  switch (akord) {
    case 9: return launch_remap3<T, 9>(p, stream);
    case 12: return launch_remap3<T, 12>(p, stream);
    case 13: return launch_remap3<T, 13>(p, stream);
    case 14: return launch_remap3<T, 14>(p, stream);
    default: return cudaErrorInvalidValue;
  }
}

#define FV3T_FAST_INST(T)                                                                                             \
  template cudaError_t fast_prep3<T>(const Prep3Params<T>&, cudaStream_t);                                            \
  template cudaError_t fast_scale3<T>(T*, T*, T*, T*, const int*, int, int, int, cudaStream_t);                       \
  template cudaError_t fast_cab3<T>(const Cab3Params<T>&, int, cudaStream_t);                                         \
  template cudaError_t fast_advect3<T>(Adv3Params<T>, int, int, cudaStream_t);                                        \
  template cudaError_t fast_prep5<T>(const Prep5Params<T>&, cudaStream_t);                                            \
  template cudaError_t fast_pad_plane<T>(T*, const T*, int, int, int, cudaStream_t);                                  \
  template cudaError_t fast_advect5_maps<T>(Adv5Maps*, const Adv5Params<T>&, int);                                    \
  template cudaError_t fast_advect5<T>(Adv5Params<T>, const Adv5Maps&, int, int, cudaStream_t);                       \
  template cudaError_t fast_advect5_sub<T>(Adv5ParamsSub<T>, const Adv5Maps&, int, int, cudaStream_t);                \
  template size_t remap4_coef_bytes<T>(int, int);                                                                     \
  template cudaError_t fast_remap_coef4<T>(const Remap4Params<T>&, cudaStream_t);                                     \
  template cudaError_t fast_remap4<T>(Remap4Params<T>, int, cudaStream_t);                                            \
  template cudaError_t fast_remap_coef3<T>(const Remap3Params<T>&, cudaStream_t);                                     \
  template cudaError_t fast_remap3<T>(const Remap3Params<T>&, int, cudaStream_t);
#if defined(FV3T_INST_F64) || !defined(FV3T_INST_F32)
FV3T_FAST_INST(double)
#endif
#if defined(FV3T_INST_F32) || !defined(FV3T_INST_F64)
FV3T_FAST_INST(float)
#endif

}  // namespace fv3t
