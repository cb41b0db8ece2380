// fv3atm_b200: translation unit of the production ("fast") kernels -- compiled WITHOUT -fmad=false (see fv3t_fast.h).
#include "fv3t_fast.h"

#include <cstdlib>

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

template <class T> cudaError_t fast_remap_coef3(const Remap3Params<T>& p, cudaStream_t stream) {
  dim3 grid((p.n * p.n + 127) / 128, p.ntiles);
  k_remap_coef3<T><<<grid, 128, 0, stream>>>(p);
  return cudaGetLastError();
}

template <class T, int AK> static cudaError_t launch_remap3(const Remap3Params<T>& p, cudaStream_t stream) {
  dim3 grid(p.nql < 0 ? p.nq - p.iq0 : p.nql, (p.n * p.n + 127) / 128, p.ntiles);
  static const int minb = getenv("FV3T_REMAP_MINB") ? atoi(getenv("FV3T_REMAP_MINB")) : 4;  // tuning knob: 4 -> 128 regs, 5 -> 96, 6 -> 80
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
  template cudaError_t fast_remap_coef3<T>(const Remap3Params<T>&, cudaStream_t);                                     \
  template cudaError_t fast_remap3<T>(const Remap3Params<T>&, int, cudaStream_t);
FV3T_FAST_INST(double)
FV3T_FAST_INST(float)

}  // namespace fv3t
