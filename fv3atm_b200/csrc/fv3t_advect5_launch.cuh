// fv3atm_b200: launchers of k_advect5, shared by the two translation units that instantiate it: fv3t_fast.cu (FMA contraction,
// shared reciprocals; hord 8 / 11 / 2) and fv3t_exact.cu (-fmad=false, the reference's operation order: every scheme, bit-identical
// to the FMA-free oracle).
#pragma once
#include <cstdlib>

#include "fv3t_advect5.cuh"

namespace fv3t {

template <class T, int OI, int OO, int TGC, bool EX, class P> static cudaError_t launch5(const P& p, const Adv5Maps& m, dim3 grid, cudaStream_t stream) {
  constexpr int NTHR = 32 + A5_GW * TGC;
  // the named barriers of the tracer groups take all 16 hardware barriers: one CTA per SM; a one-tracer CTA uses one (two CTAs per SM)
  constexpr int MINB = TGC == 1 ? 2 : 1;
  const size_t smem = A5Stage<T>::smem_bytes(p.tg);
  cudaError_t e = cudaFuncSetAttribute(k_advect5<T, OI, OO, NTHR, MINB, EX, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_advect5<T, OI, OO, NTHR, MINB, EX, P><<<grid, 32 + A5_GW * p.tg, smem, stream>>>(p, m);
  return cudaGetLastError();
}
// P = Adv5Params<T> (whole tiles) or Adv5ParamsSub<T> (sub-tile contexts: one instantiation per scheme, nine tracer groups)
template <class T, int OI, int OO, bool EX, class P> static cudaError_t launch5_ord(P& p, const Adv5Maps& m, int nlev, cudaStream_t stream) {
  // tracers per CTA: every tracer group of a CTA shares the staged level fields.  FV3T_ADV_TG caps it (tuning knob).
  static const int cap_env = getenv("FV3T_ADV_TG") ? atoi(getenv("FV3T_ADV_TG")) : A5_MAXTG;
  const int cap = cap_env < 1 ? 1 : (cap_env > A5_MAXTG ? A5_MAXTG : cap_env);
  const int chunks = (p.nql + cap - 1) / cap;
  p.tg = (p.nql + chunks - 1) / chunks;
  const int strips = (p.n + A5_W - 1) / A5_W;
  dim3 grid(strips, nlev, chunks);
  if (EX || P::SUB) return launch5<T, OI, OO, 9, EX>(p, m, grid, stream);  // one instantiation per scheme keeps the build time bounded
  if (p.tg == 1) return launch5<T, OI, OO, 1, false>(p, m, grid, stream);
  if (p.tg > 5) return launch5<T, OI, OO, 9, false>(p, m, grid, stream);
  return launch5<T, OI, OO, 5, false>(p, m, grid, stream);
}

}  // namespace fv3t
