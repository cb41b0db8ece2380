// fv3atm_b200: support kernels of the horizontal tracer advection (tracer_2d) for sm_100a: per-level Courant maxima,
// intra-GPU edge-halo fill, level gather.  The transport kernels are fv3t_advect2.cuh (strict), fv3t_advect3.cuh /
// fv3t_advect4.cuh (fast: register prefetch / cp.async input ring).
//
// Data layout in HBM = the Fortran layout of the host arrays, tile-major:
//   q   (isd:ied, jsd:jed, npz, nq)  two buffers (ping-pong: a sub-step reads one and writes the other,
//                                    because neighbouring CTAs read each other's cells through the halo)
//   dp1 (isd:ied, jsd:jed, npz)   cx (is:ie+1, jsd:jed, npz)   cy (isd:ied, js:je+1, npz)
//   mfx (is:ie+1, js:je, npz)     mfy (is:ie, js:je+1, npz)
#pragma once
#include "fv3t_ppm.cuh"

namespace fv3t {

template <class T> struct GridDev {
  const T *area, *rarea, *dx, *dy, *dxa, *dya, *sin_sg;  // tile-major, Fortran extents
};

// ---------------------------------------------------------------------------------------------------
// cmax(k) per tile (fv_tracer2d.F90:409-424).  One CTA per (tile, level); max is order independent.
// ---------------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(512) k_cmax(const T* __restrict__ cx, const T* __restrict__ cy, const T* __restrict__ sin_sg,
                                             T* __restrict__ cmax_t, int n, int npz) {
  const int t = blockIdx.x / npz, kz = blockIdx.x % npz;
  const long nd = n + 6;
  const T* cxp = cx + ((long)t * npz + kz) * (long)(n + 1) * nd;
  const T* cyp = cy + ((long)t * npz + kz) * nd * (long)(n + 1);
  const T* s5 = sin_sg + (long)t * nd * nd * 5 + 4 * nd * nd;
  const bool upper = (kz + 1) < npz / 6;
  T cm = T(0);
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    const int j = idx / n + 1, i = idx % n + 1;
    const T ax = f_abs(cxp[(long)(j + 2) * (n + 1) + (i - 1)]);
    const T ay = f_abs(cyp[(long)(j - 1) * nd + (i + 2)]);
    if (upper)
      cm = f_max(cm, ax, ay);
    else
      cm = f_max(cm, f_max(ax, ay) + T(1) - s5[(long)(j + 2) * nd + (i + 2)]);
  }
  __shared__ T red[512];
  red[threadIdx.x] = cm;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] = f_max(red[threadIdx.x], red[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) cmax_t[blockIdx.x] = red[0];
}

// Scalar edge-halo fill of q for all tracers and the levels still active at sub-step `it`
// (complete_group_halo_update, fv_tracer2d.F90:499).  dst/src: flat offsets into the tile-major stack of
// (n+6)^2 planes, built from the 12-contact mosaic (fv_mp_mod.F90:581-629).
template <class T>
__global__ void k_halo_fill(T* __restrict__ q, const int* __restrict__ dst, const int* __restrict__ src, int len, int n, int npz,
                            int nq, const int* __restrict__ ksplt, int it, int pl0 = 0) {
  const long plane = (long)(n + 6) * (n + 6);
  const long tile_stride = plane * npz * nq;
  const int pl = pl0 + blockIdx.y;  // iq*npz + kz (pl0: first plane of a tracer sub-range)
  const int kz = pl % npz;
  if (it > ksplt[kz]) return;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < len; e += gridDim.x * blockDim.x) {
    const int d = dst[e], s = src[e];
    const long dt = d / plane, dr = d % plane, st = s / plane, sr = s % plane;
    q[dt * tile_stride + (long)pl * plane + dr] = q[st * tile_stride + (long)pl * plane + sr];
  }
}

// Move levels that an earlier call left in buffer 1 back to buffer 0 (only when tracer_2d is called twice
// without a remap in between).
template <class T>
__global__ void k_copy_levels(T* __restrict__ q0, const T* __restrict__ q1, const int* __restrict__ par, long plane, int npz,
                              long total) {
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int kz = (int)((e / plane) % npz);
    if (par[kz]) q0[e] = q1[e];
  }
}

}  // namespace fv3t
