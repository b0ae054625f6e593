// Per-(point, neuron) jet arithmetic shared by the tcgen05 jet kernels: streams (value, d/dx, d/dy, laplacian).
//   forward  : a0 = t = tanh z0, ax = d1 zx, ay = d1 zy, a_lap = d2 (zx^2 + zy^2) + d1 z_lap,   d1 = 1 - t^2, d2 = -2 t d1
//   reverse  : z-bar from a-bar and the stashed (t, zx, zy, z_lap)                              d3 = -2 d1 (1 - 3 t^2)
// (SURVEY 8a "math contract"; the reference obtains the same quantities by seven autograd sweeps, ev-NSFnet/pinn_solver.py:301-309.)
#pragma once
#include <cuda_runtime.h>
#include "nsf_math.cuh"

__device__ __forceinline__ void nsf_jet_fwd(const float z[4], float v[4]) {
  const float t = nsf_tanh_fast(z[0]);
  const float d1 = fmaf(-t, t, 1.f), d2 = -2.f * t * d1;
  v[0] = t; v[1] = d1 * z[1]; v[2] = d1 * z[2];
  v[3] = fmaf(d2, fmaf(z[1], z[1], z[2] * z[2]), d1 * z[3]);
}
// activations of a layer from its stashed (t, zx, zy, z_lap)
__device__ __forceinline__ void nsf_act_from_stash(const float4 s, float a[4]) {
  const float d1 = fmaf(-s.x, s.x, 1.f), d2 = -2.f * s.x * d1;
  a[0] = s.x; a[1] = d1 * s.y; a[2] = d1 * s.z;
  a[3] = fmaf(d2, fmaf(s.y, s.y, s.z * s.z), d1 * s.w);
}
// adjoint through tanh: ab (adjoint of the activations), stash of the layer -> zb
__device__ __forceinline__ void nsf_zbar_from(const float4 st, const float ab[4], float zb[4]) {
  const float t = st.x, zx = st.y, zy = st.z, zl = st.w;
  const float d1 = fmaf(-t, t, 1.f), d2 = -2.f * t * d1, d3 = -2.f * d1 * fmaf(-3.f * t, t, 1.f);
  const float q = fmaf(zx, zx, zy * zy);
  const float c = 2.f * ab[3] * d2;
  zb[3] = ab[3] * d1;
  zb[1] = fmaf(ab[1], d1, c * zx);
  zb[2] = fmaf(ab[2], d1, c * zy);
  zb[0] = fmaf(ab[0], d1, fmaf(ab[1] * d2, zx, fmaf(ab[2] * d2, zy, ab[3] * fmaf(d3, q, d2 * zl))));
}
