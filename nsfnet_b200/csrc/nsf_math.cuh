// Branch-free fp32 tanh for the jet epilogues (CUDA's tanhf branches on |x| < 0.55, which diverges inside a warp).
//   |x| <  0.55 : x + x^3 q(x^2), q = degree-4 least-squares fit of (tanh x - x)/x^3 on [0, 0.6] (max rel. error 7.8e-8)
//   |x| >= 0.55 : 1 - 2 / (exp(2|x|) + 1) with ex2.approx / rcp-based division (rel. error <= ~3e-7 at the switch point)
#pragma once
#include <cuda_runtime.h>

__device__ __forceinline__ float nsf_tanh_fast(float x) {
  const float ax = fabsf(x);
  const float u = x * x;
  float p = fmaf(-6.276390147e-03f, u, 2.111621603e-02f);
  p = fmaf(p, u, -5.387513875e-02f);
  p = fmaf(p, u, 1.333292401e-01f);
  p = fmaf(p, u, -3.333333043e-01f);
  const float small = fmaf(x * u, p, x);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * 2.885390082f));   // exp(2|x|)
  const float big = copysignf(1.f - __fdividef(2.f, e + 1.f), x);
  return ax < 0.55f ? small : big;
}
