// Error norms of evaluate() / test() on the device: the reference copies the predictions to the host and
// calls numpy.linalg.norm (ev-NSFnet/pinn_solver.py:684-688, NSFnet/pinn_solver.py:321-325); here one kernel reduces
//   sum (u - u_pred)^2, sum u^2, sum (v - v_pred)^2, sum v^2, and -- over the points whose reference pressure is not NaN --
//   sum (p - p_pred)^2, sum p^2, the number of such points
// in fp64 (thread partials -> warp shuffles -> one atomicAdd per warp and sum).  HBM bound: 24 B per point.
#include "nsf_internal.h"

namespace {

__global__ void __launch_bounds__(256) nsf_error_norms_kernel(const float* __restrict__ pred, const float* __restrict__ u, const float* __restrict__ v,
                                                              const float* __restrict__ p, long long n, double* __restrict__ out) {
  double s[7] = {0, 0, 0, 0, 0, 0, 0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float pu = pred[3 * i], pv = pred[3 * i + 1], pp = pred[3 * i + 2];
    const double ur = u[i], vr = v[i];
    const double du = ur - (double)pu, dv = vr - (double)pv;
    s[0] += du * du; s[1] += ur * ur; s[2] += dv * dv; s[3] += vr * vr;
    const float pr = p ? p[i] : nanf("");
    if (pr == pr) {       // the reference masks NaN pressure (ev :684)
      const double dp = (double)pr - (double)pp;
      s[4] += dp * dp; s[5] += (double)pr * (double)pr; s[6] += 1.0;
    }
  }
#pragma unroll
  for (int k = 0; k < 7; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_down_sync(0xffffffffu, s[k], o);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 7; ++k) atomicAdd(out + k, s[k]);
  }
}

}  // namespace

extern "C" int nsf_error_norms(const float* uvp_pred, const float* u_ref, const float* v_ref, const float* p_ref_or_null, int64_t n,
                               double* sums8, void* stream) {
  if (n < 0 || !sums8 || (reinterpret_cast<uintptr_t>(sums8) & 7u) || (n > 0 && (!uvp_pred || !u_ref || !v_ref))) {
    nsf_set_error("nsf_error_norms: bad argument"); return NSF_E_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  NSF_CUDA_OK(cudaMemsetAsync(sums8, 0, 8 * sizeof(double), st));
  if (n == 0) return NSF_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  nsf_error_norms_kernel<<<(unsigned)blocks, 256, 0, st>>>(uvp_pred, u_ref, v_ref, p_ref_or_null, n, sums8);
  NSF_CUDA_OK(cudaGetLastError());
  return NSF_OK;
}
