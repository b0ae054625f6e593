// Step-level kernels around the jet kernel: device-resident Adam (hyper-parameters and the step
// counter live on the device so ONE captured CUDA graph replays every training iteration), and the on-device point
// layer: Latin-hypercube collocation points, distance to the nearest boundary point, SDF weights.
//
// The per-element math is written once (NSF_HD) and driven by CUDA kernels in the product build and by host loops in
// the tests/emu build (-DNSF_EMU), like the FFMA kernels.
#include "nsf_internal.h"

#include <cmath>

namespace {

// ---- Adam ------------------------------------------------------------------------------------------------------
// torch.optim.Adam without amsgrad / weight decay (ev-NSFnet/pinn_solver.py:126-129):
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps),  bc_k = 1 - b_k^t
NSF_HD void adam_update(float& p, float g, float& m, float& v, float lr, float b1, float b2, float eps, float bc1, float bc2, float gs) {
  const float gi = g * gs;
  const float mi = b1 * m + (1.f - b1) * gi;
  const float vi = b2 * v + (1.f - b2) * gi * gi;
  m = mi; v = vi;
  const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
  p -= (lr / bc1) * (mi / denom);
}

// ---- counter-based randomness -----------------------------------------------------------------------------------
NSF_HD uint32_t mix32(uint32_t x) {   // lowbias32 finaliser
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
NSF_HD uint32_t hash3(uint32_t a, uint32_t b, uint32_t c) { return mix32(a ^ mix32(b ^ mix32(c ^ 0x9e3779b9U))); }

// Pseudo-random permutation of [0, n): 6-round balanced Feistel network on 2*half bits, cycle-walked into range.
// Stateless and O(1) memory, so every rank can produce any row range of the SAME global design.
NSF_HD uint64_t feistel_perm(uint64_t i, uint64_t n, int half, uint32_t key) {
  const uint32_t mask = (1u << half) - 1u;
  uint64_t v = i;
  do {
    uint32_t l = (uint32_t)(v >> half) & mask, r = (uint32_t)v & mask;
#pragma unroll
    for (int round = 0; round < 6; ++round) {
      const uint32_t t = l ^ (hash3(r, key, (uint32_t)round) & mask);
      l = r; r = t;
    }
    v = ((uint64_t)l << half) | r;
  } while (v >= n);
  return v;
}
NSF_HD int feistel_half_bits(uint64_t n) {
  int bits = 1;
  while (((uint64_t)1 << bits) < n) ++bits;
  return (bits + 1) / 2 < 1 ? 1 : (bits + 1) / 2;
}
// uniform in [0, 1) with 24 random bits
NSF_HD float u01(uint32_t h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }

// one Latin-hypercube coordinate: stratum perm(i) of n, uniform inside the stratum (tools.py:30-57), in double so that
// the stratum survives the rounding to fp32 for n up to 2^24 strata per unit length
NSF_HD float lhs_coord(uint64_t i, uint64_t n, int half, uint32_t key, double lo, double hi) {
  const uint64_t cell = feistel_perm(i, n, half, key);
  const double u = (double)u01(hash3((uint32_t)i, (uint32_t)(i >> 32) ^ key, 0x51ed270bU));
  const double t = ((double)cell + u) / (double)n;
  float tf = (float)t;
  // fp32 rounding must not leave the stratum (as long as the stratum is wider than an fp32 ulp)
  if ((double)tf * (double)n >= (double)(cell + 1)) tf = nextafterf(tf, 0.f);
  else if ((double)tf * (double)n < (double)cell) tf = nextafterf(tf, 2.f);
  return (float)(lo + (double)tf * (hi - lo));
}

NSF_HD float sdf_weight(float d, float min_w, float decay) { return min_w + (1.f - min_w) * expf(-decay * d); }

}  // namespace

#ifndef NSF_EMU
// ================================================= CUDA =================================================
namespace {

__global__ void nsf_adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                    long long n, const NsfAdamDev* __restrict__ st) {
  __shared__ float s_bc[2];
  const NsfAdamDev h = *st;
  if (threadIdx.x == 0) {
    const double t = (double)(h.step + 1);
    s_bc[0] = (float)(1.0 - pow((double)h.beta1, t));
    s_bc[1] = (float)(1.0 - pow((double)h.beta2, t));
  }
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float pi = p[i], mi = m[i], vi = v[i];
  adam_update(pi, g[i], mi, vi, h.lr, h.beta1, h.beta2, h.eps, s_bc[0], s_bc[1], h.grad_scale);
  p[i] = pi; m[i] = mi; v[i] = vi;
}

__global__ void nsf_adam_tick_kernel(NsfAdamDev* st) { st->step += 1; }

__global__ void nsf_lhs_kernel(long long n_total, long long first, long long count, uint32_t seed, int half, double x0, double x1,
                               double y0, double y1, float* __restrict__ x, float* __restrict__ y) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const uint64_t i = (uint64_t)(first + t);
  x[t] = lhs_coord(i, (uint64_t)n_total, half, seed * 2u + 0x1234567u, x0, x1);
  y[t] = lhs_coord(i, (uint64_t)n_total, half, seed * 2u + 0x89abcdeu, y0, y1);
}

// distance to the nearest of nb boundary points (what cKDTree.query returns, cavity_data.py:118-121): the boundary set is
// staged through shared memory in tiles, every thread scans it for its own point (nb = 2052: 2 k distance evaluations per point)
constexpr int WD_TILE = 1024;
__global__ void nsf_wall_distance_kernel(const float* __restrict__ x, const float* __restrict__ y, long long n,
                                         const float* __restrict__ xb, const float* __restrict__ yb, int nb,
                                         float* __restrict__ dist, float* __restrict__ w, float min_w, float decay, double* __restrict__ w_sum) {
  __shared__ float2 sb[WD_TILE];
  __shared__ double s_part[8];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = i < n;
  const float px = ok ? x[i] : 0.f, py = ok ? y[i] : 0.f;
  float best = INFINITY;
  for (int b0 = 0; b0 < nb; b0 += WD_TILE) {
    const int cnt = nb - b0 < WD_TILE ? nb - b0 : WD_TILE;
    __syncthreads();
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) sb[k] = make_float2(xb[b0 + k], yb[b0 + k]);
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < cnt; ++k) {
      const float dx = sb[k].x - px, dy = sb[k].y - py;
      best = fminf(best, fmaf(dx, dx, dy * dy));
    }
  }
  const float d = sqrtf(best);
  float wi = 0.f;
  if (ok) {
    if (dist) dist[i] = d;
    if (w) { wi = sdf_weight(d, min_w, decay); w[i] = wi; }
  }
  if (w_sum) {   // block sum in double -> one atomic per block
    double acc = (double)wi;
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int k = 0; k < (int)(blockDim.x >> 5); ++k) tot += s_part[k];
      atomicAdd(w_sum, tot);
    }
  }
}

}  // namespace

static int adam_dev_run(float* p, const float* g, float* m, float* v, long long n, NsfAdamDev* st, nsf_stream_t s) {
  if (n > 0) {
    nsf_adam_dev_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, g, m, v, n, st);
    NSF_CUDA_OK(cudaGetLastError());
  }
  return NSF_OK;
}
static int adam_tick_run(NsfAdamDev* st, nsf_stream_t s) {
  nsf_adam_tick_kernel<<<1, 1, 0, s>>>(st);
  NSF_CUDA_OK(cudaGetLastError());
  return NSF_OK;
}
static int lhs_run(long long n_total, long long first, long long count, uint32_t seed, double x0, double x1, double y0, double y1,
                   float* x, float* y, nsf_stream_t s) {
  if (count <= 0) return NSF_OK;
  nsf_lhs_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(n_total, first, count, seed, feistel_half_bits((uint64_t)n_total), x0, x1, y0, y1, x, y);
  NSF_CUDA_OK(cudaGetLastError());
  return NSF_OK;
}
static int wall_run(const float* x, const float* y, long long n, const float* xb, const float* yb, int nb, float* dist, float* w,
                    float min_w, float decay, double* w_sum, nsf_stream_t s) {
  if (n <= 0) return NSF_OK;
  nsf_wall_distance_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, y, n, xb, yb, nb, dist, w, min_w, decay, w_sum);
  NSF_CUDA_OK(cudaGetLastError());
  return NSF_OK;
}

#else
// ============================================ host emulation (tests/emu) ============================================
static int adam_dev_run(float* p, const float* g, float* m, float* v, long long n, NsfAdamDev* st, nsf_stream_t) {
  const NsfAdamDev h = *st;
  const double t = (double)(h.step + 1);
  const float bc1 = (float)(1.0 - std::pow((double)h.beta1, t)), bc2 = (float)(1.0 - std::pow((double)h.beta2, t));
  for (long long i = 0; i < n; ++i) adam_update(p[i], g[i], m[i], v[i], h.lr, h.beta1, h.beta2, h.eps, bc1, bc2, h.grad_scale);
  return NSF_OK;
}
static int adam_tick_run(NsfAdamDev* st, nsf_stream_t) { st->step += 1; return NSF_OK; }
static int lhs_run(long long n_total, long long first, long long count, uint32_t seed, double x0, double x1, double y0, double y1,
                   float* x, float* y, nsf_stream_t) {
  const int half = feistel_half_bits((uint64_t)n_total);
  for (long long t = 0; t < count; ++t) {
    const uint64_t i = (uint64_t)(first + t);
    x[t] = lhs_coord(i, (uint64_t)n_total, half, seed * 2u + 0x1234567u, x0, x1);
    y[t] = lhs_coord(i, (uint64_t)n_total, half, seed * 2u + 0x89abcdeu, y0, y1);
  }
  return NSF_OK;
}
static int wall_run(const float* x, const float* y, long long n, const float* xb, const float* yb, int nb, float* dist, float* w,
                    float min_w, float decay, double* w_sum, nsf_stream_t) {
  double tot = 0.0;
  for (long long i = 0; i < n; ++i) {
    float best = INFINITY;
    for (int k = 0; k < nb; ++k) {
      const float dx = xb[k] - x[i], dy = yb[k] - y[i];
      best = std::fmin(best, std::fmaf(dx, dx, dy * dy));
    }
    const float d = std::sqrt(best);
    if (dist) dist[i] = d;
    if (w) { w[i] = sdf_weight(d, min_w, decay); tot += (double)w[i]; }
  }
  if (w_sum) *w_sum += tot;
  return NSF_OK;
}
#endif

// out[i] = scale * |in[i]|: the lag state of `init_vis_t` (ev :138-140) from the net_1 output this step computed anyway
#ifndef NSF_EMU
namespace {
__global__ void nsf_scale_abs_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, float scale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = scale * fabsf(in[i]);
}
}  // namespace
int nsf_scale_abs_launch(const float* in, float* out, long long n, float scale, nsf_stream_t st) {
  if (n <= 0) return NSF_OK;
  nsf_scale_abs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, n, scale);
  NSF_CUDA_OK(cudaGetLastError());
  return NSF_OK;
}
#else
int nsf_scale_abs_launch(const float* in, float* out, long long n, float scale, nsf_stream_t) {
  for (long long i = 0; i < n; ++i) out[i] = scale * std::fabs(in[i]);
  return NSF_OK;
}
#endif

static int ok_ptr(const void* p) { return p != nullptr && (reinterpret_cast<uintptr_t>(p) & 3u) == 0; }

extern "C" int nsf_adam_dev(float* params, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, NsfAdamDev* state,
                            void* stream) {
  if (n < 0 || !state || (n > 0 && (!ok_ptr(params) || !ok_ptr(grad) || !ok_ptr(exp_avg) || !ok_ptr(exp_avg_sq)))) {
    nsf_set_error("nsf_adam_dev: bad argument"); return NSF_E_ARG;
  }
  return adam_dev_run(params, grad, exp_avg, exp_avg_sq, n, state, (nsf_stream_t)stream);
}

extern "C" int nsf_adam_tick(NsfAdamDev* state, void* stream) {
  if (!state) { nsf_set_error("nsf_adam_tick: null state"); return NSF_E_ARG; }
  return adam_tick_run(state, (nsf_stream_t)stream);
}

extern "C" int nsf_lhs_points(int64_t n_total, int64_t first, int64_t count, uint32_t seed, float x_min, float x_max, float y_min,
                              float y_max, float* x_out, float* y_out, void* stream) {
  if (n_total <= 0 || first < 0 || count < 0 || first + count > n_total || n_total > ((int64_t)1 << 40) ||
      (count > 0 && (!ok_ptr(x_out) || !ok_ptr(y_out))) || x_min > x_max || y_min > y_max) {
    nsf_set_error("nsf_lhs_points: bad argument"); return NSF_E_ARG;
  }
  return lhs_run(n_total, first, count, seed, x_min, x_max, y_min, y_max, x_out, y_out, (nsf_stream_t)stream);
}

extern "C" int nsf_wall_distance(const float* x, const float* y, int64_t n, const float* xb, const float* yb, int32_t n_b,
                                 float* dist_out, void* stream) {
  if (n < 0 || n_b <= 0 || !ok_ptr(xb) || !ok_ptr(yb) || (n > 0 && (!ok_ptr(x) || !ok_ptr(y) || !ok_ptr(dist_out)))) {
    nsf_set_error("nsf_wall_distance: bad argument"); return NSF_E_ARG;
  }
  return wall_run(x, y, n, xb, yb, n_b, dist_out, nullptr, 0.f, 0.f, nullptr, (nsf_stream_t)stream);
}

extern "C" int nsf_sdf_weights(const float* x, const float* y, int64_t n, const float* xb, const float* yb, int32_t n_b,
                               float min_weight, float decay, float* w_out, double* w_sum, void* stream) {
  if (n < 0 || n_b <= 0 || !ok_ptr(xb) || !ok_ptr(yb) || (n > 0 && (!ok_ptr(x) || !ok_ptr(y) || !ok_ptr(w_out))) ||
      (w_sum && (reinterpret_cast<uintptr_t>(w_sum) & 7u))) {
    nsf_set_error("nsf_sdf_weights: bad argument"); return NSF_E_ARG;
  }
  // the reference's clamps (cavity_data.py:123-126)
  const float mw = min_weight < 1e-6f ? 1e-6f : (min_weight > 1.f ? 1.f : min_weight);
  const float dc = decay < 0.f ? 0.f : decay;
  return wall_run(x, y, n, xb, yb, n_b, nullptr, w_out, mw, dc, w_sum, (nsf_stream_t)stream);
}
