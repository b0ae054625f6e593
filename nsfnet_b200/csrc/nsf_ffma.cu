// FP32 FFMA kernels of libnsf_b200.so: the jet step / residual / value kernels (nsf_ffma_body.h),
// the parameter packer, the gradient-row reduction and a fused Adam.
#include "nsf_internal.h"
#include "nsf_ffma_body.h"

#include <vector>

int nsf_ffma_pt(int ns, int hp) { return (ns == 4 && hp > 80) ? 16 : 32; }

#ifndef NSF_EMU
// ---------------------------------------------------------------------------------------------
// CUDA
// ---------------------------------------------------------------------------------------------
// threads per CTA = (HP/4) * (PT/4): <4,32> serves HP <= 80 (160 threads), <4,16> HP <= 128 (128), <1,32> HP <= 128 (256)
template <int NS, int PT>
__global__ void __launch_bounds__(NS == 1 ? 256 : (PT == 32 ? 160 : 128)) nsf_ffma_kernel(const NsfKernelArgs a) {
  extern __shared__ __align__(16) float nsf_smem[];
  nsf_cta_program<NS, PT>(a, nsf_smem, (int)blockIdx.x, (int)gridDim.x, (int)blockDim.x);
}

template <int NS, int PT>
static int launch_t(const NsfKernelArgs& a, int grid, int nt, size_t smem, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    NSF_CUDA_OK(cudaFuncSetAttribute(nsf_ffma_kernel<NS, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  nsf_ffma_kernel<NS, PT><<<grid, nt, smem, st>>>(a);
  NSF_CUDA_OK(cudaGetLastError());
  return NSF_OK;
}

template <int NS, int PT>
static int occ_t(int nt, size_t smem) {
  int nb = 0;
  if (cudaFuncSetAttribute(nsf_ffma_kernel<NS, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, nsf_ffma_kernel<NS, PT>, nt, smem) != cudaSuccess) return 0;
  return nb;
}

int nsf_ffma_occupancy(int ns, int hp) {
  const int pt = nsf_ffma_pt(ns, hp);
  const int nt = (hp / 4) * (pt / 4);
  const size_t smem = (size_t)nsf_ffma_smem_floats(ns, pt, hp) * sizeof(float);
  if (ns == 4) return pt == 32 ? occ_t<4, 32>(nt, smem) : occ_t<4, 16>(nt, smem);
  return occ_t<1, 32>(nt, smem);
}

int nsf_ffma_launch(NsfKernelArgs& a, int ns, int grid, nsf_stream_t st) {
  const int hp = a.g.HP;
  const int pt = nsf_ffma_pt(ns, hp);
  const int nt = (hp / 4) * (pt / 4);
  const size_t smem = (size_t)nsf_ffma_smem_floats(ns, pt, hp) * sizeof(float);
  a.n_tiles = (int)((a.n + pt - 1) / pt);
  if (a.n_tiles <= 0 || grid <= 0) return NSF_OK;
  if (ns == 4) return pt == 32 ? launch_t<4, 32>(a, grid, nt, smem, st) : launch_t<4, 16>(a, grid, nt, smem, st);
  return launch_t<1, 32>(a, grid, nt, smem, st);
}

__global__ void nsf_pack_kernel(NsfNetGeom g, const float* __restrict__ flat, float* __restrict__ pk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < g.pk_size()) pk[i] = nsf_pack_value(g, flat, i);
}

int nsf_pack_launch(const NsfNetGeom& g, const float* flat, float* pk, nsf_stream_t st) {
  const int n = g.pk_size();
  nsf_pack_kernel<<<(n + 255) / 256, 256, 0, st>>>(g, flat, pk);
  NSF_CUDA_OK(cudaGetLastError());
  return NSF_OK;
}

// grad[i] = sum over rows of scratch[row][map[i]]; loss_parts[s] likewise.  Rows [0, split) may use another layout of the row
// (map0: the tcgen05 kernel writes its hidden-layer weight gradients in the order it drains tensor memory).
// A warp covers 32 consecutive columns; the rows are dealt round-robin over the 4 warps of the block, each summing its rows in
// row order in fp64 (eight loads in flight), and the four partial sums are added in a fixed order: bitwise reproducible.
// (One thread per column walked 150 - 300 rows alone: 24 - 35 us, 3 % of a 120 000-point training iteration.)
__global__ void __launch_bounds__(128) nsf_finalize_kernel(NsfNetGeom g, const float* __restrict__ scratch, int rows, const int* __restrict__ map,
                                                           float* __restrict__ grad, float* __restrict__ loss_parts, const int* __restrict__ map0, int split) {
  constexpr int RG = 4;
  __shared__ double part[RG][32];
  const int c = threadIdx.x & 31, rgp = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + c;
  const int np = g.n_params;
  const bool is_loss = i >= np;
  const bool on = i < np + NSF_LOSS_SLOTS && (is_loss ? (loss_parts != nullptr) : (grad != nullptr));
  double acc = 0.0;
  if (on) {
    const int col1 = is_loss ? g.gs_loss() + (i - np) : map[i];
    const int col0 = (split > 0 && !is_loss) ? map0[i] : col1;
    const long long stride = g.gs_row();
    int r = rgp;
    for (; r + 7 * RG < rows; r += 8 * RG) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) { const int rr = r + k * RG; v[k] = __ldcg(scratch + (long long)rr * stride + (rr < split ? col0 : col1)); }
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += (double)v[k];
    }
    for (; r < rows; r += RG) acc += (double)__ldcg(scratch + (long long)r * stride + (r < split ? col0 : col1));
  }
  part[rgp][c] = acc;
  __syncthreads();
  if (rgp == 0 && on) {
    const float t = (float)((part[0][c] + part[1][c]) + (part[2][c] + part[3][c]));
    if (is_loss) loss_parts[i - np] = t;
    else grad[i] = t;
  }
}

int nsf_finalize_launch(const NsfNetGeom& g, const float* scratch, int rows, const int* map, float* grad,
                        float* loss_parts, nsf_stream_t st, const int* map0, int split) {
  const int n = g.n_params + NSF_LOSS_SLOTS;
  if (!map0 || split < 0) split = 0;
  if (split > rows) split = rows;
  nsf_finalize_kernel<<<(n + 31) / 32, 128, 0, st>>>(g, scratch, rows, map, grad, loss_parts, map0, split);
  NSF_CUDA_OK(cudaGetLastError());
  return NSF_OK;
}

// torch.optim.Adam (no amsgrad, wd = 0): m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
// p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void nsf_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                long long n, float lr, float b1, float b2, float eps, float bc1, float bc2, float gs) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i] * gs;
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi; v[i] = vi;
  const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
  p[i] -= (lr / bc1) * (mi / denom);
}

int nsf_adam_launch(float* params, const float* grad, float* m, float* v, long long n, float lr, float b1, float b2,
                    float eps, float bc1, float bc2, float grad_scale, nsf_stream_t st) {
  if (n <= 0) return NSF_OK;
  nsf_adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(params, grad, m, v, n, lr, b1, b2, eps, bc1, bc2, grad_scale);
  NSF_CUDA_OK(cudaGetLastError());
  return NSF_OK;
}

#else
// ---------------------------------------------------------------------------------------------
// host SIMT emulation (tests/emu only)
// ---------------------------------------------------------------------------------------------
#include <cmath>

int nsf_emu_reverse = 0;
extern "C" void nsf_emu_set_reverse(int v) { nsf_emu_reverse = v; }

template <int NS, int PT>
static void emu_run(const NsfKernelArgs& a, int grid, int nt, size_t smem_floats) {
  std::vector<float> smem(smem_floats);
  std::vector<NsfRegs<NS>> regs(nt);
  for (int bid = 0; bid < grid; ++bid) {
    std::fill(smem.begin(), smem.end(), NAN);  // catch reads of unwritten shared memory
    nsf_cta_program<NS, PT>(a, smem.data(), bid, grid, nt, regs.data());
  }
}

int nsf_ffma_occupancy(int ns, int hp) { return ns == 4 ? (nsf_ffma_pt(ns, hp) == 32 ? 2 : 3) : 4; }

int nsf_ffma_launch(NsfKernelArgs& a, int ns, int grid, nsf_stream_t) {
  const int hp = a.g.HP;
  const int pt = nsf_ffma_pt(ns, hp);
  const int nt = (hp / 4) * (pt / 4);
  const size_t sf = (size_t)nsf_ffma_smem_floats(ns, pt, hp);
  a.n_tiles = (int)((a.n + pt - 1) / pt);
  if (a.n_tiles <= 0 || grid <= 0) return NSF_OK;
  if (ns == 4) { if (pt == 32) emu_run<4, 32>(a, grid, nt, sf); else emu_run<4, 16>(a, grid, nt, sf); }
  else emu_run<1, 32>(a, grid, nt, sf);
  return NSF_OK;
}

int nsf_pack_launch(const NsfNetGeom& g, const float* flat, float* pk, nsf_stream_t) {
  for (int i = 0; i < g.pk_size(); ++i) pk[i] = nsf_pack_value(g, flat, i);
  return NSF_OK;
}

int nsf_finalize_launch(const NsfNetGeom& g, const float* scratch, int rows, const int* map, float* grad,
                        float* loss_parts, nsf_stream_t, const int*, int) {
  const long long stride = g.gs_row();
  if (grad)
    for (int i = 0; i < g.n_params; ++i) {
      double acc = 0;
      for (int r = 0; r < rows; ++r) acc += (double)scratch[r * stride + map[i]];
      grad[i] = (float)acc;
    }
  if (loss_parts)
    for (int s = 0; s < NSF_LOSS_SLOTS; ++s) {
      double acc = 0;
      for (int r = 0; r < rows; ++r) acc += (double)scratch[r * stride + g.gs_loss() + s];
      loss_parts[s] = (float)acc;
    }
  return NSF_OK;
}

int nsf_adam_launch(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2,
                    float eps, float bc1, float bc2, float gs, nsf_stream_t) {
  for (long long i = 0; i < n; ++i) {
    const float gi = g[i] * gs;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= (lr / bc1) * (mi / (std::sqrt(vi) / std::sqrt(bc2) + eps));
  }
  return NSF_OK;
}
#endif
