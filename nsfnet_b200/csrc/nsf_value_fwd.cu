// Value-only forward of a small FCNet with ONE output (the entropy-viscosity net, ev-NSFnet/pinn_solver.py:280-288:
// e = net_1(x, y), 2 -> 40 x 4 -> 1 in the shipped configuration), one thread per point.
//
// The FFMA tile kernel (nsf_ffma_body.h) spends a CTA barrier per layer and phase and reaches ~19 % of the FP32 pipe
// on this net (692 us per 10^6 points, round-1 launch list), 7 % of a training step.  Here a thread owns a point:
//   * all parameters sit in shared memory, hidden weights transposed (Wt[k][j]) so that one broadcast LDS.128 yields
//     W[j..j+3][k]: 1 shared-memory instruction per 4 FFMA, no bank conflicts (every lane reads the same address);
//   * the thread's activations of the previous layer are registers (k loop unrolled), the new ones go to a
//     per-thread column of shared memory (act[j][tid]: consecutive threads -> consecutive banks) because the output
//     neuron index j is a run-time loop variable;
//   * no barrier after the parameter load: threads only touch their own columns;
//   * a thread carries TWO points (i, i + 128), so every weight fetch is used twice and eight accumulation chains are in flight.
// Template on the hidden width (40: production.yaml / config.py; 20: the reference ctor default hidden_size_1).
#include "nsf_internal.h"
#include "nsf_math.cuh"

namespace {

constexpr int VT = 128;   // threads per CTA
constexpr int NP = 2;     // points per thread and pass

template <int H>
__global__ void __launch_bounds__(VT) nsf_value_fwd_kernel(int L, const float* __restrict__ flat, const float* __restrict__ x,
                                                           const float* __restrict__ y, long long n, float* __restrict__ out) {
  extern __shared__ __align__(16) float vsm[];
  // parameter block: w0x[H] w0y[H] b0[H] | (Wt_l[H k][H j], b_l[H]) l = 1..L-1 | wl[H] bl
  float* w0x = vsm; float* w0y = vsm + H; float* b0 = vsm + 2 * H;
  float* hid = vsm + 3 * H;
  constexpr int PER = H * H + H;
  float* wl = hid + (L - 1) * PER;
  const int psz = 3 * H + (L - 1) * PER + H + 4;
  float* act = vsm + ((psz + 3) & ~3);          // [H][NP * VT]
  const int tid = threadIdx.x;

  // flat (state_dict) order: W0[H][2], b0[H], (W_l[H j][H k], b_l[H]) l = 1..L-1, WL[1][H], bL[1]
  for (int i = tid; i < H; i += VT) { w0x[i] = flat[2 * i]; w0y[i] = flat[2 * i + 1]; b0[i] = flat[2 * H + i]; }
  for (int l = 1; l < L; ++l) {
    const float* src = flat + 3 * H + (l - 1) * PER;
    float* dst = hid + (l - 1) * PER;
    for (int i = tid; i < H * H; i += VT) { const int j = i / H, k = i % H; dst[k * H + j] = src[i]; }
    for (int i = tid; i < H; i += VT) dst[H * H + i] = src[H * H + i];
  }
  {
    const float* src = flat + 3 * H + (L - 1) * PER;
    for (int i = tid; i < H; i += VT) wl[i] = src[i];
    if (tid == 0) wl[H] = src[H];
  }
  __syncthreads();

  // two points per thread (i and i + VT): one broadcast LDS.128 of W[j..j+3][k] feeds 8 FFMA on 8 independent chains
  float* col = act + tid;
  constexpr int RS = NP * VT;                    // row stride of the activation columns
  for (long long base = (long long)blockIdx.x * RS; base < n; base += (long long)gridDim.x * RS) {
    float xv[NP], yv[NP];
    bool ok[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      const long long i = base + q * VT + tid;
      ok[q] = i < n;
      xv[q] = ok[q] ? __ldg(x + i) : 0.f;
      yv[q] = ok[q] ? __ldg(y + i) : 0.f;
    }
#pragma unroll 2
    for (int j = 0; j < H; j += 4) {
      const float4 wx = *reinterpret_cast<const float4*>(w0x + j), wy = *reinterpret_cast<const float4*>(w0y + j),
                   bb = *reinterpret_cast<const float4*>(b0 + j);
#pragma unroll
      for (int q = 0; q < NP; ++q) {
        col[(j + 0) * RS + q * VT] = nsf_tanh_fast(fmaf(wx.x, xv[q], fmaf(wy.x, yv[q], bb.x)));
        col[(j + 1) * RS + q * VT] = nsf_tanh_fast(fmaf(wx.y, xv[q], fmaf(wy.y, yv[q], bb.y)));
        col[(j + 2) * RS + q * VT] = nsf_tanh_fast(fmaf(wx.z, xv[q], fmaf(wy.z, yv[q], bb.z)));
        col[(j + 3) * RS + q * VT] = nsf_tanh_fast(fmaf(wx.w, xv[q], fmaf(wy.w, yv[q], bb.w)));
      }
    }
    for (int l = 1; l < L; ++l) {
      float a[NP][H];
#pragma unroll
      for (int k = 0; k < H; ++k)
#pragma unroll
        for (int q = 0; q < NP; ++q) a[q][k] = col[k * RS + q * VT];
      const float* Wt = hid + (l - 1) * PER;
#pragma unroll 1
      for (int j = 0; j < H; j += 4) {
        const float4 bz = *reinterpret_cast<const float4*>(Wt + H * H + j);
        float4 z[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) z[q] = bz;
#pragma unroll
        for (int k = 0; k < H; ++k) {
          const float4 w = *reinterpret_cast<const float4*>(Wt + k * H + j);
#pragma unroll
          for (int q = 0; q < NP; ++q) {
            z[q].x = fmaf(w.x, a[q][k], z[q].x); z[q].y = fmaf(w.y, a[q][k], z[q].y);
            z[q].z = fmaf(w.z, a[q][k], z[q].z); z[q].w = fmaf(w.w, a[q][k], z[q].w);
          }
        }
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          col[(j + 0) * RS + q * VT] = nsf_tanh_fast(z[q].x);
          col[(j + 1) * RS + q * VT] = nsf_tanh_fast(z[q].y);
          col[(j + 2) * RS + q * VT] = nsf_tanh_fast(z[q].z);
          col[(j + 3) * RS + q * VT] = nsf_tanh_fast(z[q].w);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      float o = wl[H];
#pragma unroll 8
      for (int k = 0; k < H; ++k) o = fmaf(wl[k], col[k * RS + q * VT], o);
      if (ok[q]) out[base + q * VT + tid] = o;
    }
  }
}

template <int H>
size_t value_fwd_smem(int L) {
  const int psz = 3 * H + (L - 1) * (H * H + H) + H + 4;
  return sizeof(float) * (size_t)(((psz + 3) & ~3) + H * VT * NP);
}

}  // namespace

int nsf_value_fwd_supported(const NsfNetGeom& g) {
  if (g.n_out != 1 || g.L < 1 || g.L > NSF_MAX_LAYERS) return 0;
  if (g.H == 40) return value_fwd_smem<40>(g.L) <= 200 * 1024;
  if (g.H == 20) return value_fwd_smem<20>(g.L) <= 200 * 1024;
  return 0;
}

// out[i] = net(x[i], y[i]) for the flat (state_dict order) parameters `flat`; asynchronous on `st`
int nsf_value_fwd_launch(const NsfNetGeom& g, int sms, const float* flat, const float* x, const float* y, long long n, float* out,
                         nsf_stream_t st) {
  if (n <= 0) return NSF_OK;
  const size_t smem = g.H == 40 ? value_fwd_smem<40>(g.L) : value_fwd_smem<20>(g.L);
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  long long blocks = (n + VT * NP - 1) / (VT * NP);
  const long long cap = (long long)sms * per_sm;
  if (blocks > cap) blocks = cap;
  if (g.H == 40) {
    NSF_CUDA_OK(cudaFuncSetAttribute(nsf_value_fwd_kernel<40>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nsf_value_fwd_kernel<40><<<(unsigned)blocks, VT, smem, st>>>(g.L, flat, x, y, n, out);
  } else {
    NSF_CUDA_OK(cudaFuncSetAttribute(nsf_value_fwd_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nsf_value_fwd_kernel<20><<<(unsigned)blocks, VT, smem, st>>>(g.L, flat, x, y, n, out);
  }
  NSF_CUDA_OK(cudaGetLastError());
  return NSF_OK;
}
