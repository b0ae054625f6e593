// Probe of the CTA-pair tensor-core path (tcgen05 .cta_group::2) on B200 -- not part of the library.
//   1. forward-like MMA: M = 128 over a CTA pair (64 rows = 16 points x 4 streams per CTA), A = the per-CTA activation
//      image (MN-major, descriptor layout type 1), B = the layer weights, K-major, N/2 rows per CTA.  Where do the
//      results land in each CTA's tensor memory?  (expected: lane = row + 64 * (n >= N/2), column = n mod N/2)
//   2. weight-gradient-like MMA with .cta_group::1 per CTA issued in the SAME kernel after the pair MMAs
//      (both operands = the type-1 images read K-major, contraction over the 64 rows)
//   3. cross-CTA hand-over: remote mbarrier arrive (peer -> leader), multicast commit (leader -> both)
//   4. cycles per MMA of both kinds
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I nsfnet_b200/csrc scripts/probe_2sm.cu -o scripts/_bin/probe_2sm
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "nsf_tc.cuh"
using namespace nsftc;

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma2_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit2(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
               "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ uint64_t desc_full(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t ltype) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(ltype & 7) << 61;
  return d;
}

// type-1 image of a [64 rows m][KP neurons k] tile: group g = m / 32, then atoms of 4 neurons x 128 B
template <int KP> __host__ __device__ constexpr uint32_t img_group() { return (KP / 4) * 512; }
template <int KP> __host__ __device__ inline uint32_t img_off(int m, int k) {
  const int g = m >> 5, ml = m & 31;
  return (uint32_t)g * img_group<KP>() + (uint32_t)(k >> 2) * 512u + (uint32_t)(k & 3) * 128u + (uint32_t)(((ml >> 3) ^ (k & 3)) * 32) + (uint32_t)(ml & 7) * 4u;
}
// K-major no-swizzle weight image: rows n (local), 16-byte K chunks 128 B apart, 8-row bands (KP/4)*128 apart
template <int KP> __host__ __device__ inline uint32_t w_off(int n, int k) {
  return (uint32_t)(n >> 3) * (uint32_t)((KP / 4) * 128) + (uint32_t)(k >> 2) * 128u + (uint32_t)(n & 7) * 16u + (uint32_t)(k & 3) * 4u;
}

template <int KP, int NP>   // NP = padded N of the pair MMA (multiple of 16)
struct Lay {
  static constexpr uint32_t IMG = 2 * img_group<KP>();            // one 64-row image
  static constexpr uint32_t IMG_PAD = ((128 / 4) * 512 > IMG ? (128 / 4) * 512 : IMG);   // the M = 128 K-major read of the wgrad touches 32 atoms
  static constexpr uint32_t WIMG = (NP / 2 / 8) * (KP / 4) * 128;
  static constexpr uint32_t OFF_P = 0, OFF_Q = OFF_P + 2 * IMG_PAD, OFF_W = OFF_Q + 2 * IMG_PAD;
  static constexpr uint32_t BYTES = OFF_W + ((WIMG + 1023) / 1024) * 1024 + 1024;
};

template <int KP, int NP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) probe2(const float* __restrict__ A /* [2][64][KP] */, const float* __restrict__ Q /* [2][64][KP] */,
                                                                       const float* __restrict__ W /* [NP][KP] */, float* __restrict__ out1 /* [2][128][NP/2] */,
                                                                       float* __restrict__ out2 /* [2][128][128] */, long long* __restrict__ tim /* [2][8] */, int mix) {
  using L = Lay<KP, NP>;
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t ready, done, done1;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  if (warp == 0) tmem_alloc2(&tmem_base, 512);
  if (tid == 0) { mbar_init(&ready, 2); mbar_init(&done, 1); mbar_init(&done1, 1); mbar_fence_init(); }
  for (uint32_t i = tid; i < L::BYTES / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
  __syncthreads();
  for (int i = tid; i < 64 * KP; i += 128) {
    const int m = i / KP, k = i % KP;
    *reinterpret_cast<float*>(smem + L::OFF_P + img_off<KP>(m, k)) = A[(size_t)rank * 64 * KP + i];
    *reinterpret_cast<float*>(smem + L::OFF_Q + img_off<KP>(m, k)) = Q[(size_t)rank * 64 * KP + i];
  }
  for (int i = tid; i < (NP / 2) * KP; i += 128) {
    const int n = i / KP, k = i % KP;
    *reinterpret_cast<float*>(smem + L::OFF_W + w_off<KP>(n, k)) = W[(size_t)(rank * (NP / 2) + n) * KP + k];
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // both CTAs' barriers initialised and tensor memory allocated before any remote arrive / pair MMA
  tc_fence_after();
  const uint32_t tb = tmem_base;
  const uint32_t sb = smem_u32(smem);
  // hand-over under test: each CTA's thread 0 arrives on the LEADER's `ready`
  if (tid == 0) mbar_arrive_remote(&ready, 0);
  long long t_issue = 0, t_done = 0;
  if (rank == 0 && tid == 0) {
    mbar_wait_cluster(&ready, 0);
    tc_fence_after();
    const uint32_t idesc = idesc_tf32(128, NP, 1, 0);
    const long long t0 = clock64();
    for (int ks = 0; ks < KP / 8; ++ks) {
      const uint64_t da = desc_full(sb + L::OFF_P + ks * 1024, img_group<KP>(), 512, 1);
      const uint64_t db = desc_full(sb + L::OFF_W + ks * 256, 128, (KP / 4) * 128, 0);
      mma2_tf32(tb, da, db, idesc, ks > 0);
    }
    mma_commit2(&done, 3);
    t_issue = clock64() - t0;
  }
  mbar_wait(&done, 0);
  tc_fence_after();
  // dump D: 128 lanes x NP/2 columns
  for (int c0 = 0; c0 < NP / 2; c0 += 8) {
    float v[8];
    tmem_ld8(tb + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    tmem_ld_wait();
    for (int i = 0; i < 8; ++i) out1[((size_t)rank * 128 + tid) * (NP / 2) + c0 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (mix) {
    // per-CTA weight-gradient-like MMA, cta_group::1: D2[j, k'] = sum_r P[r, j] Q[r, k'], M = 128 (rows j), N = 128, K = 64 rows r
    if (tid == 0) {
      const uint32_t idesc = idesc_tf32(128, 128, 0, 0);
      for (int ks = 0; ks < 8; ++ks) {
        const uint32_t o = (ks >> 2) * img_group<KP>() + (ks & 3) * 32;
        const uint64_t da = desc_full(sb + L::OFF_P + o, 0, 512, 1);
        const uint64_t db = desc_full(sb + L::OFF_Q + o, 0, 512, 1);
        mma_tf32(tb + 256, da, db, idesc, ks > 0);
      }
      mma_commit(&done1);
    }
    mbar_wait(&done1, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < 128; c0 += 8) {
      float v[8];
      tmem_ld8(tb + 256 + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      for (int i = 0; i < 8; ++i) out2[((size_t)rank * 128 + tid) * 128 + c0 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  // ---- timing: 3 * KP/8 pair MMAs back to back (one 3xTF32 stage), 8 repetitions; then 24 single-CTA wgrad MMAs ----
  for (int rep = 0; rep < 8; ++rep) {
    if (rank == 0 && tid == 0) {
      const uint32_t idesc = idesc_tf32(128, NP, 1, 0);
      const long long t0 = clock64();
      for (int i = 0; i < 3 * (KP / 8); ++i) {
        const int ks = i % (KP / 8);
        const uint64_t da = desc_full(sb + L::OFF_P + ks * 1024, img_group<KP>(), 512, 1);
        const uint64_t db = desc_full(sb + L::OFF_W + ks * 256, 128, (KP / 4) * 128, 0);
        mma2_tf32(tb, da, db, idesc, i > 0);
      }
      mma_commit2(&done, 3);
      t_issue = clock64() - t0;
      mbar_wait(&done, (rep + 1) & 1);
      t_done = clock64() - t0;
      if (rep >= 2) { tim[rep] = t_done; tim[8 + rep] = t_issue; }
    } else {
      mbar_wait(&done, (rep + 1) & 1);
    }
    tc_fence_after();
  }
  if (mix) {
    for (int rep = 0; rep < 8; ++rep) {
      if (tid == 0) {
        const uint32_t idesc = idesc_tf32(128, NP == 80 ? 80 : 128, 0, 0);
        const long long t0 = clock64();
        for (int i = 0; i < 24; ++i) {
          const int ks = i % 8;
          const uint32_t o = (ks >> 2) * img_group<KP>() + (ks & 3) * 32;
          mma_tf32(tb + 256, desc_full(sb + L::OFF_P + o, 0, 512, 1), desc_full(sb + L::OFF_Q + o, 0, 512, 1), idesc, i > 0);
        }
        mma_commit(&done1);
        mbar_wait(&done1, (rep + 1) & 1);
        if (rep >= 2 && rank == 0) tim[16 + rep] = clock64() - t0;
      }
      __syncthreads();
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2(tb, 512);
}

template <int KP, int NP>
static int run(int mix) {
  using L = Lay<KP, NP>;
  printf("==== KP=%d NP=%d mix=%d smem=%u\n", KP, NP, mix, L::BYTES);
  std::vector<float> A(2 * 64 * KP), Q(2 * 64 * KP), W((size_t)NP * KP, 0.f);
  srand(1234 + KP);
  for (auto& v : A) v = (float)(rand() % 9 - 4);
  for (auto& v : Q) v = (float)(rand() % 7 - 3);
  for (int n = 0; n < KP; ++n) for (int k = 0; k < KP; ++k) W[(size_t)n * KP + k] = (float)(rand() % 5 - 2);
  float *dA, *dQ, *dW, *o1, *o2; long long* dt;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dQ, Q.size() * 4)); CK(cudaMalloc(&dW, W.size() * 4));
  CK(cudaMalloc(&o1, 2 * 128 * (NP / 2) * 4)); CK(cudaMalloc(&o2, 2 * 128 * 128 * 4)); CK(cudaMalloc(&dt, 64 * 8));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dQ, Q.data(), Q.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(o1, 0xff, 2 * 128 * (NP / 2) * 4)); CK(cudaMemset(o2, 0xff, 2 * 128 * 128 * 4)); CK(cudaMemset(dt, 0, 64 * 8));
  CK(cudaFuncSetAttribute(probe2<KP, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::BYTES));
  probe2<KP, NP><<<2, 128, L::BYTES>>>(dA, dQ, dW, o1, o2, dt, mix);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 2; }
  std::vector<float> h1(2 * 128 * (NP / 2)), h2(2 * 128 * 128);
  long long t[64];
  CK(cudaMemcpy(h1.data(), o1, h1.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h2.data(), o2, h2.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(t, dt, sizeof(t), cudaMemcpyDeviceToHost));
  // check 1: expected layout lane = m + 64 * (n >= NP/2), col = n % (NP/2)
  long bad = 0, tot = 0;
  for (int c = 0; c < 2; ++c)
    for (int m = 0; m < 64; ++m)
      for (int n = 0; n < KP; ++n) {
        double ref = 0;
        for (int k = 0; k < KP; ++k) ref += (double)A[((size_t)c * 64 + m) * KP + k] * W[(size_t)n * KP + k];
        const int lane = m + 64 * (n >= NP / 2), col = n % (NP / 2);
        const float got = h1[((size_t)c * 128 + lane) * (NP / 2) + col];
        ++tot;
        if (got != (float)ref) { if (bad < 6) printf("  pair MMA mismatch cta %d m %d n %d: got %g want %g\n", c, m, n, got, ref); ++bad; }
      }
  printf("pair MMA (cta_group::2, M=128, N=%d): %ld / %ld mismatches\n", NP, bad, tot);
  if (bad) {   // diagnostic: where did D[cta 0][m = 1][n = 2] and [m = 40][n = NP/2 + 3] go?
    for (int c = 0; c < 2; ++c) {
      printf("  cta %d lane 0 cols 0..7:", c);
      for (int i = 0; i < 8; ++i) printf(" %g", h1[((size_t)c * 128) * (NP / 2) + i]);
      printf("\n  cta %d lane 64 cols 0..7:", c);
      for (int i = 0; i < 8; ++i) printf(" %g", h1[((size_t)c * 128 + 64) * (NP / 2) + i]);
      printf("\n");
    }
  }
  if (mix) {
    long bad2 = 0, tot2 = 0;
    for (int c = 0; c < 2; ++c)
      for (int j = 0; j < KP; ++j)
        for (int k = 0; k < KP; ++k) {
          double ref = 0;
          for (int r = 0; r < 64; ++r) ref += (double)A[((size_t)c * 64 + r) * KP + j] * Q[((size_t)c * 64 + r) * KP + k];
          const float got = h2[((size_t)c * 128 + j) * 128 + k];
          ++tot2;
          if (got != (float)ref) { if (bad2 < 6) printf("  wgrad MMA mismatch cta %d j %d k %d: got %g want %g\n", c, j, k, got, ref); ++bad2; }
        }
    printf("wgrad MMA (cta_group::1 after cta_group::2, K-major type-1 images): %ld / %ld mismatches\n", bad2, tot2);
  }
  long long bd = 1LL << 60, bi = 1LL << 60, bw = 1LL << 60;
  for (int r = 2; r < 8; ++r) { if (t[r] < bd) bd = t[r]; if (t[8 + r] < bi) bi = t[8 + r]; if (t[16 + r] && t[16 + r] < bw) bw = t[16 + r]; }
  const int nm = 3 * (KP / 8);
  printf("pair MMA timing: %d MMAs issue %lld complete %lld cycles (%.1f / MMA)\n", nm, bi, bd, (double)bd / nm);
  if (mix) printf("wgrad timing: 24 MMAs complete %lld cycles (%.1f / MMA)\n", bw, (double)bw / 24);
  cudaFree(dA); cudaFree(dQ); cudaFree(dW); cudaFree(o1); cudaFree(o2); cudaFree(dt);
  return 0;
}

int main(int argc, char** argv) {
  CK(cudaSetDevice(0));
  const int which = argc > 1 ? atoi(argv[1]) : 0;
  const int mix = argc > 2 ? atoi(argv[2]) : 1;
  if (which == 0) return run<80, 80>(mix);
  if (which == 1) return run<120, 128>(mix);
  return 0;
}
