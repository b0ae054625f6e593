// Probe for the "points on M" jet kernel (not part of the library):
//   1. forward-like MMA in ONE CTA: A = activation image of a tile (rows m = 4*point + stream, MN-major, descriptor layout
//      type 1), B = one layer's weights in k-step sub-blocks (K-major, no swizzle), M = 128 (32 points) or M = 64 (16 points,
//      two tiles sharing the accumulator columns in the two lane halves of every TMEM quadrant); result layout + cycles
//   2. the 16-lane TMEM load (tcgen05.ld.16x32bx2) the M = 64 epilogue needs
//   3. the 4x4 quad transpose (streams <-> neurons) of the epilogue, by shuffles
//   4. L2 -> shared-memory streaming rate of TMA bulk copies through a ring, all SMs at once (weight streaming for hidden = 120)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I nsfnet_b200/csrc scripts/probe_pm.cu -o scripts/_bin/probe_pm
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "nsf_tc.cuh"
using namespace nsftc;

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t desc_full(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t ltype) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(ltype & 7) << 61;
  return d;
}
template <int KP> __host__ __device__ constexpr uint32_t img_group() { return (KP / 4) * 512; }
template <int KP> __host__ __device__ inline uint32_t img_off(int m, int k) {
  const int g = m >> 5, ml = m & 31;
  return (uint32_t)g * img_group<KP>() + (uint32_t)(k >> 2) * 512u + (uint32_t)(k & 3) * 128u + (uint32_t)(((ml >> 3) ^ (k & 3)) * 32) + (uint32_t)(ml & 7) * 4u;
}
// weights: k-step sub-blocks of NP*32 bytes: [NP/8 bands of 256 B][2 chunks of 128 B][8 rows x 16 B]
template <int NP> __host__ __device__ inline uint32_t w_off(int n, int k) {
  return (uint32_t)(k >> 3) * (uint32_t)(NP * 32) + (uint32_t)(n >> 3) * 256u + (uint32_t)((k >> 2) & 1) * 128u + (uint32_t)(n & 7) * 16u + (uint32_t)(k & 3) * 4u;
}
__device__ __forceinline__ void tmem_ld_16x32bx2_x4(uint32_t taddr, float* v) {
  uint32_t r0, r1, r2, r3;
  asm volatile("tcgen05.ld.sync.aligned.16x32bx2.x4.b32 {%0,%1,%2,%3}, [%4], 4;" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
  v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}

template <int KP, int NP, int MT>
__global__ void __launch_bounds__(128) probe_fwd(const float* __restrict__ A /* [2][MT][KP] */, const float* __restrict__ W /* [NP][KP] */,
                                                 float* __restrict__ out /* [2][128][NP] */, long long* __restrict__ tim) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_base;
  constexpr uint32_t IMG = (MT / 32) * img_group<KP>();
  constexpr uint32_t OFF_P0 = 0, OFF_P1 = IMG, OFF_W = 2 * IMG;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  if (tid == 0) { mbar_init(&done, 1); mbar_fence_init(); }
  for (int i = tid; i < MT * KP; i += 128) {
    const int m = i / KP, k = i % KP;
    *reinterpret_cast<float*>(smem + OFF_P0 + img_off<KP>(m, k)) = A[i];
    *reinterpret_cast<float*>(smem + OFF_P1 + img_off<KP>(m, k)) = A[MT * KP + i];
  }
  for (int i = tid; i < NP * KP; i += 128) {
    const int n = i / KP, k = i % KP;
    *reinterpret_cast<float*>(smem + OFF_W + w_off<NP>(n, k)) = W[i];
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base, sb = smem_u32(smem);
  const uint32_t idesc = idesc_tf32(MT, NP, 1, 0);
  constexpr int NT = MT == 64 ? 2 : 1;
  if (tid == 0) {
    for (int t = 0; t < NT; ++t)
      for (int ks = 0; ks < KP / 8; ++ks)
        mma_tf32(tb + ((uint32_t)(16 * t) << 16), desc_full(sb + (t ? OFF_P1 : OFF_P0) + ks * 1024, img_group<KP>(), 512, 1),
                 desc_full(sb + OFF_W + ks * NP * 32, 128, 256, 0), idesc, ks > 0);
    mma_commit(&done);
  }
  mbar_wait(&done, 0);
  tc_fence_after();
  if (MT == 128) {
    for (int c0 = 0; c0 < NP; c0 += 8) {
      float v[8];
      tmem_ld8(tb + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      for (int i = 0; i < 8; ++i) out[(size_t)tid * NP + c0 + i] = v[i];      // lane tid = row m
    }
  } else {
    // warp w, tile t: lanes 32w + 16t + (0..15) = rows 16w + (0..15); half-warps take columns c0..c0+3 / c0+4..c0+7
    for (int t = 0; t < 2; ++t)
      for (int c0 = 0; c0 < NP; c0 += 8) {
        float v[4];
        tmem_ld_16x32bx2_x4(tb + ((uint32_t)(warp * 32 + 16 * t) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
        const int m = 16 * warp + (lane & 15), c = c0 + 4 * (lane >> 4);
        for (int i = 0; i < 4; ++i) out[((size_t)t * 128 + m) * NP + c + i] = v[i];
      }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    for (int rep = 0; rep < 8; ++rep) {
      const long long t0 = clock64();
#pragma unroll 1
      for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int ks = 0; ks < KP / 8; ++ks)
          mma_tf32(tb, desc_full(sb + OFF_P0 + ks * 1024, img_group<KP>(), 512, 1), desc_full(sb + OFF_W + ks * NP * 32, 128, 256, 0), idesc, (p | ks) > 0);
      mma_commit(&done);
      mbar_wait(&done, (rep + 1) & 1);
      tim[rep] = clock64() - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

template <int KP, int NP, int MT>
static void run_fwd() {
  constexpr uint32_t IMG = (MT / 32) * img_group<KP>();
  const size_t smem = 2 * IMG + (size_t)(KP / 8) * NP * 32 + 1024;
  std::vector<float> A(2 * MT * KP), W((size_t)NP * KP);
  srand(99 + KP + MT);
  for (auto& v : A) v = (float)(rand() % 9 - 4);
  for (auto& v : W) v = (float)(rand() % 5 - 2);
  float *dA, *dW, *o; long long* dt;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dW, W.size() * 4)); CK(cudaMalloc(&o, 2 * 128 * NP * 4)); CK(cudaMalloc(&dt, 64));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(o, 0xff, 2 * 128 * NP * 4));
  CK(cudaFuncSetAttribute(probe_fwd<KP, NP, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_fwd<KP, NP, MT><<<1, 128, smem>>>(dA, dW, o, dt);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("[fwd KP=%d NP=%d MT=%d] CUDA error: %s\n", KP, NP, MT, cudaGetErrorString(e)); exit(2); }
  std::vector<float> h(2 * 128 * NP); long long t[8];
  CK(cudaMemcpy(h.data(), o, h.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(t, dt, 64, cudaMemcpyDeviceToHost));
  long bad = 0, tot = 0;
  for (int tl = 0; tl < (MT == 64 ? 2 : 1); ++tl)
    for (int m = 0; m < MT; ++m)
      for (int n = 0; n < NP; ++n) {
        double ref = 0;
        for (int k = 0; k < KP; ++k) ref += (double)A[((size_t)tl * MT + m) * KP + k] * W[(size_t)n * KP + k];
        const float got = h[((size_t)tl * 128 + m) * NP + n];
        ++tot;
        if (got != (float)ref) { if (bad < 5) printf("  mismatch tile %d m %d n %d: got %g want %g\n", tl, m, n, got, ref); ++bad; }
      }
  long long best = 1LL << 60;
  for (int r = 2; r < 8; ++r) if (t[r] < best) best = t[r];
  printf("fwd MMA KP=%3d N=%3d M=%3d: %ld / %ld mismatches; %d MMAs in %lld cycles (%.1f / MMA)\n", KP, NP, MT, bad, tot, 3 * (KP / 8), best,
         (double)best / (3 * (KP / 8)));
  cudaFree(dA); cudaFree(dW); cudaFree(o); cudaFree(dt);
}

// ---- quad transpose -------------------------------------------------------------------------------------------------
// lane s of a quad holds v[i] = value of (stream s, neuron i); afterwards lane i holds w[s] = value of (stream s, neuron i)
__device__ __forceinline__ void quad_transpose(const float v[4], float w[4], int lane) {
  const bool b0 = lane & 1, b1 = lane & 2;
  const float sa = b0 ? v[0] : v[1], sb_ = b0 ? v[2] : v[3];
  const float ra = __shfl_xor_sync(0xffffffffu, sa, 1), rb = __shfl_xor_sync(0xffffffffu, sb_, 1);
  const float p0 = b0 ? ra : v[0], p1 = b0 ? v[1] : ra;     // neuron b0: streams (2*b1', 2*b1'+1) of this lane pair
  const float q0 = b0 ? rb : v[2], q1 = b0 ? v[3] : rb;     // neuron 2 + b0
  const float s0 = b1 ? p0 : q0, s1 = b1 ? p1 : q1;
  const float r0 = __shfl_xor_sync(0xffffffffu, s0, 2), r1 = __shfl_xor_sync(0xffffffffu, s1, 2);
  const float k0 = b1 ? q0 : p0, k1 = b1 ? q1 : p1;
  w[0] = b1 ? r0 : k0; w[1] = b1 ? r1 : k1; w[2] = b1 ? k0 : r0; w[3] = b1 ? k1 : r1;
}
__global__ void probe_transpose(int* bad) {
  const int lane = threadIdx.x & 31;
  float v[4], w[4];
  for (int i = 0; i < 4; ++i) v[i] = (float)(100 * lane + i);
  quad_transpose(v, w, lane);
  for (int s = 0; s < 4; ++s)
    if (w[s] != (float)(100 * ((lane & ~3) + s) + (lane & 3))) atomicAdd(bad, 1);
}

// ---- TMA bulk streaming through a ring ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__global__ void __launch_bounds__(64) probe_stream(const unsigned char* __restrict__ src, size_t src_bytes, uint32_t chunk, int ring, int n_chunks,
                                                  long long* __restrict__ cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < ring; ++i) mbar_init(&full[i], 1);
    mbar_fence_init();
    const size_t nsrc = src_bytes / chunk;
    size_t pos = (size_t)blockIdx.x * 7 % nsrc;
    const long long t0 = clock64();
    for (int i = 0; i < ring && i < n_chunks; ++i) {
      mbar_expect_tx(&full[i], chunk);
      tma_bulk_g2s(smem + (size_t)i * chunk, src + pos * chunk, chunk, &full[i]);
      pos = (pos + 1) % nsrc;
    }
    for (int i = 0; i < n_chunks; ++i) {
      const int s = i % ring;
      mbar_wait(&full[s], (uint32_t)((i / ring) & 1));
      if (i + ring < n_chunks) {
        mbar_expect_tx(&full[s], chunk);
        tma_bulk_g2s(smem + (size_t)s * chunk, src + pos * chunk, chunk, &full[s]);
        pos = (pos + 1) % nsrc;
      }
    }
    cyc[blockIdx.x] = clock64() - t0;
  }
}

int main(int argc, char** argv) {
  CK(cudaSetDevice(0));
  run_fwd<80, 80, 128>();
  run_fwd<80, 16, 128>();
  run_fwd<80, 80, 64>();
  run_fwd<120, 120, 64>();
  run_fwd<120, 128, 128>();
  {
    int* bad; CK(cudaMalloc(&bad, 4)); CK(cudaMemset(bad, 0, 4));
    probe_transpose<<<1, 64>>>(bad);
    CK(cudaDeviceSynchronize());
    int h; CK(cudaMemcpy(&h, bad, 4, cudaMemcpyDeviceToHost));
    printf("quad transpose: %d mismatches\n", h);
  }
  {
    const size_t src_bytes = 1 << 20;
    unsigned char* src; long long* cyc;
    CK(cudaMalloc(&src, src_bytes)); CK(cudaMemset(src, 1, src_bytes)); CK(cudaMalloc(&cyc, 148 * 8));
    CK(cudaFuncSetAttribute(probe_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const uint32_t chunks[] = {5120, 7680, 12800, 19200, 25600};
    for (uint32_t ch : chunks)
      for (int ring : {2, 4, 8}) {
        if ((size_t)ch * ring > 190 * 1024) continue;
        const int n_chunks = 4000;
        probe_stream<<<148, 64, (size_t)ch * ring>>>(src, src_bytes, ch, ring, 64, cyc);      // warm-up
        CK(cudaEventRecord(e0));
        probe_stream<<<148, 64, (size_t)ch * ring>>>(src, src_bytes, ch, ring, n_chunks, cyc);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        long long h[148]; CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
        double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
        printf("TMA ring stream: chunk %5u B, ring %d: %.1f B/clk/SM (%.2f TB/s aggregate, %.0f cycles per chunk)\n", ch, ring,
               (double)ch * n_chunks / avg, 148.0 * ch * n_chunks / (ms * 1e-3) / 1e12, avg / n_chunks);
      }
  }
  printf("done\n");
  return 0;
}
