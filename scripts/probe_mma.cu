// Exploratory tcgen05 probe (not part of the library):
//   1. address decode: which shared-memory word does the tensor core read for B element (n, k) under a given
//      descriptor (major-ness, layout type, LBO, SBO)?  A is a K-major selector, the B region holds its own word index.
//   2. D lane layout for M = 64 / 128.
//   3. cycles of a batch of back-to-back MMAs (SS and A-from-TMEM) for several N.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I nsfnet_b200/csrc scripts/probe_mma.cu -o scripts/_bin/probe_mma
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
__device__ __forceinline__ uint32_t idesc_gen(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a_mn & 1) << 15) | ((uint32_t)(b_mn & 1) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct DecodeCfg { int which;  /* 0: probe B, 1: probe A */ int mn; int ltype; uint32_t lbo, sbo; int N; int M; uint32_t start; };

// Region under test: 16 KB at smem + 32768 (1024-aligned).  Selector operand at smem + 0 (K-major no swizzle, 128 rows x 8).
__global__ void __launch_bounds__(128) decode_kernel(DecodeCfg c, float* out /* [2][128][256] */) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base, 256);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  __syncthreads();
  for (int pass = 0; pass < 2; ++pass) {
    // selector: rows r (0..255 space), K = 8: sel[r][k] = (r == k); K-major no swizzle: SBO = 256 (8-row group: 2 chunks x 128 B), LBO = 128
    for (int i = tid; i < 256 * 8; i += 128) {
      const int r = i >> 3, k = i & 7;
      const uint32_t off = (uint32_t)(r >> 3) * 256u + (uint32_t)(k >> 2) * 128u + (uint32_t)(r & 7) * 16u + (uint32_t)(k & 3) * 4u;
      *reinterpret_cast<float*>(smem + off) = (r == k) ? 1.f : 0.f;
    }
    for (int i = tid; i < 8192; i += 128) {   // 32 KB region
      *reinterpret_cast<float*>(smem + 32768 + i * 4) = pass == 0 ? (float)(i & 2047) : (float)(i >> 11);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_base;
    if (tid == 0) {
      const uint64_t sel = desc_full(smem_u32(smem), 128, 256, 0);
      const uint64_t tst = desc_full(smem_u32(smem) + 32768 + c.start, c.lbo, c.sbo, c.ltype);
      if (c.which == 0) mma_tf32(tb, sel, tst, idesc_gen(c.M, c.N, 0, c.mn), 0);    // D[m, n] = sum_k sel[m,k] B[n,k] = B[n, m] (m < 8)
      else mma_tf32(tb, tst, sel, idesc_gen(c.M, c.N, c.mn, 0), 0);                 // D[m, n] = A[m, n] (n < 8)
      mma_commit(&bar);
    }
    mbar_wait(&bar, pass & 1);
    tc_fence_after();
    for (int c0 = 0; c0 < 256; c0 += 8) {
      float v[8];
      tmem_ld8(tb + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      for (int i = 0; i < 8; ++i) out[((size_t)pass * 128 + tid) * 256 + c0 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

struct TimeCfg { int M, N, nmma, ts; uint32_t a_lbo, a_sbo, a_step, b_lbo, b_sbo, b_step; int ksteps; };

__global__ void __launch_bounds__(128) time_kernel(TimeCfg c, long long* out /* [8] */) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  for (int i = tid; i < 50000; i += 128) *reinterpret_cast<float*>(smem + i * 4) = 0.001f * (float)(i & 63);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base;
  if (warp == 0) {
    const uint32_t leader = elect_one();
    const uint32_t idesc = idesc_gen(c.M, c.N, 0, 0);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 102400;
    for (int rep = 0; rep < 8; ++rep) {
      __syncwarp();
      const long long t0 = clock64();
      for (int i = 0; i < c.nmma; ++i) {
        const int ks = i % c.ksteps;
        if (c.ts) mma_tf32_ts_elect(tb, tb + 256 + (uint32_t)(ks * 8), desc_lo(b0 + ks * c.b_step, c.b_lbo), desc_hi(c.b_sbo), idesc, i > 0, leader);
        else mma_tf32_elect2(tb, desc_lo(a0 + ks * c.a_step, c.a_lbo), desc_hi(c.a_sbo), desc_lo(b0 + ks * c.b_step, c.b_lbo), desc_hi(c.b_sbo), idesc, i > 0, leader);
      }
      mma_commit_elect(&bar, leader);
      const long long t1 = clock64();
      mbar_wait(&bar, rep & 1);
      const long long t2 = clock64();
      if (tid == 0) { out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}


template <int M, int N, int TS, int NM, int SAMED>
__global__ void __launch_bounds__(128) time_kernel_t(long long* out /* [16] */) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  for (int i = tid; i < 50000; i += 128) *reinterpret_cast<float*>(smem + i * 4) = 0.001f * (float)(i & 63);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base;
  if (warp == 0) {
    const uint32_t leader = elect_one();
    const uint32_t idesc = idesc_gen(M, N, 0, 0);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 102400;
    constexpr uint32_t AHI = desc_hi(2560), BHI = desc_hi(2880);
    const uint32_t ah0 = desc_lo(a0, 128), al0 = desc_lo(a0 + 25600, 128);
    const uint32_t bh0 = desc_lo(b0, 144), bl0 = desc_lo(b0 + 46080, 144);
    for (int rep = 0; rep < 8; ++rep) {
      __syncwarp();
      const long long t0 = clock64();
#pragma unroll
      for (int i = 0; i < NM; ++i) {
        const int ks = i % 10;
        const uint32_t da = ks * (256 >> 4), db = ks * (288 >> 4);
        const uint32_t d = SAMED ? tb : tb + (uint32_t)((i & 1) * 256);
        if (TS) mma_tf32_ts_elect(d, tb + 448 + (uint32_t)((i % 8) * 8), ((i / 10) & 1 ? bl0 : bh0) + db, BHI, idesc, i > 1, leader);
        else mma_tf32_elect2(d, ((i / 10) & 1 ? al0 : ah0) + da, AHI, ((i / 10) & 1 ? bl0 : bh0) + db, BHI, idesc, i > 1, leader);
      }
      mma_commit_elect(&bar, leader);
      const long long t1 = clock64();
      mbar_wait(&bar, rep & 1);
      const long long t2 = clock64();
      if (tid == 0) { out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// B operand MN-major, descriptor layout type 1 (the jet kernel's R image: 128-byte rows of 32 columns n, 512-byte atoms
// of 4 k-rows, the next 32 columns LBO bytes on): does a wider N amortise the fetch of the A operand here as well?
template <int M, int N, int NM>
__global__ void __launch_bounds__(128) time_kernel_mn(long long* out, uint32_t lbo_bytes) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  for (int i = tid; i < 55000; i += 128) *reinterpret_cast<float*>(smem + i * 4) = 0.001f * (float)(i & 63);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base;
  if (warp == 0) {
    const uint32_t leader = elect_one();
    const uint32_t idesc = idesc_gen(M, N, 0, 1);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 61440;     // A: 2 x 25600 (hi | lo); B: images of 10240 B, 40960 B apart
    constexpr uint32_t AHI = desc_hi(2560);
    const uint32_t BHI = ((512u >> 4) & 0x3FFF) | (1u << 14) | (1u << 29);
    const uint32_t ah0 = desc_lo(a0, 128), al0 = desc_lo(a0 + 25600, 128);
    const uint32_t bh0 = desc_lo(b0, lbo_bytes), bl0 = desc_lo(b0 + 10240, lbo_bytes);
    for (int rep = 0; rep < 8; ++rep) {
      __syncwarp();
      const long long t0 = clock64();
#pragma unroll
      for (int i = 0; i < NM; ++i) {
        const int ks = i % 10;
        const uint32_t da = ks * (256 >> 4), db = ks * (1024 >> 4);
        mma_tf32_elect2(tb, ((i / 10) & 1 ? al0 : ah0) + da, AHI, ((i / 10) & 1 ? bl0 : bh0) + db, BHI, idesc, i > 1, leader);
      }
      mma_commit_elect(&bar, leader);
      const long long t1 = clock64();
      mbar_wait(&bar, rep & 1);
      const long long t2 = clock64();
      if (tid == 0) { out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}
template <int M, int N, int NM>
static void run_time_mn(uint32_t lbo, long long* d_t) {
  CK(cudaFuncSetAttribute(time_kernel_mn<M, N, NM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  time_kernel_mn<M, N, NM><<<1, 128, 220 * 1024>>>(d_t, lbo);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("[mn] CUDA error: %s\n", cudaGetErrorString(e)); exit(2); }
  long long t[16]; CK(cudaMemcpy(t, d_t, sizeof(t), cudaMemcpyDeviceToHost));
  long long bi = 1LL << 60, bt = 1LL << 60;
  for (int r = 2; r < 8; ++r) { if (t[2 * r] < bi) bi = t[2 * r]; if (t[2 * r + 1] < bt) bt = t[2 * r + 1]; }
  printf("SS, B MN-major type 1 (LBO %5u) M=%3d N=%3d nmma=%3d : issue %6lld cyc, complete %6lld cyc (%.1f / MMA)\n", lbo, M, N, NM, bi, bt, (double)bt / NM);
}

template <int M, int N, int TS, int NM, int SAMED>
static void run_time_t(const char* name, long long* d_t) {
  CK(cudaFuncSetAttribute(time_kernel_t<M, N, TS, NM, SAMED>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  time_kernel_t<M, N, TS, NM, SAMED><<<1, 128, 220 * 1024>>>(d_t);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("[%s] CUDA error: %s\n", name, cudaGetErrorString(e)); exit(2); }
  long long t[16]; CK(cudaMemcpy(t, d_t, sizeof(t), cudaMemcpyDeviceToHost));
  long long bi = 1LL << 60, bt = 1LL << 60;
  for (int r = 2; r < 8; ++r) { if (t[2 * r] < bi) bi = t[2 * r]; if (t[2 * r + 1] < bt) bt = t[2 * r + 1]; }
  printf("unrolled %-10s M=%3d N=%3d TS=%d nmma=%3d sameD=%d : issue %6lld cyc, complete %6lld cyc (%.1f / MMA)\n", name, M, N, TS, NM, SAMED, bi, bt, (double)bt / NM);
}

static void run_decode(const char* name, DecodeCfg c, float* d_out, std::vector<float>& h) {
  CK(cudaMemset(d_out, 0xff, 2 * 128 * 256 * sizeof(float)));
  decode_kernel<<<1, 128, 65536 + 1024>>>(c, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("[%s] CUDA error: %s\n", name, cudaGetErrorString(e)); exit(2); }
  CK(cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
  printf("== %s: which=%d mn=%d ltype=%d lbo=%u sbo=%u N=%d M=%d start=%u\n", name, c.which, c.mn, c.ltype, c.lbo, c.sbo, c.N, c.M, c.start);
  // word index read for element (idx, k): which=0 -> D[m=k][n=idx]; which=1 -> D[m=idx][n=k]
  const int cnt = c.which == 0 ? c.N : c.M;
  for (int k = 0; k < 8; ++k) {
    printf("  k=%d byte offsets:", k);
    for (int idx = 0; idx < cnt && idx < 40; ++idx) {
      const float lo = c.which == 0 ? h[(size_t)(0 * 128 + k) * 256 + idx] : h[(size_t)(0 * 128 + idx) * 256 + k];
      const float hi = c.which == 0 ? h[(size_t)(1 * 128 + k) * 256 + idx] : h[(size_t)(1 * 128 + idx) * 256 + k];
      const long w = (long)hi * 2048 + (long)lo;
      printf(" %ld", w * 4);
    }
    printf("\n");
  }
}

int main(int argc, char** argv) {
  CK(cudaSetDevice(0));
  if (argc == 2 && !strcmp(argv[1], "mn")) {   // timing of MN-major (layout type 1) B operands only
    long long* d_tm; CK(cudaMalloc(&d_tm, 16 * sizeof(long long)));
    run_time_mn<128, 32, 30>(40960, d_tm);
    run_time_mn<128, 64, 30>(40960, d_tm);
    run_time_mn<128, 96, 30>(40960, d_tm);
    run_time_mn<128, 64, 30>(10240, d_tm);
    run_time_mn<128, 64, 30>(4096, d_tm);
    run_time_mn<64, 32, 30>(40960, d_tm);
    run_time_mn<64, 64, 30>(40960, d_tm);
    printf("done\n");
    return 0;
  }
  if (argc == 9) {   // single decode: which mn ltype lbo sbo N M start   (one process per configuration: a bad descriptor faults)
    CK(cudaFuncSetAttribute(decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024));
    float* d_out1; CK(cudaMalloc(&d_out1, 2 * 128 * 256 * sizeof(float)));
    std::vector<float> h1(2 * 128 * 256);
    DecodeCfg c{atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), (uint32_t)atoi(argv[4]), (uint32_t)atoi(argv[5]), atoi(argv[6]), atoi(argv[7]), (uint32_t)atoi(argv[8])};
    run_decode("cli", c, d_out1, h1);
    return 0;
  }
  CK(cudaFuncSetAttribute(decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024));
  CK(cudaFuncSetAttribute(time_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  float* d_out; CK(cudaMalloc(&d_out, 2 * 128 * 256 * sizeof(float)));
  std::vector<float> h(2 * 128 * 256);

  // known-good reference: K-major no swizzle B, LBO = 128, SBO = 256
  run_decode("B K-major none", DecodeCfg{0, 0, 0, 128, 256, 32, 128, 0}, d_out, h);
  // MN-major no swizzle (expected broken for tf32)
  run_decode("B MN-major none", DecodeCfg{0, 1, 0, 1024, 128, 32, 128, 0}, d_out, h);
  // MN-major layout type 1 (128B swizzle, 32B base): atoms of 4 K-rows x 128 B; SBO between K atoms, LBO between 32-element MN groups
  run_decode("B MN-major type1 lbo=4096 sbo=512", DecodeCfg{0, 1, 1, 4096, 512, 64, 128, 0}, d_out, h);
  run_decode("B MN-major type1 lbo=512 sbo=4096", DecodeCfg{0, 1, 1, 512, 4096, 64, 128, 0}, d_out, h);
  run_decode("B MN-major type1 lbo=4096 sbo=1024", DecodeCfg{0, 1, 1, 4096, 1024, 64, 128, 0}, d_out, h);
  run_decode("B MN-major type2 (SW128) lbo=4096 sbo=1024", DecodeCfg{0, 1, 2, 4096, 1024, 64, 128, 0}, d_out, h);
  run_decode("B MN-major type1 start=512", DecodeCfg{0, 1, 1, 4096, 512, 64, 128, 512}, d_out, h);
  run_decode("B MN-major type1 start=128", DecodeCfg{0, 1, 1, 4096, 512, 64, 128, 128}, d_out, h);
  // K-major with layout type 1 / 2 / 4 / 6: what does the hardware do?
  run_decode("B K-major type2 (SW128) sbo=1024", DecodeCfg{0, 0, 2, 16, 1024, 32, 128, 0}, d_out, h);
  run_decode("B K-major type4 (SW64) sbo=512", DecodeCfg{0, 0, 4, 16, 512, 32, 128, 0}, d_out, h);
  run_decode("B K-major type6 (SW32) sbo=256", DecodeCfg{0, 0, 6, 16, 256, 32, 128, 0}, d_out, h);
  // A operand probes
  run_decode("A K-major none", DecodeCfg{1, 0, 0, 128, 256, 16, 128, 0}, d_out, h);
  run_decode("A MN-major type1 lbo=4096 sbo=512", DecodeCfg{1, 1, 1, 4096, 512, 16, 128, 0}, d_out, h);
  // M = 64: where do the rows land?  (which=1, idx = m up to 64; read lanes 0..127)
  {
    DecodeCfg c{1, 0, 0, 128, 256, 16, 64, 0};
    CK(cudaMemset(d_out, 0xff, 2 * 128 * 256 * sizeof(float)));
    decode_kernel<<<1, 128, 65536 + 1024>>>(c, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("[M64] CUDA error: %s\n", cudaGetErrorString(e)); exit(2); }
    CK(cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
    printf("== M=64 D lane map (lane: word read for column 0 / column 1)\n ");
    for (int lane = 0; lane < 128; ++lane) printf(" %d:%g/%g", lane, h[(size_t)lane * 256 + 0], h[(size_t)lane * 256 + 1]);
    printf("\n");
  }

  // ---- timing ----
  long long* d_t; CK(cudaMalloc(&d_t, 16 * sizeof(long long)));
  auto run_time = [&](const char* name, TimeCfg c) {
    time_kernel<<<1, 128, 220 * 1024>>>(c, d_t);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("[%s] CUDA error: %s\n", name, cudaGetErrorString(e)); exit(2); }
    long long t[16]; CK(cudaMemcpy(t, d_t, sizeof(t), cudaMemcpyDeviceToHost));
    long long bi = 1LL << 60, bt = 1LL << 60;
    for (int r = 2; r < 8; ++r) { if (t[2 * r] < bi) bi = t[2 * r]; if (t[2 * r + 1] < bt) bt = t[2 * r + 1]; }
    printf("time %-44s M=%3d N=%3d nmma=%3d : issue %6lld cyc, complete %6lld cyc (%.1f / MMA)\n", name, c.M, c.N, c.nmma, bi, bt, (double)bt / c.nmma);
  };
  // jet-kernel forward: A weight image (LBO 128, SBO 2560, step 256), B R image (LBO 144, SBO 2880, step 288), K = 80 -> 10 k-steps
  for (int N : {16, 32, 64, 128, 256}) {
    run_time("SS fwd (A 128x8 LBO128/SBO2560, B LBO144)", TimeCfg{128, N, 30, 0, 128, 2560, 256, 144, 2880, 288, 10});
    run_time("TS fwd (A in TMEM)", TimeCfg{128, N, 30, 1, 128, 2560, 256, 144, 2880, 288, 10});
  }
  for (int N : {32, 64, 128}) run_time("SS M=64", TimeCfg{64, N, 30, 0, 128, 2560, 256, 144, 2880, 288, 10});
  // wgrad-like: A / B C-images (SBO 128, LBO 1280, step 2560), N = 80, K-steps = 4
  run_time("SS wgrad (C images) N=80", TimeCfg{128, 80, 12, 0, 1280, 128, 2560, 1280, 128, 2560, 4});
  run_time("TS wgrad (A in TMEM) N=80", TimeCfg{128, 80, 12, 1, 1280, 128, 2560, 1280, 128, 2560, 4});
  run_time("SS N=80 x120", TimeCfg{128, 80, 120, 0, 128, 2560, 256, 144, 2880, 288, 10});
  run_time("SS N=32 x120", TimeCfg{128, 32, 120, 0, 128, 2560, 256, 144, 2880, 288, 10});
  run_time("TS N=32 x120", TimeCfg{128, 32, 120, 1, 128, 2560, 256, 144, 2880, 288, 10});
  run_time("SS N=128 x120", TimeCfg{128, 128, 120, 0, 128, 2560, 256, 144, 2880, 288, 10});
  run_time("TS N=128 x120", TimeCfg{128, 128, 120, 1, 128, 2560, 256, 144, 2880, 288, 10});
  run_time("SS N=256 x120", TimeCfg{128, 256, 120, 0, 128, 2560, 256, 144, 2880, 288, 10});

  run_time_t<128, 16, 0, 30, 1>("SS", d_t);
  run_time_t<128, 32, 0, 30, 1>("SS", d_t);
  run_time_t<128, 64, 0, 30, 1>("SS", d_t);
  run_time_t<128, 128, 0, 30, 1>("SS", d_t);
  run_time_t<128, 32, 0, 60, 1>("SS", d_t);
  run_time_t<128, 32, 0, 60, 0>("SS", d_t);
  run_time_t<128, 64, 0, 60, 0>("SS", d_t);
  run_time_t<128, 128, 0, 60, 0>("SS", d_t);
  run_time_t<128, 80, 0, 60, 1>("SS", d_t);
  run_time_t<64, 32, 0, 60, 1>("SS", d_t);
  run_time_t<64, 128, 0, 60, 1>("SS", d_t);
  run_time_t<128, 16, 1, 30, 1>("TS", d_t);
  run_time_t<128, 32, 1, 30, 1>("TS", d_t);
  run_time_t<128, 64, 1, 30, 1>("TS", d_t);
  run_time_t<128, 128, 1, 30, 1>("TS", d_t);
  run_time_t<128, 32, 1, 60, 1>("TS", d_t);
  run_time_t<128, 32, 1, 60, 0>("TS", d_t);
  run_time_t<128, 80, 1, 60, 1>("TS", d_t);
  printf("done\n");
  return 0;
}
