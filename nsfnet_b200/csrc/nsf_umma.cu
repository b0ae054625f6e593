// tcgen05 (UMMA) path of libnsf_b200.so.
//
// nsf_selftest_umma: one CTA computes D[128,n] = A[128,k] * B[n,k]^T with kind::tf32 MMAs using
// exactly the shared-memory layouts / descriptor strides of the three contractions of the jet kernel:
//   variant 2 "wgrad"        :  A, B K-major, the contraction index runs over 16-byte chunks LBO apart
//                               (point blocks), 8-row groups are 128 B apart (SBO = 128)
//   variant 3 "forward/dgrad":  A = weight image (K-major, LBO = 128, SBO = one 8-row band),
//                               B = activation / adjoint rows with the chunk stride padded to 144 B so that
//                               the epilogue's per-neuron scalar stores are bank-conflict free
//   variant 4 "A from TMEM"  :  A written to tensor memory with tcgen05.st (lane = row, column = K index), B as in
//                               variant 3; tcgen05.mma reads A from TMEM and only B from shared memory
// flag 16: 3xTF32 (lo*hi + hi*lo + hi*hi, fp32 accumulation in TMEM).
// Measured on B200: MN-major tf32 operands in the no-swizzle layout produce zeros (CUTLASS: "for mn-major
// tf32 operands, SW128_32B is the only available smem layout"), so every contraction of the jet kernel is
// arranged to be K-major.  tests/test_gpu_umma.py checks the variants against a float64 product.
#include "nsf_internal.h"
#include "nsf_tc.cuh"

using namespace nsftc;

struct SelftestArgs {
  const float* a; const float* b; float* d;
  int n, k, variant, flags;
};

// byte offset of logical element (r, c) of an operand image; r = M/N index, c = contraction index
__device__ __forceinline__ uint32_t img_off_kmajor(int r, int c, uint32_t sbo, uint32_t lbo) {
  return (uint32_t)(r >> 3) * sbo + (uint32_t)(r & 7) * 16u + (uint32_t)(c >> 2) * lbo + (uint32_t)(c & 3) * 4u;
}
__device__ __forceinline__ uint32_t img_off_mnmajor(int r, int c, uint32_t sbo, uint32_t lbo) {
  return (uint32_t)(r >> 2) * sbo + (uint32_t)(r & 3) * 4u + (uint32_t)(c & 7) * 16u + (uint32_t)(c >> 3) * lbo;
}

__global__ void __launch_bounds__(128) nsf_selftest_kernel(SelftestArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int N = a.n, K = a.k, M = 128;
  const bool split3 = (a.flags & 16) != 0;

  // layouts ---------------------------------------------------------------------------------
  int a_mn, b_mn;
  uint32_t a_sbo, a_lbo, a_step, b_sbo, b_lbo, b_step, a_bytes, b_bytes;
  if (a.variant == 3 || a.variant == 4) {  // forward / dgrad: A rows j, K chunks 128 B apart; B rows n, K chunks 144 B apart
    a_mn = 0; a_lbo = 128; a_sbo = (uint32_t)(K / 4) * 128; a_step = 256; a_bytes = (uint32_t)(M / 8) * a_sbo;
    b_mn = 0; b_lbo = 144; b_sbo = (uint32_t)(K / 4) * 144; b_step = 288; b_bytes = (uint32_t)(N / 8) * b_sbo;
  } else {  // wgrad: A (j, n) K-major with point blocks as K chunks, B (k, n) likewise
    a_mn = 0; a_sbo = 128; a_lbo = (uint32_t)(M / 8) * 128; a_step = 2 * a_lbo; a_bytes = (uint32_t)(K / 4) * a_lbo;
    b_mn = 0; b_sbo = 128; b_lbo = (uint32_t)(N / 8) * 128; b_step = 2 * b_lbo; b_bytes = (uint32_t)(K / 4) * b_lbo;
  }
  unsigned char* a_hi = smem;
  unsigned char* a_lo = a_hi + a_bytes;
  unsigned char* b_hi = a_lo + a_bytes;
  unsigned char* b_lo = b_hi + b_bytes;

  for (int i = tid; i < M * K; i += 128) {
    const int r = i / K, c = i % K;
    float hi, lo;
    split_tf32(a.a[i], hi, lo);
    if (!split3) { hi = a.a[i]; lo = 0.f; }
    const uint32_t off = a_mn ? img_off_mnmajor(r, c, a_sbo, a_lbo) : img_off_kmajor(r, c, a_sbo, a_lbo);
    *reinterpret_cast<float*>(a_hi + off) = hi;
    *reinterpret_cast<float*>(a_lo + off) = lo;
  }
  for (int i = tid; i < N * K; i += 128) {
    const int r = i / K, c = i % K;
    float hi, lo;
    split_tf32(a.b[i], hi, lo);
    if (!split3) { hi = a.b[i]; lo = 0.f; }
    const uint32_t off = b_mn ? img_off_mnmajor(r, c, b_sbo, b_lbo) : img_off_kmajor(r, c, b_sbo, b_lbo);
    *reinterpret_cast<float*>(b_hi + off) = hi;
    *reinterpret_cast<float*>(b_lo + off) = lo;
  }
  uint32_t ncols = 32;
  const int need = a.variant == 4 ? N + 2 * K : N;
  while ((int)ncols < need) ncols <<= 1;
  if (warp == 0) tmem_alloc(&tmem_base, ncols);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base;

  if (a.variant == 4) {
    // A -> TMEM: thread (row) writes its K values, hi at columns [N, N+K), lo at [N+K, N+2K)
    for (int c0 = 0; c0 < K; c0 += 8) {
      float hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float v = a.a[(size_t)tid * K + c0 + i];
        split_tf32(v, hi[i], lo[i]);
        if (!split3) { hi[i] = v; lo[i] = 0.f; }
      }
      tmem_st8(tb + ((uint32_t)(warp * 32) << 16) + (uint32_t)(N + c0), hi);
      tmem_st8(tb + ((uint32_t)(warp * 32) << 16) + (uint32_t)(N + K + c0), lo);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
      const uint32_t leader = elect_one();
      const uint32_t idesc = idesc_tf32(M, N, 0, 0);
      const uint32_t bhi = desc_hi(b_sbo);
      uint32_t acc = 0;
      for (int ks = 0; ks < K / 8; ++ks) {
        const uint32_t bh = desc_lo(smem_u32(b_hi) + ks * b_step, b_lbo), bl = desc_lo(smem_u32(b_lo) + ks * b_step, b_lbo);
        const uint32_t ah = tb + (uint32_t)(N + ks * 8), al = tb + (uint32_t)(N + K + ks * 8);
        if (split3) {
          mma_tf32_ts_elect(tb, al, bh, bhi, idesc, acc, leader); acc = 1;
          mma_tf32_ts_elect(tb, ah, bl, bhi, idesc, acc, leader);
        }
        mma_tf32_ts_elect(tb, ah, bh, bhi, idesc, acc, leader); acc = 1;
      }
      mma_commit_elect(&bar, leader);
    }
  } else if (tid == 0) {
    const uint32_t alb = a_lbo, asb = a_sbo, blb = b_lbo, bsb = b_sbo;
    const uint32_t idesc = idesc_tf32(M, N, a_mn, b_mn);
    uint32_t acc = 0;
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint64_t ah = smem_desc(smem_u32(a_hi) + ks * a_step, alb, asb);
      const uint64_t al = smem_desc(smem_u32(a_lo) + ks * a_step, alb, asb);
      const uint64_t bh = smem_desc(smem_u32(b_hi) + ks * b_step, blb, bsb);
      const uint64_t bl = smem_desc(smem_u32(b_lo) + ks * b_step, blb, bsb);
      if (split3) {
        mma_tf32(tb, al, bh, idesc, acc); acc = 1;   // small terms first
        mma_tf32(tb, ah, bl, idesc, acc);
      }
      mma_tf32(tb, ah, bh, idesc, acc); acc = 1;
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  // epilogue: warp w reads TMEM lanes 32w..32w+31 (= rows), 8 columns at a time
  const int row = tid;
  for (int c0 = 0; c0 < N; c0 += 8) {
    float v[8];
    tmem_ld8(tb + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) a.d[(size_t)row * N + c0 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, ncols);
}

extern "C" int nsf_selftest_umma(int device, int32_t variant, const float* a, const float* b, float* d, int32_t n, int32_t k,
                                 void* stream) {
  const int v = variant & 15, flags = variant & ~15;
  if (!a || !b || !d || (v < 2 || v > 4) || (flags & ~16) || n < 16 || n > 256 || (n % 16) || k < 8 || (k % 8)) {
    nsf_set_error("nsf_selftest_umma: bad argument (variant 2, 3 or 4 [+16], n in [16,256] multiple of 16, k multiple of 8)");
    return NSF_E_ARG;
  }
  cudaDeviceProp prop;
  NSF_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) { nsf_set_error("nsf_selftest_umma: device is not sm_100"); return NSF_E_ARCH; }
  NSF_CUDA_OK(cudaSetDevice(device));
  size_t a_bytes, b_bytes;
  if (v == 3 || v == 4) { a_bytes = (size_t)16 * (k / 4) * 128; b_bytes = (size_t)(n / 8) * (k / 4) * 144; }
  else { a_bytes = (size_t)(k / 4) * 16 * 128; b_bytes = (size_t)(k / 4) * (n / 8) * 128; }
  const size_t smem = 2 * a_bytes + 2 * b_bytes + 1024;
  if (smem > 220 * 1024) { nsf_set_error("nsf_selftest_umma: operands do not fit in shared memory (%zu bytes)", smem); return NSF_E_SHAPE; }
  NSF_CUDA_OK(cudaFuncSetAttribute(nsf_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SelftestArgs args{a, b, d, n, k, v, flags};
  nsf_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(args);
  NSF_CUDA_OK(cudaGetLastError());
  return NSF_OK;
}
